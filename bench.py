#!/usr/bin/env python
"""bench.py — QGT evaluations per second on synthetic hardware-efficient ansätze (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--impl ours|reference]

A step = one full QGT evaluation (metric + Berry curvature) of the workload's circuit through the C-ABI.
Default workload: c3 (28 qubits, 256 parameters, complex double) = the largest single-GPU configuration of
BASELINE.json.  N > 1 (torchrun): the SAME state sharded over the N ranks on its top qubits (strong scaling, NCCL
exchanges + allreduce); --replicas runs N independent evaluations instead.

  value         whole-job evals/s from CUDA-event device time (library stream), max over ranks
  e2e           the same metric through the public call with HOST buffers: theta and the gate table go in, metric
                and Berry curvature (2 * P^2 doubles) come back, then the regularised natural-gradient solve on the
                host; wall clock between device syncs
  roofline      dominant kernel against peaks measured in this run (FP64 tensor pipe: DMMA micro-kernel of the
                library; HBM: MEASURED_PEAKS.json, else the library's copy measurement)
  cpu_baseline  the reference's own CPU routines (oracle/_ref: sim_execute_circuit + diffgeo_*) on a bounded
                sample, timed ONCE, extrapolated (the sample says how)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from quantum_geometric_tensor_b200 import circuits  # noqa: E402

FALLBACK_HBM_GBS = 6650.0          # B200_PROFILING.md fallback
NOMINAL_FP64_TENSOR_TFLOPS = 40.0  # B200 nominal
NVLINK_GBS_PER_DIR = 900.0         # NVLink 5, per GPU and direction


def load_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            if "hbm_gbs" in j:
                return float(j["hbm_gbs"]), "MEASURED_PEAKS.json"
        except Exception:
            pass
    return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def workload_config(args, circ):
    """The same dict in both arms (the driver compares them)."""
    return {"workload": f"{args.workload}: {circ.name}", "qubits": circ.num_qubits, "params": circ.num_params,
            "gates": len(circ.gates), "amplitude_type": "complex128",
            "l2": "working set (>= 2 statevectors of %.1f GB) exceeds the 126 MB L2: no flush needed" % (16 * 2.0 ** circ.num_qubits / 1e9)
                  if 16 << circ.num_qubits > (128 << 20) else "L2 flushed between steps"}


def cpu_baseline(circ, theta, budget_s: float = 25.0):
    """The reference's CPU path on a bounded sample, timed once.  1 core: the path is serial by construction
    (no OpenMP in quantum_simulator.c / differential_geometry.c).

    n <= 22: forward circuit + k derivative columns + a k x k metric/curvature assembly, extrapolated to P columns.
    n  > 22: the first g gates of the circuit on the full 2^n state (per-gate time x gate count = one circuit; a
             derivative column costs one circuit), and the assembly of a 2 x 2 block on a 2^m-amplitude slice,
             scaled by (P/2)^2 and 2^(n-m)."""
    from oracle.oracle import Oracle, Reference
    kind = "reference" if Reference.available() else "port"
    eng = Reference() if kind == "reference" else Oracle()
    if kind == "reference" and (circ.initial_state != 0 or any(g[0] not in circuits.REFERENCE_KINDS for g in circ.gates)):
        kind, eng = "port", Oracle()     # the reference simulator has no such gate: time the restatement
    P, n, G = circ.num_params, circ.num_qubits, len(circ.gates)
    if n <= 22:
        t0 = time.perf_counter()
        psi = eng.apply(circ, theta)
        t_fwd = time.perf_counter() - t0
        cols, t_cols, k = [], 0.0, 0
        while k < P and (k < 2 or t_cols + t_fwd < budget_s * 0.6) and k < 16:
            t0 = time.perf_counter()
            cols.append(eng.derivative(circ, theta, k))
            t_cols += time.perf_counter() - t0
            k += 1
        J = np.stack(cols)
        t0 = time.perf_counter()
        eng.fubini_berry(psi, J) if kind == "reference" else eng.qgt_from_columns(psi, J)
        t_asm = time.perf_counter() - t0
        total = t_fwd + (t_cols / k) * P + t_asm * (P / k) ** 2
        sample = (f"forward {t_fwd:.2f}s + {k}/{P} columns {t_cols:.2f}s + {k}x{k} assembly {t_asm:.2f}s"
                  + ("" if k == P else f", columns x{P / k:.0f}, assembly x{(P / k) ** 2:.0f}"))
    else:
        g = 6
        sub = circuits.Circuit(n, gates=list(circ.gates[:g]), num_params=circ.num_params)
        st = np.zeros(1 << n, dtype=np.complex128)
        st[0] = 1.0
        t0 = time.perf_counter()
        st = eng.apply(sub, theta, st)
        t_g = (time.perf_counter() - t0) / g
        m = min(n, 26)
        sl = slice(0, 1 << m)
        J = np.stack([st[sl], st[sl][::-1]])
        t0 = time.perf_counter()
        eng.fubini_berry(st[sl], J) if kind == "reference" else eng.qgt_from_columns(st[sl], J)
        t_asm = (time.perf_counter() - t0) * 2.0 ** (n - m)
        total = t_g * G * (P + 1) + t_asm * (P / 2) ** 2
        sample = (f"{g}/{G} gates on 2^{n} amplitudes {t_g * g:.1f}s (circuit = x{G / g:.0f}, {P}+1 circuits) + 2x2 assembly on "
                  f"2^{m} amplitudes x{2 ** (n - m)} x{(P / 2) ** 2:.0f}: extrapolated")
    return {"value": 1.0 / total, "unit": "QGT evals/s", "cores": 1, "kind": kind, "sample": sample, "seconds_per_eval": total}


def run_reference(args, circ, theta):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    base = cpu_baseline(circ, theta)          # ONE bounded sample; steps / warmup do not repeat minutes of CPU work
    sec = base["seconds_per_eval"]
    line = {"impl": "reference", "metric": "QGT evals/sec", "value": 1.0 / sec, "unit": "QGT evals/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "strong" if args.gpus > 1 and not args.replicas else "weak", "vs_baseline": None,
            "dtype": "f64 (complex double)", "data": "synthetic", "config": workload_config(args, circ),
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": 1.0 / sec, "unit": "QGT evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "timing": "one bounded sample extrapolated to a full evaluation (see cpu_baseline.sample); not repeated per step"}
    print(json.dumps(line))


def golden_parity(ctx, api, name="c1_hea_n12_l2"):
    """C1 through the very same context (sharded when the bench is) against the committed reference golden."""
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    c = circuits.Circuit(int(z["n"]))
    for row in z["gates"]:
        c.add(int(row[0]), int(row[1]), int(row[2]), int(row[3]), float(row[4]), float(row[5]))
    q = ctx.qgt(c, z["theta"])
    e_g = np.abs(q.real - z["metric"]).max() / np.abs(z["metric"]).max()
    e_f = np.abs(-2 * q.imag - z["curvature"]).max() / np.abs(z["curvature"]).max()
    return {"case": name + " (reference-generated golden, tests/golden)", "rel_err_metric": float(e_g), "rel_err_curvature": float(e_f),
            "ok": bool(e_g < 1e-10 and e_f < 1e-10)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--replicas", action="store_true", help="N > 1: N independent evaluations (weak scaling) instead of one sharded state")
    ap.add_argument("--sharded", action="store_true", help="(default for N > 1; kept for compatibility)")
    ap.add_argument("--max-seconds", type=float, default=1500.0, help="cap on the timed region: fewer steps are timed (and reported) if K would exceed it")
    ap.add_argument("--explore", action="store_true", help="allow fewer than 3 warm-up steps (exploration only, never a reported number)")
    ap.add_argument("--opt", action="append", default=[], help="library tuning option key=value (qgt_b200_set_option)")
    args = ap.parse_args()
    if args.impl == "ours" and not args.explore:
        args.warmup = max(args.warmup, 3)      # timing rule: at least 3 warm-up steps

    circ = circuits.config(args.workload)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    sharded = world > 1 and not args.replicas
    theta = circuits.default_angles(circ.num_params, circuits.SEED_ANGLES + (0 if sharded else rank))

    if args.impl == "reference":
        run_reference(args, circ, theta)
        return

    import torch
    import torch.distributed as dist
    from quantum_geometric_tensor_b200 import api

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    ctx = api.Context(local_rank)          # raises if the CUDA library or an sm_100 device is missing
    for kv in args.opt:
        k, v = kv.split("=")
        ctx.set_option(k, float(v))
    peaks = ctx.measure_peaks()            # before the workspace is taken: DMMA micro-kernel + 1 GiB D2D copy
    if sharded:                            # one state over all ranks: rank 0's NCCL id goes round over torch.distributed
        uid = [api.Context.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.dist_init(rank, world, uid[0])
    P = circ.num_params
    cfg = workload_config(args, circ)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda") if "flushed" in cfg["l2"] else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    g_host = np.zeros((P, P))
    b_host = np.zeros((P, P))

    import ctypes as C
    cc = circ.to_c()                      # the host-side gate table (POD structs of include/qgt_b200.h)

    # BASELINE config 2 is "QGT + natural-gradient step": the e2e call sequence ends with the regularised solve
    # (G + lambda I)^-1 grad on the host (reference semantics), fed with a fixed synthetic gradient
    grad_host = np.random.default_rng(7).normal(size=P)
    dx_host = np.zeros(P)
    lam = C.c_double(0)
    natgrad_ms = [0.0]

    def step():
        # the public call: host theta + gate table in, host metric + Berry curvature out
        rc = ctx.L.qgt_b200_qgt(ctx.h, C.byref(cc), theta.ctypes.data_as(api._DP), g_host.ctypes.data_as(api._DP),
                                b_host.ctypes.data_as(api._DP), None, None)
        if rc:
            raise api.QgtError(rc, ctx.L.qgt_b200_last_error().decode())
        st_ = ctx.stats()
        if rank == 0 or not sharded:
            t_ng = time.perf_counter()
            rc = ctx.L.qgt_b200_natural_gradient(ctx.h, g_host.ctypes.data_as(api._DP), grad_host.ctypes.data_as(api._DP), P, None,
                                                 dx_host.ctypes.data_as(api._DP), C.byref(lam))
            natgrad_ms[0] += (time.perf_counter() - t_ng) * 1e3
            if rc:
                raise api.QgtError(rc, ctx.L.qgt_b200_last_error().decode())
        return st_

    t_w0 = time.perf_counter()
    for _ in range(args.warmup):
        if flush is not None:
            flush.zero_()
        step()
    t_step_est = (time.perf_counter() - t_w0) / max(1, args.warmup)
    steps = args.steps
    if t_step_est * steps > args.max_seconds:
        steps = max(1, int(args.max_seconds / t_step_est))
    if world > 1:                          # every rank must time the same number of steps
        ts = torch.tensor([steps], dtype=torch.int64, device="cuda")
        dist.all_reduce(ts, op=dist.ReduceOp.MIN)
        steps = int(ts[0])
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    dev_ms = 0.0
    keys = ("ms_sweep", "ms_gram", "ms_other", "ms_exchange", "sweep_bytes", "gram_flops", "gram_bytes", "tensor_flops", "exchange_bytes",
            "sweep_launches", "gram_launches", "other_launches", "sweep_column_passes")
    agg = {k: 0.0 for k in keys}
    wall = 0.0
    st = None
    for _ in range(steps):
        if flush is not None:
            flush.zero_()
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        st = step()
        wall += time.perf_counter() - t0
        dev_ms += st["ms_total"]
        for k in agg:
            agg[k] += st[k]
    barrier()
    clocks = sampler.stop()

    t = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_ms_max = float(t[0]), float(t[1])
    parity = golden_parity(ctx, api)       # after the timed region: C1 through the same (possibly sharded) context
    if rank == 0:
        hbm, hbm_src = load_hbm_peak()
        if hbm is None:
            hbm, hbm_src = peaks["copy_gbs"], "measured in this run (1 GiB device-to-device copy)"
        dmma, dmma_src = peaks["dmma_tflops"], "measured in this run (qgt_b200_measure_peaks: mma.sync m8n8k4 f64, 4 chains/warp)"
        evals = steps * (1 if sharded or world == 1 else world)
        value = evals / (dev_ms_max * 1e-3)
        e2e = evals / (wall_ms_max * 1e-3)
        fused = bool(st["fused"])
        ms_sw, ms_gr = max(agg["ms_sweep"], 1e-9), max(agg["ms_gram"], 1e-9)
        sweep_gbs = agg["sweep_bytes"] / (ms_sw * 1e-3) * 1e-9
        sweep_tf = agg["tensor_flops"] / (ms_sw * 1e-3) * 1e-12
        gram_tf = agg["gram_flops"] / (ms_gr * 1e-3) * 1e-12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(args.workload + ("_fused" if fused else ""))
            except Exception:
                traffic = None
        nl = max(1, agg["sweep_launches"])
        hbm_roof = {"bound": "hbm", "kernel": "qgt_fused_direct_kernel" if fused else "qgt_sweep_kernel", "achieved": sweep_gbs, "peak": hbm,
                    "unit": "GB/s", "frac": sweep_gbs / hbm, "traffic": (traffic or {}).get("sweep_bytes_per_launch"),
                    "traffic_source": "profiles/traffic.json (ncu dram__bytes of the same command, committed)" if traffic else None,
                    "peak_source": hbm_src, "share_of_step": agg["ms_sweep"] / dev_ms, "launches_per_step": agg["sweep_launches"] / steps,
                    "algorithmic_bytes_per_launch": agg["sweep_bytes"] / nl}
        if fused:
            # the fused kernel carries the stage applications AND the transition-matrix products: FP64-tensor bound
            roof = {"bound": "tensor", "kernel": "qgt_fused_direct_kernel (sweep + transition matrices)", "achieved": sweep_tf, "peak": dmma,
                    "unit": "TFLOP/s", "frac": sweep_tf / dmma, "traffic": (traffic or {}).get("sweep_bytes_per_launch"),
                    "peak_source": dmma_src, "peak_nominal": NOMINAL_FP64_TENSOR_TFLOPS, "share_of_step": agg["ms_sweep"] / dev_ms,
                    "launches_per_step": agg["sweep_launches"] / steps, "algorithmic_flops_per_launch": agg["tensor_flops"] / nl,
                    "flops_definition": "DMMA flops issued per amplitude: 64 per dense stage application (32 in diagonal-real form), 48 per "
                                        "transition-matrix stage (3M complex product)"}
            other = hbm_roof
        else:
            gram_roof = {"bound": "tensor", "kernel": "qgt_gram_kernel", "achieved": gram_tf, "peak": dmma, "unit": "TFLOP/s", "frac": gram_tf / dmma,
                         "traffic": (traffic or {}).get("gram_bytes_per_launch"), "peak_source": dmma_src, "peak_nominal": NOMINAL_FP64_TENSOR_TFLOPS,
                         "share_of_step": agg["ms_gram"] / dev_ms, "launches_per_step": agg["gram_launches"] / steps,
                         "algorithmic_flops_per_launch": agg["gram_flops"] / max(1, agg["gram_launches"])}
            roof, other = (hbm_roof, gram_roof) if agg["ms_sweep"] >= agg["ms_gram"] else (gram_roof, hbm_roof)
        line = {"metric": "QGT evals/sec", "value": value, "unit": "QGT evals/s", "n_gpus": world, "steps": steps, "warmup": args.warmup,
                "ms_per_step": dev_ms_max / steps, "higher_is_better": True, "scaling": "strong" if sharded else "weak",
                "vs_baseline": None, "dtype": "f64 (complex double)", "data": "synthetic", "config": cfg,
                "plan": {"parallelism": (f"one state sharded over {world} GPUs on its top qubits, NCCL exchange + allreduce" if sharded
                                         else f"replicas x{world} (independent parameter points)" if world > 1 else "single GPU"),
                         "schedule": "fused (transition matrices inside the sweeps)" if fused else "columns + Gram",
                         "runs": st["num_runs"], "resident_columns": st["resident_columns"], "blocks": st["blocks"],
                         "tile_qubits": st["tile_qubits"], "column_passes_per_step": agg["sweep_column_passes"] / steps},
                "e2e": {"value": e2e, "unit": "QGT evals/s", "h2d_bytes_per_step": 8 * P + 32 * len(circ.gates),
                        "d2h_bytes_per_step": 16 * P * P,
                        "includes": "host planning, H2D of the plan, D2H of metric + Berry curvature, natural-gradient solve on the host",
                        "natgrad_ms_per_step": natgrad_ms[0] / (steps + args.warmup)},
                "gpu_launches": int(agg["sweep_launches"] + 2 * agg["gram_launches"] + agg["other_launches"]),
                "clocks": clocks, "roofline": roof, "roofline_secondary": other,
                "peaks_measured": {"dmma_tflops": peaks["dmma_tflops"], "copy_gbs": peaks["copy_gbs"]},
                "parity_check": parity}
        if steps != args.steps:
            line["steps_requested"] = args.steps
            line["steps_note"] = f"timed region capped at --max-seconds {args.max_seconds:.0f}"
        passes = agg["sweep_column_passes"] / steps
        line["gate_bw_effective_gbs"] = (circ.unfused_bytes() / max(1, st["num_runs"])) * passes / (agg["ms_sweep"] / steps * 1e-3) * 1e-9 \
            if agg["ms_sweep"] > 0 else 0.0
        if sharded:
            ex_ms = max(agg["ms_exchange"], 1e-9)
            line["exchange"] = {"bytes_sent_per_rank_per_step": agg["exchange_bytes"] / steps, "ms_per_step": agg["ms_exchange"] / steps,
                                "gbs_per_direction": agg["exchange_bytes"] / (ex_ms * 1e-3) * 1e-9, "nvlink_peak_gbs_per_direction": NVLINK_GBS_PER_DIR,
                                "frac": agg["exchange_bytes"] / (ex_ms * 1e-3) * 1e-9 / NVLINK_GBS_PER_DIR,
                                "share_of_step": agg["ms_exchange"] / dev_ms,
                                "allreduce": "one ncclAllReduce of %d doubles per step (inside ms_other)" % (2 * (P + 1) * (P + 1))}
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = {k: v for k, v in cpu_baseline(circ, theta).items() if k != "seconds_per_eval"}
            except Exception as ex:  # the checker is optional infrastructure; the GPU numbers stand on their own
                line["cpu_baseline"] = {"value": None, "unit": "QGT evals/s", "cores": 1, "kind": "port", "sample": f"failed: {ex}"}
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
