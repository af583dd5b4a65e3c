#!/usr/bin/env python
"""bench.py — QGT evaluations per second on synthetic hardware-efficient ansätze (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--impl ours|reference]

A step = one full QGT evaluation (metric + Berry curvature) of the workload's circuit through the C-ABI.
  value  whole-job evals/s from CUDA-event device time (library stream), max over ranks
  e2e    the same metric through the public call with HOST buffers: theta and the gate table go in,
         metric and Berry curvature (2 * P^2 doubles) come back, wall clock between device syncs
  roofline      dominant kernel (gate sweep: HBM; Gram: FP64 tensor pipe) against measured peaks
  cpu_baseline  the reference's own CPU routines (oracle/_ref: sim_execute_circuit + diffgeo_*) or, where
                the reference tree was never built, the oracle port, on a bounded sample, extrapolated
N > 1: independent replicas of the workload at different parameter points, one per rank (weak scaling);
workloads too large for one GPU (c5) use the sharded path.
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
NOMINAL_FP64_TENSOR_TFLOPS = 40.0  # B200 nominal; replaced by the measured DMMA peak when profiles/ has it


def load_peaks():
    peaks = {"hbm_gbs": FALLBACK_HBM_GBS, "hbm_source": "fallback", "dmma_tflops": NOMINAL_FP64_TENSOR_TFLOPS,
             "dmma_source": "nominal"}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            if "hbm_gbs" in j:
                peaks["hbm_gbs"], peaks["hbm_source"] = float(j["hbm_gbs"]), "measured"
            if "fp64_tensor_tflops" in j:
                peaks["dmma_tflops"], peaks["dmma_source"] = float(j["fp64_tensor_tflops"]), "measured"
        except Exception:
            pass
    p = os.path.join(ROOT, "profiles", "peaks.json")
    if os.path.exists(p) and peaks["dmma_source"] != "measured":
        try:
            j = json.load(open(p))
            peaks["dmma_tflops"], peaks["dmma_source"] = float(j["dmma_tflops"]), "measured (tools/peaks.cu, profiles/peaks.json)"
        except Exception:
            pass
    return peaks


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


def cpu_baseline(circ, theta, budget_s: float = 20.0):
    """Reference CPU path on a bounded sample: forward circuit, k derivative columns, a k x k g/F assembly,
    extrapolated to P columns (columns ~ P, assembly ~ P^2).  1 core: the reference path is serial."""
    from oracle.oracle import Oracle, Reference
    kind = "reference" if Reference.available() else "port"
    eng = Reference() if kind == "reference" else Oracle()
    if kind == "reference" and (circ.initial_state != 0 or any(g[0] not in circuits.REFERENCE_KINDS for g in circ.gates)):
        kind, eng = "port", Oracle()     # the reference simulator has no such gate: time the restatement
    P = circ.num_params
    t0 = time.perf_counter()
    psi = eng.apply(circ, theta)
    t_fwd = time.perf_counter() - t0
    cols, t_cols = [], 0.0
    k = 0
    while k < P and (k < 2 or t_cols + t_fwd < budget_s * 0.6) and k < 16:
        t0 = time.perf_counter()
        cols.append(eng.derivative(circ, theta, k))
        t_cols += time.perf_counter() - t0
        k += 1
    J = np.stack(cols)
    t0 = time.perf_counter()
    if kind == "reference":
        eng.fubini_berry(psi, J)
    else:
        eng.qgt_from_columns(psi, J)
    t_asm = time.perf_counter() - t0
    total = t_fwd + (t_cols / k) * P + t_asm * (P / k) ** 2
    full = (k == P)
    return {"value": 1.0 / total, "unit": "QGT evals/s", "cores": 1, "kind": kind,
            "sample": (f"forward circuit {t_fwd:.2f}s + {k} of {P} derivative columns {t_cols:.2f}s + {k}x{k} metric/curvature "
                       f"assembly {t_asm:.2f}s" + ("" if full else f"; extrapolated to P={P} (columns x{P / k:.1f}, assembly x{(P / k) ** 2:.1f})")),
            "seconds_per_eval": total}


def run_reference(args, circ, theta):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    base = None
    for i in range(args.warmup + args.steps):
        base = cpu_baseline(circ, theta, budget_s=6.0)
        if i >= args.warmup:
            vals.append(base["seconds_per_eval"])
    sec = float(np.mean(vals))
    line = {"impl": "reference", "metric": "QGT evals/sec", "value": 1.0 / sec, "unit": "QGT evals/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex double)", "data": "synthetic",
            "config": workload_config(args, circ),
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": 1.0 / sec, "unit": "QGT evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    line["cpu_baseline"]["value"] = line["value"]
    print(json.dumps(line))


def workload_config(args, circ):
    return {"workload": f"{args.workload}: {circ.name}", "qubits": circ.num_qubits, "params": circ.num_params,
            "gates": len(circ.gates), "amplitude_type": "complex128",
            "l2": "derivative-column working set exceeds the 126 MB L2 (no flush needed)"
                  if (circ.num_params + 1) * (16 << circ.num_qubits) > (256 << 20) else "L2 flushed between steps"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sharded", action="store_true", help="N > 1: shard ONE state over the ranks (NCCL exchanges + Gram allreduce) "
                                                           "instead of running N independent replicas")
    ap.add_argument("--explore", action="store_true", help="allow fewer than 3 warm-up steps (exploration only, never a reported number)")
    ap.add_argument("--opt", action="append", default=[], help="library tuning option key=value (qgt_b200_set_option)")
    args = ap.parse_args()
    if args.impl == "ours" and not args.explore:
        args.warmup = max(args.warmup, 3)      # timing rule: at least 3 warm-up steps

    circ = circuits.config(args.workload)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    sharded = world > 1 and (args.sharded or args.workload == "c5")
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
    if sharded:                            # one state over all ranks: rank 0's NCCL id goes round over torch.distributed
        uid = [api.Context.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.dist_init(rank, world, uid[0])
    P = circ.num_params
    flush = None
    if "flushed" in workload_config(args, circ)["l2"]:
        flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

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

    for _ in range(args.warmup):
        if flush is not None:
            flush.zero_()
        step()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    dev_ms = 0.0
    agg = {"ms_sweep": 0.0, "ms_gram": 0.0, "ms_other": 0.0, "sweep_bytes": 0.0, "gram_flops": 0.0, "gram_bytes": 0.0,
           "sweep_launches": 0, "gram_launches": 0, "other_launches": 0, "sweep_column_passes": 0}
    wall = 0.0
    st = None
    for _ in range(args.steps):
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
    if rank == 0:
        peaks = load_peaks()
        evals = args.steps * (1 if sharded else world)
        value = evals / (dev_ms_max * 1e-3)
        e2e = evals / (wall_ms_max * 1e-3)
        sweep_gbs = agg["sweep_bytes"] / (agg["ms_sweep"] * 1e-3) * 1e-9 if agg["ms_sweep"] > 0 else 0.0
        gram_tf = agg["gram_flops"] / (agg["ms_gram"] * 1e-3) * 1e-12 if agg["ms_gram"] > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(args.workload)
            except Exception:
                traffic = None
        sweep_roof = {"bound": "hbm", "kernel": "qgt_sweep_kernel", "achieved": sweep_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                      "frac": sweep_gbs / peaks["hbm_gbs"], "traffic": (traffic or {}).get("sweep_bytes_per_launch"),
                      "peak_source": peaks["hbm_source"], "share_of_step": agg["ms_sweep"] / dev_ms,
                      "launches_per_step": agg["sweep_launches"] / args.steps,
                      "algorithmic_bytes_per_launch": agg["sweep_bytes"] / max(1, agg["sweep_launches"])}
        gram_roof = {"bound": "tensor", "kernel": "qgt_gram_kernel", "achieved": gram_tf, "peak": peaks["dmma_tflops"], "unit": "TFLOP/s",
                     "frac": gram_tf / peaks["dmma_tflops"], "traffic": (traffic or {}).get("gram_bytes_per_launch"),
                     "peak_source": peaks["dmma_source"], "share_of_step": agg["ms_gram"] / dev_ms,
                     "launches_per_step": agg["gram_launches"] / args.steps,
                     "algorithmic_flops_per_launch": agg["gram_flops"] / max(1, agg["gram_launches"])}
        dominant, other = (sweep_roof, gram_roof) if agg["ms_sweep"] >= agg["ms_gram"] else (gram_roof, sweep_roof)
        line = {"metric": "QGT evals/sec", "value": value, "unit": "QGT evals/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
                "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f64 (complex double)", "data": "synthetic",
                "config": dict(workload_config(args, circ),
                               parallelism=(f"state sharded over {world} GPUs (top qubits), NCCL exchange + allreduce" if sharded
                                            else f"replicas x{world} (independent parameter points)" if world > 1 else "single GPU"),
                               runs=st["num_runs"], resident_columns=st["resident_columns"], blocks=st["blocks"],
                               tile_qubits=st["tile_qubits"]),
                "e2e": {"value": e2e, "unit": "QGT evals/s", "h2d_bytes_per_step": 8 * P + 32 * len(circ.gates),
                        "d2h_bytes_per_step": 16 * P * P,
                        "includes": "host planning, H2D of the plan, D2H of metric + Berry curvature, natural-gradient solve on the host",
                        "natgrad_ms_per_step": natgrad_ms[0] / (args.steps + args.warmup)},
                "gpu_launches": int(agg["sweep_launches"] + 2 * agg["gram_launches"] + agg["other_launches"]),
                "clocks": clocks, "roofline": dominant, "roofline_secondary": other,
                "gate_bw_effective_gbs": circ.unfused_bytes() * 0 + 0.0}
        # effective per-gate bandwidth: what one-pass-per-gate would have had to move, over the sweep time
        passes = agg["sweep_column_passes"] / args.steps
        line["gate_bw_effective_gbs"] = (circ.unfused_bytes() / max(1, st["num_runs"])) * passes / (agg["ms_sweep"] / args.steps * 1e-3) * 1e-9 \
            if agg["ms_sweep"] > 0 else 0.0
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
