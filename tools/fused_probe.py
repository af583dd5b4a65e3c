"""Exploration: time schedules on a workload.  python tools/fused_probe.py c2 [reps] [mode,mode,...]
mode = g (Gram schedule), r (fused, phi recomputed), t (fused, trajectory, 8-warp kernel), p (trajectory, pipelined
16-warp kernel); suffix :N = fused_debug bits"""
import json, os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from quantum_geometric_tensor_b200 import api, circuits as K

name = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
modes = sys.argv[3].split(",") if len(sys.argv) > 3 else ["g", "r", "t", "p"]
extra = sys.argv[4:]
ctx = api.Context(0)
for kv in extra:
    k_, v_ = kv.split("=")
    ctx.set_option(k_, float(v_))
c = K.config(name)
th = K.default_angles(c.num_params)
res = {}
for mode in modes:
    m, _, dbg = mode.partition(":")
    ctx.set_option("fused", 0 if m == "g" else 1)
    ctx.set_option("fused_traj", 1 if m in "tpld" else 0)
    ctx.set_option("fused_pipeline", {"p": 1, "l": 2, "d": 3}.get(m, 0))
    ctx.set_option("fused_debug", float(dbg or 0))
    best = None
    for _ in range(reps):
        q = ctx.qgt(c, th)
        st = ctx.stats()
        if best is None or st["ms_total"] < best["ms_total"]:
            best = st
    if not dbg:
        res[mode] = q
    keys = ("ms_total", "ms_sweep", "ms_gram", "ms_other", "sweep_launches", "sweep_column_passes", "blocks",
            "resident_columns", "fused", "ms_wall", "ms_host_plan")
    d = {k: round(best[k], 3) if isinstance(best[k], float) else best[k] for k in keys}
    d["sweep_TBs"] = round(best["sweep_bytes"] / max(best["ms_sweep"], 1e-9) * 1e-9, 3)
    d["sweep_TFs"] = round(best["tensor_flops"] / max(best["ms_sweep"], 1e-9) * 1e-9, 2)
    print(name, mode, json.dumps(d), flush=True)
ks = list(res)
for k in ks[1:]:
    print("rel diff", ks[0], k, np.abs(res[ks[0]] - res[k]).max() / np.abs(res[ks[0]]).max())
