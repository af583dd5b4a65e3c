"""Exploration: time the fused and the Gram schedule on a workload.  python tools/fused_probe.py c2 [reps]"""
import json, os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from quantum_geometric_tensor_b200 import api, circuits as K

name = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
modes = sys.argv[3].split(",") if len(sys.argv) > 3 else ["0", "1"]
ctx = api.Context(0)
c = K.config(name)
th = K.default_angles(c.num_params)
res = {}
for mode in modes:
    ctx.set_option("fused", float(mode))
    best = None
    for _ in range(reps):
        q = ctx.qgt(c, th)
        st = ctx.stats()
        if best is None or st["ms_total"] < best["ms_total"]:
            best = st
    res[mode] = q
    keys = ("ms_total", "ms_sweep", "ms_gram", "ms_other", "sweep_bytes", "tensor_flops", "sweep_launches", "sweep_column_passes", "blocks",
            "resident_columns", "fused", "ms_wall", "ms_host_plan")
    d = {k: best[k] for k in keys}
    d["sweep_TBs"] = best["sweep_bytes"] / max(best["ms_sweep"], 1e-9) * 1e-9
    d["sweep_TFs"] = best["tensor_flops"] / max(best["ms_sweep"], 1e-9) * 1e-9
    print(name, "fused=" + mode, json.dumps(d), flush=True)
if len(res) == 2:
    a, b = res[modes[0]], res[modes[1]]
    print("rel diff between schedules", np.abs(a - b).max() / np.abs(a).max())
