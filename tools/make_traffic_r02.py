#!/usr/bin/env python3
"""profiles/r02_dram_c3.csv (ncu dram__bytes_* of the qgt_fused_direct_kernel launches of one C3 evaluation inside
`bench.py --steps 1 --warmup 1`) -> the "c3_fused" entry of profiles/traffic.json (what bench.py reports as roofline.traffic).
Per-launch average over the evaluation's launches (phi launches and column launches alike), like `achieved`.

    python tools/make_traffic_r02.py profiles/r02_dram_c3.csv profiles/traffic.json
"""
import csv
import json
import os
import sys

src, dst = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
per = {}
for r in rows:
    lid, metric, unit, val = int(r[0]), r[-3], r[-2], float(r[-1].replace(",", ""))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1.0)
    d = per.setdefault(lid, {"bytes": 0.0, "ms": 0.0})
    if metric.startswith("dram__bytes"):
        d["bytes"] += val * scale
    elif metric.startswith("gpu__time"):
        d["ms"] += val * scale
L = [per[k] for k in sorted(per)]
cols = [l for l in L if l["ms"] > 20.0]          # column launches (hundreds of ms) vs phi launches (a few ms)
out = json.load(open(dst)) if os.path.exists(dst) else {}
out["c3_fused"] = {
    "sweep_bytes_per_launch": sum(l["bytes"] for l in L) / len(L),
    "launches": len(L), "column_launches": len(cols),
    "column_launch_bytes": sum(l["bytes"] for l in cols) / max(1, len(cols)),
    "column_launch_ms_under_ncu": sum(l["ms"] for l in cols) / max(1, len(cols)),
    "dram_gbs_over_all_launches": sum(l["bytes"] for l in L) / sum(l["ms"] for l in L) * 1e-6,
    "note": f"dram__bytes_read.sum + dram__bytes_write.sum per qgt_fused_direct_kernel launch, averaged over the {len(L)} launches of one "
            f"C3 evaluation ({src}); algorithmic bytes and flops per launch are in the bench line"}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out["c3_fused"]))
