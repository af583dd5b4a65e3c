#!/usr/bin/env python3
"""profiles/rNN_dram_c2.csv (ncu dram__bytes_* per launch) -> profiles/traffic.json (what bench.py reports as roofline.traffic).

The csv holds the first sweep / Gram launches of `bench.py --steps 1 --warmup 1` = a few identical evaluations, each
ending with its Gram launch; the last complete evaluation is used.  Per-launch averages, like `achieved`.
"""
import csv
import json
import sys

src, dst = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
per = {}
for r in rows:
    lid, name, metric, unit, val = int(r[0]), r[4], r[-3], r[-2], float(r[-1].replace(",", ""))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    d = per.setdefault(lid, {"name": name, "bytes": 0.0})
    if metric.startswith("dram__bytes"):
        d["bytes"] += val * scale
launches = [per[k] for k in sorted(per)]
ends = [i for i, l in enumerate(launches) if "gram" in l["name"]]          # an evaluation ends with its Gram
assert len(ends) >= 2, "need at least two complete evaluations"
last = launches[ends[-2] + 1: ends[-1] + 1]
sweeps = [l["bytes"] for l in last if "sweep" in l["name"]]
grams = [l["bytes"] for l in last if "gram" in l["name"]]
out = {"c2": {"sweep_bytes_per_launch": sum(sweeps) / max(1, len(sweeps)), "gram_bytes_per_launch": sum(grams) / max(1, len(grams)),
              "sweep_launches": len(sweeps), "gram_launches": len(grams),
              "note": f"dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the {len(sweeps)} sweep launches / "
                      f"{len(grams)} Gram launch of one C2 evaluation ({src}); algorithmic bytes per sweep launch are in the bench line"}}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out))
