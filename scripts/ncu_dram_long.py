#!/usr/bin/env python
"""DRAM bytes per image and launch time per kernel from a long-format ncu CSV
(`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file F`).

usage: ncu_dram_long.py F.csv IMAGES_PER_LAUNCH [SKIP_LAUNCHES_PER_KERNEL]
Launches of a kernel are averaged after skipping the first SKIP (warm-up batches see cold caches / first-touch pages).
"""
import csv
import json
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
path, nimg = sys.argv[1], int(sys.argv[2])
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
col = {c: i for i, c in enumerate(rows[0])}
per = {}
for r in rows[1:]:
    name = r[col["Kernel Name"]].split("(")[0].split("<")[0].replace("void ", "")
    d = per.setdefault((name, r[col["ID"]]), {})
    d[r[col["Metric Name"]]] = float(r[col["Metric Value"]].replace(",", "")) * UNIT.get(r[col["Metric Unit"]], 1.0)
agg = {}
for (name, _id), d in per.items():
    agg.setdefault(name, []).append(d)
out = {}
for name, ls in agg.items():
    ls = ls[skip:] if len(ls) > skip else ls
    n = len(ls)
    rd = sum(d.get("dram__bytes_read.sum", 0.0) for d in ls) / n
    wr = sum(d.get("dram__bytes_write.sum", 0.0) for d in ls) / n
    ms = sum(d.get("gpu__time_duration.sum", 0.0) for d in ls) / n
    out[name] = dict(launches=n, dram_read_per_image=round(rd / nimg), dram_write_per_image=round(wr / nimg),
                     dram_bytes_per_image=round((rd + wr) / nimg), ms_per_launch=round(ms, 3))
print(json.dumps(out, indent=1))
