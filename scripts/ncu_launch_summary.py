#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total time and share per kernel.

usage: ncu_launch_summary.py LAUNCHES.csv > SUMMARY.csv
"""
import csv
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.reader(lines))
col = {c: i for i, c in enumerate(rows[0])}
agg = {}
for r in rows[1:]:
    if r[col["Metric Name"]] != "gpu__time_duration.sum":
        continue
    unit = r[col["Metric Unit"]]
    ms = float(r[col["Metric Value"]].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[unit]
    name = r[col["Kernel Name"]].split("(")[0].split("<")[0]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
total = sum(a[1] for a in agg.values())
print("kernel,launches,total_ms,share")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k},{a[0]},{a[1]:.3f},{a[1] / total:.4f}")
