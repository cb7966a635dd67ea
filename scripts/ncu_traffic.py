#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: DRAM bytes, duration, issue utilisation and occupancy per kernel launch.

usage: ncu_traffic.py RAW.csv IMAGES_PER_LAUNCH [--update profiles/ncu_traffic_per_image.json]

Prints one line per kernel (launches of the same kernel averaged) and, with --update, rewrites the per-image DRAM byte
counts `bench.py` reports next to the algorithmic ones (only the kernels present in the capture are touched).
"""
import csv
import json
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "%": 1.0, "": 1.0}

# kernel name in the capture -> key of bench.py's `kernels` table (stage.kernel); resize / hamming run in two stages and are
# told apart by their grid, so they are left to the capture that isolates them
KEYS = {
    "k_fast_nms": "orb.k_fast_nms", "k_orb_blur7_fast": "orb.k_orb_blur7", "k_ic_angle": "orb.k_ic_angle",
    "k_rbrief": "orb.k_rbrief", "k_select_sort": "orb.k_select_sort", "k_blur_q8_fast": "lsd.k_blur_q8",
    "k_lsd_grad": "lsd.k_lsd_grad", "k_lsd_rowhist": "lsd.k_lsd_rowhist", "k_lsd_binscan": "lsd.k_lsd_binscan",
    "k_lsd_scatter": "lsd.k_lsd_scatter", "k_lsd_grow": "lsd.k_lsd_grow", "k_lsd_rects": "lsd.k_lsd_rects",
    "k_blur5_sobel_fast": "lbd.k_blur5_sobel", "k_lbd": "lbd.k_lbd",
}


def main():
    path, nimg = sys.argv[1], int(sys.argv[2])
    upd = sys.argv[sys.argv.index("--update") + 1] if "--update" in sys.argv else None
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {c: i for i, c in enumerate(hdr)}

    def val(r, name):
        i = col[name]
        return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)

    agg = {}
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = r[col["Kernel Name"]].split("(")[0].split("<")[0]
        a = agg.setdefault(name, dict(n=0, dram=0.0, ms=0.0, issue=0.0, occ=0.0, inst=0.0))
        a["n"] += 1
        a["dram"] += val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
        a["ms"] += val(r, "gpu__time_duration.sum")
        a["issue"] += val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")
        a["occ"] += val(r, "sm__warps_active.avg.pct_of_peak_sustained_active")
        a["inst"] += val(r, "sm__inst_executed.sum") if "sm__inst_executed.sum" in col else 0.0
    out = {}
    print(f"{'kernel':24s} {'launches':>8s} {'dram B/image':>13s} {'ms/launch':>10s} {'issue %':>8s} {'occ %':>7s} {'warp inst/image':>16s}")
    for k, a in agg.items():
        n = a["n"]
        per_img = a["dram"] / n / nimg
        print(f"{k:24s} {n:8d} {per_img:13.0f} {a['ms'] / n:10.3f} {a['issue'] / n:8.1f} {a['occ'] / n:7.1f} {a['inst'] / n / nimg:16.0f}")
        if k in KEYS:
            out[KEYS[k]] = int(round(per_img))
    if upd:
        d = json.load(open(upd))
        d.update(out)
        json.dump(d, open(upd, "w"), indent=1)
        print("updated", upd, sorted(out))


if __name__ == "__main__":
    main()
