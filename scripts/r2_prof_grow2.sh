#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'^k_lsd_grow$' -s 1 -c 1 -o /tmp/r02_grow python bench.py --batch 64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_grow_prof.log 2>&1
ncu -i /tmp/r02_grow.ncu-rep --page source --csv > gpurun_out/r02_src_grow.csv 2>/dev/null
ncu -i /tmp/r02_grow.ncu-rep --page raw --csv > gpurun_out/r02_raw_grow.csv 2>/dev/null
ls -la gpurun_out/r02_src_grow.csv
