#!/bin/bash
mkdir -p gpurun_out
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/r02_inst_b64.csv python bench.py --batch 64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_inst_b64_bench.log 2>&1
tail -2 gpurun_out/r02_inst_b64.csv | cut -c1-200
