#!/bin/bash
mkdir -p gpurun_out
build/tma_probe | tail -12
# 1. TMA-staged tile kernels + padded pitches with the default (warp) growing kernel: whole GPU suite + bench
unset PLF_GROW_CFG
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_tma_gputests.log 2>&1
tail -4 gpurun_out/r2_tma_gputests.log
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_tma_bench.json 2> gpurun_out/r2_tma_bench.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_tma_bench.json"))
    print("TMA+warp", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],1), [(k["kernel"],round(k["ms"],2)) for k in d["kernels"] if k["ms"]>1.0])
except Exception as e:
    print("tma bench failed", e)
PY
# 2. lane-per-image growing kernel variants
for L in 4 2 1; do
  export PLF_GROW_CFG=$L
  timeout 600 python -m pytest tests/test_lsd_gpu.py tests/test_pipeline_gpu.py tests/test_lbd_gpu.py -x -q 2>&1 | tail -3 > gpurun_out/r2_grow_s_test_$L.log
  cat gpurun_out/r2_grow_s_test_$L.log
  timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_grow_s_$L.json 2> gpurun_out/r2_grow_s_$L.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_grow_s_$L.json"))
    g=[k for k in d["kernels"] if "grow" in k["kernel"]]
    print("L=$L", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "grow", [(k["kernel"],k["ms"]) for k in g], "tracked", d["config"].get("tracked_fraction"))
except Exception as e:
    print("L=$L failed", e)
PY
done
export PLF_GROW_CFG=2
ncu --set full --clock-control none --import-source on -k regex:k_lsd_grow_s -s 2 -c 1 -o gpurun_out/r2_grow_s2 python bench.py --batch 256 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_prof2.log 2>&1
tail -2 gpurun_out/r2_prof2.log | cut -c1-300
