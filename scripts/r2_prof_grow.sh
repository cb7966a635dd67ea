#!/bin/bash
mkdir -p gpurun_out
export PLF_GROW_CFG=411
ncu --set full --clock-control none --import-source on -k regex:k_lsd_grow_t -s 2 -c 1 -o gpurun_out/r2_grow_t411 python bench.py --batch 256 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_prof1.log 2>&1
tail -3 gpurun_out/r2_prof1.log
export PLF_GROW_CFG=121
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_sweep_121.json 2> gpurun_out/r2_sweep_121.err
export PLF_GROW_CFG=221
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_sweep_221.json 2> gpurun_out/r2_sweep_221.err
python - <<PY
import json
for m in ("121","221"):
    try:
        d=json.load(open(f"gpurun_out/r2_sweep_{m}.json"))
        g=[k for k in d["kernels"] if "grow" in k["kernel"]]
        print(m, "value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "grow", [(k["kernel"],k["ms"]) for k in g])
    except Exception as e: print(m,"failed",e)
PY
