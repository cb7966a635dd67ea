#!/bin/bash
# round-2 first GPU pass: full GPU test suite, default bench line, warp-vs-thread growing A/B
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_gpu.txt
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_gputests.log 2>&1
tail -5 gpurun_out/r2_gputests.log
python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
tail -c 600 gpurun_out/r2_bench_default.err
for mode in warp 411 421 811 821 1621; do
  if [ "$mode" = "warp" ]; then export PLF_GROW=warp; unset PLF_GROW_CFG; else unset PLF_GROW; export PLF_GROW_CFG=$mode; fi
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_sweep_$mode.json 2> gpurun_out/r2_sweep_$mode.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_sweep_$mode.json"))
    g=[k for k in d["kernels"] if "grow" in k["kernel"]]
    print("$mode", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "grow", [(k["kernel"],k["ms"]) for k in g], "tracked", d["config"].get("tracked_fraction"))
except Exception as e:
    print("$mode failed", e)
PY
done
