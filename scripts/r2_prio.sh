#!/bin/bash
mkdir -p gpurun_out
for P in 1 0 1 0; do
PLF_G_PRIORITY=$P timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_prio_$P.json 2> gpurun_out/r2_prio_$P.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_prio_$P.json"))
    print("G priority $P: value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],1), d["pipeline_timeline_ms"]["last_two_batches"][1])
except Exception as e:
    print("bench failed", e, open("gpurun_out/r2_prio_$P.err").read()[-500:])
PY
done
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_pipeline_large_gpu.py -x -q 2>&1 | tail -2
