#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_par_gputests.log 2>&1
tail -4 gpurun_out/r2_par_gputests.log
for P in 2 1; do
  export PLF_LSD_PARITIES=$P
  timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2_par_$P.json 2> gpurun_out/r2_par_$P.err
  tail -2 gpurun_out/r2_par_$P.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_par_$P.json"))
    print("parities=$P", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],1), "timeline", d.get("pipeline_timeline_ms",{}).get("last_two_batches"))
except Exception as e:
    print("par $P failed", e)
PY
done
