#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_fit_gputests.log 2>&1
grep -E "passed|failed" gpurun_out/r2_fit_gputests.log | tail -1
if grep -qE "failed|error" gpurun_out/r2_fit_gputests.log; then tail -40 gpurun_out/r2_fit_gputests.log; fi
for M in async inline async inline; do
if [ $M = inline ]; then export PLF_LSD_FIT_INLINE=1; else unset PLF_LSD_FIT_INLINE; fi
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_fit_$M.json 2> gpurun_out/r2_fit_$M.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_fit_$M.json"))
    print("fit $M: value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],1), d["pipeline_timeline_ms"]["last_two_batches"], [(k["kernel"].split('.')[-1],round(k["ms"],2)) for k in d["kernels"] if "rects" in k["kernel"]])
except Exception as e:
    print("bench failed", e, open("gpurun_out/r2_fit_$M.err").read()[-800:])
PY
done
