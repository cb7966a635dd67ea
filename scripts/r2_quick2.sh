#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lsd_gpu.py tests/test_lbd_gpu.py tests/test_pipeline_gpu.py tests/test_pipeline_large_gpu.py tests/test_shim.py -x -q 2>&1 | tail -3 > gpurun_out/r2_q_tests.log
cat gpurun_out/r2_q_tests.log
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2_q_bench.json 2> gpurun_out/r2_q_bench.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_q_bench.json"))
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],1), "serial", d["per_rank"][0]["serial_kernel_ms"], [(k["kernel"].split('.')[-1],round(k["ms"],2)) for k in d["kernels"] if k["ms"]>2.0])
except Exception as e:
    print("bench failed", e, open("gpurun_out/r2_q_bench.err").read()[-500:])
PY
