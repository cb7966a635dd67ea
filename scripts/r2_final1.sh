#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_orb_gpu.py tests/test_lsd_gpu.py tests/test_pipeline_gpu.py tests/test_pipeline_large_gpu.py -x -q 2>&1 | tail -3 > gpurun_out/r2_f1_tests.log
cat gpurun_out/r2_f1_tests.log
timeout 900 python bench.py > gpurun_out/r2_f1_kitti.json 2> gpurun_out/r2_f1_kitti.err
timeout 900 python bench.py --config euroc > gpurun_out/r2_f1_euroc.json 2> gpurun_out/r2_f1_euroc.err
timeout 900 python bench.py --config lowtex > gpurun_out/r2_f1_lowtex.json 2> gpurun_out/r2_f1_lowtex.err
python - <<PY
import json
for c in ("kitti","euroc","lowtex"):
    try:
        d=json.load(open(f"gpurun_out/r2_f1_{c}.json"))
        print(c, "B", d["config"]["pairs_per_step_per_gpu"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],1), "cpu", d.get("cpu_baseline",{}).get("value"), [(k["kernel"].split('.')[-1],round(k["ms"],1)) for k in d["kernels"] if k["ms"]>3.0])
    except Exception as e:
        print(c, "failed", e, open(f"gpurun_out/r2_f1_{c}.err").read()[-400:])
PY
