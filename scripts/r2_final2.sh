#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_final_gputests.log 2>&1
tail -3 gpurun_out/r2_final_gputests.log
python __graft_entry__.py smoke > gpurun_out/r2_final_smoke.log 2>&1; tail -1 gpurun_out/r2_final_smoke.log
timeout 900 python bench.py > gpurun_out/r2_final_kitti.json 2> gpurun_out/r2_final_kitti.err
timeout 900 python bench.py --config euroc > gpurun_out/r2_final_euroc.json 2> gpurun_out/r2_final_euroc.err
timeout 900 python bench.py --config lowtex > gpurun_out/r2_final_lowtex.json 2> gpurun_out/r2_final_lowtex.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_reference.json 2> gpurun_out/r2_final_reference.err
python - <<PY
import json
for c in ("kitti","euroc","lowtex","reference"):
    try:
        d=json.load(open(f"gpurun_out/r2_final_{c}.json"))
        print(c, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],1), "cpu", d.get("cpu_baseline",{}).get("value"), "cores", d.get("cpu_baseline",{}).get("cores"))
    except Exception as e:
        print(c, "failed", e)
PY
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_orb_gpu.py tests/test_lbd_gpu.py::test_gradients_golden tests/test_matching_gpu.py tests/test_gn_gpu.py -x -q > gpurun_out/r2_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -5 gpurun_out/r2_racecheck.log
