#!/bin/bash
mkdir -p gpurun_out
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_orb_gpu.py tests/test_lsd_gpu.py -x -q > gpurun_out/r2_memcheck_final.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/r2_memcheck_final.log
timeout 150 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_orb_gpu.py -x -q > gpurun_out/r2_racecheck_final.log 2>&1
echo "racecheck rc=$?"; tail -3 gpurun_out/r2_racecheck_final.log
