#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lba_gpu.py -x -q 2>&1 | tail -15 > gpurun_out/r2_lba_tests.log
cat gpurun_out/r2_lba_tests.log
# compute-sanitizer on the parity tests of every operator + a pipelined batch test (small images: the tools slow kernels ~50x)
export PLF_SANITIZE_SMALL=1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_pipeline_gpu.py::test_batches_in_flight_match_sequential tests/test_pipeline_gpu.py::test_pipeline_reset_and_too_few_features tests/test_matching_gpu.py tests/test_gn_gpu.py tests/test_lba_gpu.py tests/test_matchgrid_gpu.py tests/test_mapfeatures_gpu.py tests/test_loopclosure_gpu.py -x -q > gpurun_out/r2_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/r2_memcheck.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "tests/test_pipeline_gpu.py::test_pipeline_windowed_matching_strategy[euroc]" tests/test_orb_gpu.py tests/test_lbd_gpu.py -x -q > gpurun_out/r2_memcheck2.log 2>&1
echo "memcheck2 rc=$?"; tail -4 gpurun_out/r2_memcheck2.log
