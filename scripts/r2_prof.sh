#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1
# full-set capture of one launch of each hot kernel (batch 64 = 128 images); the raw page is exported here, the report stays on the box
ncu --set full --clock-control none --import-source on -k regex:'k_fast_nms|k_orb_blur7_fast|k_lsd_grad|k_rbrief|k_lsd_rects|k_blur5_sobel_fast|k_blur_q8_fast|k_lsd_rowhist|k_lsd_scatter|k_lbd|k_hamming_knn2_mma|k_resize_exact4|k_ic_angle|k_lsd_grow|k_select_sort|k_lsd_binscan' -s 70 -c 36 -o /tmp/r02_full python bench.py --batch 64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_full_bench.log 2>&1
ncu -i /tmp/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_full_raw.csv 2>/dev/null
ncu -i /tmp/r02_full.ncu-rep --page source --csv -k regex:k_fast_nms > gpurun_out/r02_src_fast.csv 2>/dev/null
ncu -i /tmp/r02_full.ncu-rep --page source --csv -k regex:k_lsd_grad > gpurun_out/r02_src_grad.csv 2>/dev/null
ls -la /tmp/r02_full.ncu-rep gpurun_out/ | tail -12
