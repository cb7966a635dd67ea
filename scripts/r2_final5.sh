#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_final_gputests.log 2>&1
grep -E "passed|failed" gpurun_out/r2_final_gputests.log | tail -1
if grep -qE "failed|error" gpurun_out/r2_final_gputests.log; then tail -30 gpurun_out/r2_final_gputests.log; echo "GPU tests failed - stopping"; exit 1; fi
python __graft_entry__.py smoke > gpurun_out/r2_final_smoke.log 2>&1; tail -1 gpurun_out/r2_final_smoke.log
ncu --set full --clock-control none -k regex:'k_lsd_grad|k_lsd_rowhist|k_lsd_binscan|k_lsd_scatter|k_lsd_rects|k_fast_nms|k_lsd_grow' -s 20 -c 14 -o /tmp/r02b python bench.py --batch 64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02b_bench.log 2>&1
ncu -i /tmp/r02b.ncu-rep --page raw --csv > gpurun_out/r02b_full_raw.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02b_launches_bench.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_lsd_|k_fast_nms|k_resize|k_rbrief|k_orb_blur7|k_lbd|k_blur' -c 200 --csv --log-file gpurun_out/r02_dram_b1536.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_dram_b1536_bench.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_kitti.json 2> gpurun_out/r2_final_kitti.err
timeout 900 python bench.py --config euroc --steps 20 --warmup 5 > gpurun_out/r2_final_euroc.json 2> gpurun_out/r2_final_euroc.err
timeout 900 python bench.py --config lowtex --steps 20 --warmup 5 > gpurun_out/r2_final_lowtex.json 2> gpurun_out/r2_final_lowtex.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_reference.json 2> gpurun_out/r2_final_reference.err
python - <<PY
import json
for c in ("kitti","euroc","lowtex","reference"):
    try:
        d=json.load(open(f"gpurun_out/r2_final_{c}.json"))
        print(c, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],1), "cpu", d.get("cpu_baseline",{}).get("value"))
    except Exception as e:
        print(c, "failed", e)
PY
