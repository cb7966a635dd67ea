#!/bin/bash
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_s2_n1.json 2> gpurun_out/r2_s2_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_s2_n2.json 2> gpurun_out/r2_s2_n2.err
python - <<PY
import json
a=json.load(open("gpurun_out/r2_s2_n1.json")); b=json.load(open("gpurun_out/r2_s2_n2.json"))
print("N=1", round(a["value"]), "e2e", round(a["e2e"]["value"]), "ms", round(a["ms_per_step"],1))
print("N=2", round(b["value"]), "e2e", round(b["e2e"]["value"]), "ms", round(b["ms_per_step"],1), "eff", round(b["value"]/2/a["value"],3), "e2e eff", round(b["e2e"]["value"]/2/a["e2e"]["value"],3))
print([(r["rank"], r["dev_ms_per_step"], r["serial_kernel_ms"], r.get("grow_ms")) for r in b.get("per_rank",[])])
PY
