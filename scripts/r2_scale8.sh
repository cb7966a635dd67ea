#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python bench.py --gpus 1 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2_scale_n1.json 2> gpurun_out/r2_scale_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2_scale_n8.json 2> gpurun_out/r2_scale_n8.err
python - <<PY
import json
a=json.load(open("gpurun_out/r2_scale_n1.json")); b=json.load(open("gpurun_out/r2_scale_n8.json"))
print("N=1", round(a["value"]), "e2e", round(a["e2e"]["value"]), "ms", round(a["ms_per_step"],1))
print("N=8", round(b["value"]), "e2e", round(b["e2e"]["value"]), "ms", round(b["ms_per_step"],1), "eff", round(b["value"]/8/a["value"],3))
print([ (r["rank"], r["dev_ms_per_step"], r.get("grow_ms")) for r in b.get("per_rank",[])])
PY
