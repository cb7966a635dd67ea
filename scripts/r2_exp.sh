#!/bin/bash
mkdir -p gpurun_out
run() { # name, env..., args
  name=$1; shift
  timeout 600 env "$@" python bench.py --steps 5 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/r2_exp_$name.json 2> gpurun_out/r2_exp_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_exp_$name.json"))
    g=[k for k in d["kernels"] if "grow" in k["kernel"]][0]
    print("$name", "B", d["config"]["pairs_per_step_per_gpu"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],1), "grow", g["ms"], "serial", d["per_rank"][0]["serial_kernel_ms"])
except Exception as e:
    print("$name failed", e, open("gpurun_out/r2_exp_$name.err").read()[-300:])
PY
}
EXTRA=""
run base PLF_X=0
run r48 PLF_GROW_CFG=48
run r40 PLF_GROW_CFG=40
EXTRA="--batch 2048"; run b2048 PLF_X=0
EXTRA="--batch 2560"; run b2560 PLF_X=0
EXTRA="--batch 2560"; run b2560_r40 PLF_GROW_CFG=40
