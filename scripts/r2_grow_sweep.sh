python -m pytest tests/test_lsd_gpu.py tests/test_pipeline_gpu.py -x -q 2>&1 | tail -5 > gpurun_out/r2_test2.log
cat gpurun_out/r2_test2.log
for cfg in 411 421 433 811 821 1621 3221; do
  if [ "$cfg" = "warp" ]; then export PLF_GROW=warp; else unset PLF_GROW; export PLF_GROW_CFG=$cfg; fi
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_sweep_$cfg.json 2> gpurun_out/r2_sweep_$cfg.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_sweep_$cfg.json"))
g=[k for k in d["kernels"] if "grow" in k["kernel"]][0]
print("$cfg", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "grow ms", g["ms"], "tracked", d["config"]["tracked_fraction"])
PY
done
