#!/bin/bash
# kernel-variant sweep: every scratch/variants/libmmidx_*.so through bench.py --profile (device-timed, no parity legs)
mkdir -p gpurun_out
for f in scratch/variants/libmmidx_*.so; do
  n=$(basename $f .so)
  MMIDX_LIB_PATH=$PWD/$f python bench.py --steps 10 --warmup 3 --profile 2> gpurun_out/sweep_$n.err | tail -1 > gpurun_out/sweep_$n.json
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/sweep_$n.json").read())
    print("$n", "qps", round(d["value"]), "stage_ms", [round(x,3) for x in d["stage_ms_per_step"]])
except Exception as e:
    print("$n failed", e)
PY
done
