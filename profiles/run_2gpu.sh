#!/bin/bash
# 2-GPU validation: sharded tests + scaling lines at N=1,2
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu ) > gpurun_out/pytest_2gpu.log 2>&1
tail -n 5 gpurun_out/pytest_2gpu.log
bash profiles/run_scaling.sh 1 > gpurun_out/scale_1.json
bash profiles/run_scaling.sh 2 > gpurun_out/scale_2.json
MMIDX_LIST_SHARDS=2 bash profiles/run_scaling.sh 2 > gpurun_out/scale_2_s2.json
for f in gpurun_out/scale_1.json gpurun_out/scale_2.json gpurun_out/scale_2_s2.json; do python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1]); print("$f", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["config"]["sharding"], d["parity"])
except Exception as e: print("$f failed", e)
PY
done
