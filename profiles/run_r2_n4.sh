#!/bin/bash
# N = 4 record of the default bench on the final kernels
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR --nproc-per-node 4 bench.py --gpus 4 --steps 20 --warmup 6 --quick > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench4.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d.get("median_ms_per_step"), json.dumps(d.get("stage_ms_per_step")))
for k in ("list_sharded", "strong_scaling_10k", "parity", "e2e"):
    if k in d: print(" ", k, json.dumps(d[k])[:900])
PY
grep -v "^\*\|OMP" gpurun_out/r2_bench4.err | tail -n 4
