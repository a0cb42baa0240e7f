#!/bin/bash
# 2-GPU check of the multi-GPU step on the current kernels: sharded tests + quick N = 2 bench
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/r2_n2d_pytest.log 2>&1
tail -n 3 gpurun_out/r2_n2d_pytest.log | cut -c1-400
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 6 --no-cpu-baseline > gpurun_out/r2_n2d_bench.json 2> gpurun_out/r2_n2d_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_n2d_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["stage_ms_per_step"], d["launches_per_step"])
for k in ("strong_scaling_10k", "list_sharded", "parity", "e2e"):
    print(k, json.dumps(d[k])[:700])
PY
grep -v "^\*\|OMP" gpurun_out/r2_n2d_bench.err | tail -n 4
