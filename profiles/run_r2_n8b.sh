#!/bin/bash
# 8-GPU record pass at the end of round 2 (kernels of the final commit): default bench at N = 8, config 4 (10M) over 8 GPUs
mkdir -p gpurun_out
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR --nproc-per-node 8 bench.py --gpus 8 --steps 20 --warmup 6 > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err
timeout 1200 $TR --nproc-per-node 8 bench.py --config 4 --gpus 8 --steps 10 --warmup 6 > gpurun_out/r2_cfg4_n8.json 2> gpurun_out/r2_cfg4_n8.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench8.json", "gpurun_out/r2_cfg4_n8.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d.get("median_ms_per_step"), json.dumps(d.get("stage_ms_per_step")))
        for k in ("list_sharded", "replicas", "strong_scaling_10k", "parity", "e2e", "roofline", "ties"):
            if k in d: print(" ", k, json.dumps(d[k])[:1300])
    except Exception as e:
        print(f, "ERR", e)
PY
for f in r2_bench8 r2_cfg4_n8; do grep -v "^\*\|OMP" gpurun_out/$f.err | tail -n 12; done
