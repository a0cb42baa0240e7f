#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_n2b_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
MMIDX_STATS=1 MMIDX_VERBOSE=1 timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 6 --config 4 --n-db ${NDB:-2000000} > gpurun_out/r2_n2b_cfg4.json 2> gpurun_out/r2_n2b_cfg4.err
MMIDX_VERBOSE=1 timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 6 --no-cpu-baseline --quick > gpurun_out/r2_n2b_bench.json 2> gpurun_out/r2_n2b_bench.err
cat gpurun_out/r2_n2b_pytest.log
python - <<'PY'
import json
for f in ("gpurun_out/r2_n2b_bench.json", "gpurun_out/r2_n2b_cfg4.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d.get("stage_ms_per_step"))
        for k in ("list_sharded", "replicas", "strong_scaling_10k", "parity", "e2e"):
            if k in d: print(" ", k, json.dumps(d[k])[:1200])
    except Exception as e:
        print(f, "ERR", e)
PY
grep -v "^\*\|OMP" gpurun_out/r2_n2b_bench.err | tail -8
grep -v "^\*\|OMP" gpurun_out/r2_n2b_cfg4.err | tail -25
