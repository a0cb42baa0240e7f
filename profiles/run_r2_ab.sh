#!/bin/bash
# quick A/B of a scan-kernel change: parity subset, configs[2] and config 4 step times on one GPU
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 6 --quick --no-cpu-baseline > gpurun_out/ab_cfg3.json 2> gpurun_out/ab_cfg3.err
timeout 900 python bench.py --config 4 --steps 10 --warmup 6 --no-cpu-baseline > gpurun_out/ab_cfg4.json 2> gpurun_out/ab_cfg4.err
python - <<'PY'
import json
for f in ("gpurun_out/ab_cfg3.json", "gpurun_out/ab_cfg4.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d.get("ms_per_step"), json.dumps(d.get("stage_ms_per_step")), json.dumps(d.get("parity")), d["roofline"]["frac"])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -n 3 gpurun_out/ab_cfg3.err; tail -n 3 gpurun_out/ab_cfg4.err
