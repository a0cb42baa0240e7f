#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q > gpurun_out/r2_n1f_pytest.log 2>&1
tail -15 gpurun_out/r2_n1f_pytest.log | cut -c1-300
python bench.py --steps 10 --warmup 6 --no-cpu-baseline 2> gpurun_out/r2_n1f_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['stage_ms_per_step'], d['e2e']['value'], d['parity'])
print(json.dumps(d['rows'])[:3000])"
tail -3 gpurun_out/r2_n1f_bench.err
