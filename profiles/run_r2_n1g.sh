#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q > gpurun_out/r2_n1g_pytest.log 2>&1
tail -12 gpurun_out/r2_n1g_pytest.log | cut -c1-300
python bench.py --steps 10 --warmup 6 --no-cpu-baseline 2> gpurun_out/r2_n1g_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['stage_ms_per_step'], d['e2e']['value'], d['parity'])
print(json.dumps(d['rows'])[:3000])"
timeout 900 python bench.py --config 4 --steps 5 --warmup 6 2> gpurun_out/r2_n1g_cfg4.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg4', d['value'], d['ms_per_step'], d['stage_ms_per_step'], d['roofline']['frac'], d['parity'])"
tail -3 gpurun_out/r2_n1g_cfg4.err
