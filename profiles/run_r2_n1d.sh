#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q > gpurun_out/r2_n1d_pytest.log 2>&1
tail -4 gpurun_out/r2_n1d_pytest.log | cut -c1-300
for i in 1 2 3; do timeout 300 python -X faulthandler -m pytest tests/test_gpu_parity.py -m gpu -x -q -k concurrent 2>&1 | tail -2 | cut -c1-200; done
for dep in 1 4; do for nq in 1 8; do MMIDX_STATS=1 MMIDX_SCAN_DEPTH=$dep python profiles/hbm_regime.py --nq $nq --reps 10 2>&1 | cut -c1-330; done; done
python bench.py --steps 10 --warmup 6 --quick --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['stage_ms_per_step'], d['e2e']['value'], d['parity'])"
