#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q > gpurun_out/r2_n1e_pytest.log 2>&1
tail -3 gpurun_out/r2_n1e_pytest.log | cut -c1-300
python bench.py --steps 10 --warmup 6 --quick --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['stage_ms_per_step'], d['e2e']['value'], d['parity'])"
rm -f gpurun_out/r2_hbm_regime.jsonl
bash profiles/run_r2_hbm.sh > gpurun_out/r2_hbm.log 2>&1
cut -c1-330 gpurun_out/r2_hbm_regime.jsonl; head -8 gpurun_out/r2_hbm_regime_ncu.txt; grep -E "long_scoreboard|stalled_barrier|issue_active" gpurun_out/r2_hbm_regime_ncu.txt
