#!/bin/bash
# 1-GPU pass: GPU tests, HBM-regime scan with the deep sweep, config 4 (10M) on one GPU with both sweep forms
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_n1b_pytest.log
rm -f gpurun_out/r2_hbm_regime.jsonl
N=${N:-134217728} bash profiles/run_r2_hbm.sh > gpurun_out/r2_hbm.log 2>&1
MMIDX_SCAN_DEPTH=1 python profiles/hbm_regime.py --nq 1 >> gpurun_out/r2_hbm_regime_depth1.jsonl 2>> gpurun_out/r2_hbm_regime.err
for dep in 4 1; do
  MMIDX_SCAN_DEPTH=$dep timeout 900 python bench.py --config 4 --steps 5 --warmup 6 > gpurun_out/r2_cfg4_n1_depth$dep.json 2> gpurun_out/r2_cfg4_n1_depth$dep.err
done
cat gpurun_out/r2_n1b_pytest.log
cut -c1-420 gpurun_out/r2_hbm_regime.jsonl; cut -c1-420 gpurun_out/r2_hbm_regime_depth1.jsonl
head -6 gpurun_out/r2_hbm_regime_ncu.txt; grep -E "long_scoreboard|registers_per_thread|issue_active" gpurun_out/r2_hbm_regime_ncu.txt
python - <<'PY'
import json
for dep in (4, 1):
    try:
        d = json.loads(open(f"gpurun_out/r2_cfg4_n1_depth{dep}.json").read().strip().splitlines()[-1])
        print("cfg4 depth", dep, d["value"], d["ms_per_step"], json.dumps(d["stage_ms_per_step"]), json.dumps(d["roofline"])[:400], json.dumps(d["parity"]), json.dumps(d["e2e"])[:200])
    except Exception as e:
        print("cfg4", dep, "ERR", e)
PY
tail -5 gpurun_out/r2_cfg4_n1_depth4.err
