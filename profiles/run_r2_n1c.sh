#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q > gpurun_out/r2_n1c_pytest.log 2>&1
head -60 gpurun_out/r2_n1c_pytest.log | cut -c1-400
for nq in 1 8; do python profiles/hbm_regime.py --nq $nq --reps 10 2>&1 | cut -c1-330; done
MMIDX_SCAN_DEPTH=1 python profiles/hbm_regime.py --nq 8 --reps 10 2>&1 | cut -c1-330
