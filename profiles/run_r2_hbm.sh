#!/bin/bash
# HBM-regime roofline of the fused scan: events-timed runs at nq = 1, 2, 8 and one ncu --set full capture at nq = 1
mkdir -p gpurun_out
N=${N:-134217728}
for nq in 1 2 8; do python profiles/hbm_regime.py --n $N --nq $nq >> gpurun_out/r2_hbm_regime.jsonl 2>> gpurun_out/r2_hbm_regime.err; done
ncu --set full --clock-control none --import-source on -k regex:k_ivfpq_scan_fast -s 2 -c 1 -o gpurun_out/prof_r2_hbm python profiles/hbm_regime.py --n $N --nq 1 --reps 3 > gpurun_out/ncu_r2_hbm.log 2>&1
python profiles/ncu_summary.py gpurun_out/prof_r2_hbm.ncu-rep > gpurun_out/r2_hbm_regime_ncu.txt 2>&1
cat gpurun_out/r2_hbm_regime.jsonl; head -30 gpurun_out/r2_hbm_regime_ncu.txt; tail -3 gpurun_out/r2_hbm_regime.err
