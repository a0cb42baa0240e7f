#!/bin/bash
mkdir -p gpurun_out
for mode in fast exact; do
  MMIDX_MODE=$mode timeout 600 python bench.py --config 4 --n-db 3000000 --steps 3 --warmup 6 --no-cpu-baseline 2>&1 >/dev/null | grep "indexed" | cut -c1-200
done
ncu --set full --clock-control none --import-source on -k regex:k_ivfpq_scan_fast -s 3 -c 1 -o gpurun_out/prof_r2_cfg4_scan python bench.py --config 4 --steps 1 --warmup 3 --profile > gpurun_out/ncu_r2_cfg4.log 2>&1
python profiles/ncu_summary.py gpurun_out/prof_r2_cfg4_scan.ncu-rep > gpurun_out/r2_cfg4_scan_ncu.txt 2>&1
head -24 gpurun_out/r2_cfg4_scan_ncu.txt
bash profiles/run_sanitizer.sh
