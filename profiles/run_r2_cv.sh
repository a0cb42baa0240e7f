#!/bin/bash
# whole GPU suite on the current build, then full ncu captures (with source) of the coarse verify / prep kernels and the scan
mkdir -p gpurun_out
timeout 1200 python -X faulthandler -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_b.log 2>&1
tail -n 3 gpurun_out/r2_pytest_b.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:'k_coarse_verify|k_fast_prep|k_coarse_mma' -s 3 -c 3 -o gpurun_out/prof_r2_cv python bench.py --steps 1 --warmup 1 --profile > gpurun_out/ncu_r2_cv.log 2>&1
tail -n 2 gpurun_out/ncu_r2_cv.log
ncu --set full --clock-control none --import-source on -k regex:k_ivfpq_scan_fast -s 1 -c 1 -o gpurun_out/prof_r2_scan2 python bench.py --steps 1 --warmup 1 --profile > gpurun_out/ncu_r2_scan2.log 2>&1
tail -n 2 gpurun_out/ncu_r2_scan2.log
