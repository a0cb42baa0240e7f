#!/bin/bash
# full ncu captures (with source) of the coarse verify / prep / filter kernels of a search step (the index build's launches are skipped by name / count)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_coarse_verify|k_fast_prep|k_fast_t2' -c 3 -o gpurun_out/prof_r2_cv python bench.py --steps 1 --warmup 1 --profile > gpurun_out/ncu_r2_cv.log 2>&1
tail -n 2 gpurun_out/ncu_r2_cv.log
ncu --set full --clock-control none --import-source on -k regex:'k_coarse_mma' -s 17 -c 1 -o gpurun_out/prof_r2_mma python bench.py --steps 1 --warmup 1 --profile > gpurun_out/ncu_r2_mma.log 2>&1
tail -n 2 gpurun_out/ncu_r2_mma.log
