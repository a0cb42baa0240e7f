#!/bin/bash
# full ncu capture of one kernel (regex $1) from a --profile bench run; report gpurun_out/prof_$2.ncu-rep
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$1 -s 1 -c 1 -o gpurun_out/prof_$2 python bench.py --steps 1 --warmup 1 --profile > gpurun_out/ncu_$2.log 2>&1
tail -n 2 gpurun_out/ncu_$2.log
