#!/bin/bash
# round 1 record pass: gpu tests, reference arm, bench with CPU baseline, launch list, full capture of the scan kernel
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu.txt
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu.log 2>&1
tail -n 5 gpurun_out/pytest_gpu.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_1.json 2> gpurun_out/bench_1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --profile > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ivfpq_scan_fast -s 1 -c 1 -o gpurun_out/prof_scan_final python bench.py --steps 1 --warmup 1 --profile > gpurun_out/ncu_full.log 2>&1
tail -n 3 gpurun_out/bench_*.json
