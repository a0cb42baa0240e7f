#!/bin/bash
# round 1, first measurement pass: bench at several LUT chunk sizes, launch list, one full capture of the scan kernel
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c64.json 2> gpurun_out/bench_c64.err
for mb in 256 1024 4096; do
  MMIDX_LUT_CHUNK_MB=$mb python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c$mb.json 2> gpurun_out/bench_c$mb.err
done
MMIDX_LUT_CHUNK_MB=1024 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --profile > gpurun_out/ncu_launch.log 2>&1
MMIDX_LUT_CHUNK_MB=1024 ncu --set full --clock-control none --import-source on -k regex:k_ivfpq_scan -s 5 -c 2 -o gpurun_out/prof_scan_r1 python bench.py --steps 1 --warmup 1 --profile > gpurun_out/ncu_full.log 2>&1
tail -n 3 gpurun_out/bench_*.json
