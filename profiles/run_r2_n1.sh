#!/bin/bash
# 1-GPU record pass of round 2: HBM-regime scan, default bench, launch list, full ncu capture of the scan kernel
mkdir -p gpurun_out
rm -f gpurun_out/r2_hbm_regime.jsonl
N=${N:-134217728} bash profiles/run_r2_hbm.sh > gpurun_out/r2_hbm.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_ref1.json 2> gpurun_out/r2_ref1.err
MMIDX_VERBOSE=1 python bench.py --steps 20 --warmup 6 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --profile > gpurun_out/r2_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ivfpq_scan_fast -s 3 -c 1 -o gpurun_out/prof_r2_scan python bench.py --steps 1 --warmup 3 --profile > gpurun_out/ncu_r2_scan.log 2>&1
python profiles/ncu_summary.py gpurun_out/prof_r2_scan.ncu-rep > gpurun_out/r2_scan_ncu.txt 2>&1
tail -5 gpurun_out/r2_hbm.log | cut -c1-600
cat gpurun_out/r2_hbm_regime.jsonl | cut -c1-700
head -12 gpurun_out/r2_hbm_regime_ncu.txt
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench1.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "median_ms_per_step", "stage_ms_per_step", "roofline", "e2e", "parity", "ties", "cpu_baseline"):
    print(k, json.dumps(d.get(k))[:600])
print(json.dumps(d.get("small_batch"))[:1500])
PY
tail -4 gpurun_out/r2_bench1.err
head -8 gpurun_out/r2_scan_ncu.txt
tail -12 gpurun_out/r2_launches.csv | cut -d, -f5,9,15 | cut -c1-160
