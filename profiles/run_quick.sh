#!/bin/bash
# quick iteration pass: gpu tests, one bench line (no CPU baseline), optional full capture of the fused scan kernel
# usage: run_quick.sh [tag] [ncu]
TAG=${1:-q}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_$TAG.log 2>&1
tail -n 4 gpurun_out/pytest_$TAG.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
    print("qps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "stages", {k: round(v,3) for k,v in d["stage_ms_per_step"].items()}, "frac", round(d["roofline"]["frac"],3), "parity", d["parity"], "recall", d["recall_at_100"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_$TAG.err").read()[-2000:])
PY
if [ "$2" = "ncu" ]; then
  ncu --set full --clock-control none --import-source on -k regex:k_ivfpq_scan_fast -s 1 -c 1 -o gpurun_out/prof_scan_$TAG python bench.py --steps 1 --warmup 1 --profile > gpurun_out/ncu_$TAG.log 2>&1
  tail -n 2 gpurun_out/ncu_$TAG.log
fi
