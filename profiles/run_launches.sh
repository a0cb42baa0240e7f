#!/bin/bash
# bench line + ncu launch list (per-launch gpu__time_duration, cold-cache and serialised: shares, not absolutes)
TAG=${1:-l}
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
print("qps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "stages", {k: round(v,3) for k,v in d["stage_ms_per_step"].items()}, "frac", round(d["roofline"]["frac"],3), "parity", d["parity"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --profile > gpurun_out/ncu_launch_$TAG.log 2>&1
tail -n 1 gpurun_out/ncu_launch_$TAG.log
