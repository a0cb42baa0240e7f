#!/bin/bash
# 2-GPU check of graph replay in the multi-GPU step + launch list of an index build at nlist = 8192
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/r2_n2c_pytest.log 2>&1
tail -5 gpurun_out/r2_n2c_pytest.log | cut -c1-400
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 6 --no-cpu-baseline > gpurun_out/r2_n2c_bench.json 2> gpurun_out/r2_n2c_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_n2c_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["stage_ms_per_step"], d["launches_per_step"])
for k in ("strong_scaling_10k", "list_sharded", "parity", "e2e"):
    print(k, json.dumps(d[k])[:900])
PY
grep -v "^\*\|OMP" gpurun_out/r2_n2c_bench.err | tail -5
CUDA_VISIBLE_DEVICES=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_build_launches.csv python bench.py --config 4 --n-db 200000 --steps 1 --warmup 3 --profile > gpurun_out/r2_build_launches.log 2>&1
python - <<'PY'
import csv, re, collections
rows = list(csv.reader(open("gpurun_out/r2_build_launches.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[hdr + 1:]:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("mmidx::", "")
    tot[name] += int(r[-1]); cnt[name] += 1
for k, v in tot.most_common(14): print(f"{k:50s} n={cnt[k]:4d} total_us={v/1e3:10.1f} avg_us={v/1e3/cnt[k]:9.1f}")
PY
