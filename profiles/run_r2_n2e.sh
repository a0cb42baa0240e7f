#!/bin/bash
# 2-GPU quick bench with step-time percentiles (where does the mean / median gap of the multi-GPU step come from?) + smoke + N=1 quick
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 6 --quick --no-cpu-baseline > gpurun_out/r2_n2e_bench.json 2> gpurun_out/r2_n2e_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_n2e_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["median_ms_per_step"], d.get("step_ms_percentiles"), d["stage_ms_per_step"])
print("e2e", d["e2e"]["ms_per_step"], d["e2e"].get("step_ms_percentiles"))
PY
CUDA_VISIBLE_DEVICES=0 timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 5 --warmup 3 --quick --no-cpu-baseline 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n1', d['value'], d['ms_per_step'], d.get('step_ms_percentiles'), d['e2e']['value'])"
