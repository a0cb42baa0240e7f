#!/bin/bash
# 1-GPU record pass at the end of round 2: whole GPU suite, index-build A/B of the filtered coarse assignment at nlist = 8192,
# default bench (configs[2]) and config 4 on one GPU
mkdir -p gpurun_out
timeout 1200 python -X faulthandler -m pytest tests -m gpu -x -q > gpurun_out/r2_final_pytest.log 2>&1
tail -4 gpurun_out/r2_final_pytest.log | cut -c1-300
for mode in fast exact; do
  MMIDX_MODE=$mode timeout 600 python bench.py --config 4 --n-db 3000000 --steps 3 --warmup 6 --no-cpu-baseline > gpurun_out/r2_ab_$mode.json 2> gpurun_out/r2_ab_$mode.err
  echo "mode=$mode"; grep -i "indexed\|vectors/s" gpurun_out/r2_ab_$mode.err | tail -3
done
timeout 900 python bench.py --steps 20 --warmup 6 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref1.json 2> gpurun_out/r2_bench_ref1.err
timeout 1200 python bench.py --config 4 --steps 10 --warmup 6 > gpurun_out/r2_cfg4_1.json 2> gpurun_out/r2_cfg4_1.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench1.json", "gpurun_out/r2_bench_ref1.json", "gpurun_out/r2_cfg4_1.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d.get("ms_per_step"), d.get("median_ms_per_step"), json.dumps(d.get("stage_ms_per_step")))
        for k in ("parity", "e2e", "roofline", "ties", "cpu_baseline", "small_batch", "rows", "clocks", "index_vectors_per_s"):
            if k in d: print(" ", k, json.dumps(d[k])[:1500])
    except Exception as e:
        print(f, "ERR", e)
PY
for f in r2_bench1 r2_cfg4_1; do grep -v "^\*\|OMP" gpurun_out/$f.err | tail -8; done
