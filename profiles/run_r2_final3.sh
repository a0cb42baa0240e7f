#!/bin/bash
# last 1-GPU record pass of round 2 (final kernels): GPU suite, default bench, reference arm, config 4, launch list of a search
# step, HBM-streaming regime of the scan (events-timed, nq = 1, 2, 8)
mkdir -p gpurun_out
rm -f gpurun_out/r2_hbm_regime_final.jsonl
timeout 1200 python -X faulthandler -m pytest tests -m gpu -x -q > gpurun_out/r2_final_pytest.log 2>&1
tail -n 3 gpurun_out/r2_final_pytest.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 6 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref1.json 2> gpurun_out/r2_bench_ref1.err
timeout 1200 python bench.py --config 4 --steps 10 --warmup 6 > gpurun_out/r2_cfg4_1.json 2> gpurun_out/r2_cfg4_1.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench1.json", "gpurun_out/r2_bench_ref1.json", "gpurun_out/r2_cfg4_1.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d.get("ms_per_step"), d.get("median_ms_per_step"), json.dumps(d.get("stage_ms_per_step")))
        for k in ("parity", "e2e", "roofline", "cpu_baseline", "small_batch", "rows", "clocks", "index_vectors_per_s", "gpu_launches"):
            if k in d: print(" ", k, json.dumps(d[k])[:1200])
    except Exception as e:
        print(f, "ERR", e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 2 --warmup 1 --profile > gpurun_out/ncu_launch_final.log 2>&1
for nq in 1 2 8; do timeout 300 python profiles/hbm_regime.py --n 134217728 --nq $nq >> gpurun_out/r2_hbm_regime_final.jsonl 2>> gpurun_out/r2_hbm_regime_final.err; done
cat gpurun_out/r2_hbm_regime_final.jsonl | cut -c1-400
