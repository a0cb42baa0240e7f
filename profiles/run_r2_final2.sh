#!/bin/bash
# 1-GPU record pass on the final kernels of round 2: GPU suite, default bench, reference arm, config 4, then the ncu launch
# list of a search step (index-build launches excluded by starting after them) and a full capture of the config-4 scan
mkdir -p gpurun_out
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
        for k in ("parity", "e2e", "roofline", "ties", "cpu_baseline", "small_batch", "rows", "clocks", "index_vectors_per_s", "gpu_launches"):
            if k in d: print(" ", k, json.dumps(d[k])[:1500])
    except Exception as e:
        print(f, "ERR", e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 2 --warmup 1 --profile > gpurun_out/ncu_launch_final.log 2>&1
tail -n 1 gpurun_out/ncu_launch_final.log
ncu --set full --clock-control none --import-source on -k regex:k_ivfpq_scan_fast -s 2 -c 1 -o gpurun_out/prof_r2_cfg4_scan2 python bench.py --config 4 --steps 1 --warmup 2 --profile > gpurun_out/ncu_r2_cfg4b.log 2>&1
tail -n 2 gpurun_out/ncu_r2_cfg4b.log
