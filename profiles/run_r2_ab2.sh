#!/bin/bash
# same-box A/B of scan-kernel variants built into ab/*.so (MMIDX_LIB_PATH selects the library): step times of configs[2] and config 4
mkdir -p gpurun_out
for rep in $(seq 1 ${REPS:-1}); do
for v in "$@"; do
  MMIDX_LIB_PATH=$PWD/ab/lib_$v.so timeout 600 python bench.py --steps 30 --warmup 6 --quick --no-cpu-baseline > gpurun_out/ab_${v}_cfg3.json 2> gpurun_out/ab_${v}_cfg3.err
  MMIDX_LIB_PATH=$PWD/ab/lib_$v.so timeout 900 python bench.py --config 4 --steps 10 --warmup 6 --no-cpu-baseline > gpurun_out/ab_${v}_cfg4.json 2> gpurun_out/ab_${v}_cfg4.err
  python - $v <<'PY'
import json, sys
v = sys.argv[1]
out = [v]
for c in ("cfg3", "cfg4"):
    try:
        d = json.loads(open(f"gpurun_out/ab_{v}_{c}.json").read().strip().splitlines()[-1])
        out.append(f"{c}: step {d['ms_per_step']:.4f} scan {d['stage_ms_per_step']['scan']:.4f} parity {list(d['parity'].values())}")
    except Exception as e:
        out.append(f"{c}: ERR {e!r}")
print(" | ".join(out))
PY
done
done
