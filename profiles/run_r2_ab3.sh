#!/bin/bash
# parity subset on the in-tree build, then the same-box A/B of the variants named on the command line with stage times
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -n 2 | cut -c1-300
bash profiles/run_r2_ab2.sh "$@"
python - "$@" <<'PY'
import json, sys
for v in dict.fromkeys(sys.argv[1:]):
    for c in ("cfg3", "cfg4"):
        d = json.loads(open(f"gpurun_out/ab_{v}_{c}.json").read().strip().splitlines()[-1])
        print(v, c, {k: round(x, 4) for k, x in d["stage_ms_per_step"].items()}, d.get("index_vectors_per_s"))
PY
