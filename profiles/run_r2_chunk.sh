#!/bin/bash
# end-to-end (host buffers) step time of configs[2] against the chunk size of mmidx_search's copy/compute pipeline
mkdir -p gpurun_out
for ch in ${CHUNKS:-2048 1024 1536 2560 3400 5000 2048}; do
  MMIDX_CHUNK=$ch timeout 600 python bench.py --steps 20 --warmup 6 --quick --no-cpu-baseline > gpurun_out/chunk_$ch.json 2> gpurun_out/chunk_$ch.err
  python - $ch <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/chunk_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("chunk", sys.argv[1], "device ms", round(d["ms_per_step"], 4), "e2e ms", round(d["e2e"]["ms_per_step"], 4), "median", round(d["e2e"]["median_ms_per_step"], 4), "e2e q/s", round(d["e2e"]["value"]))
PY
done
