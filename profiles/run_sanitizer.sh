#!/bin/bash
# compute-sanitizer passes over the small parity tests of the fast path (racecheck: shared-memory hazards of the
# barrier-free sweep / collectors; memcheck: out-of-bounds)
mkdir -p gpurun_out
SEL='test_ivfpq_vs_oracle and fast or test_coarse_probes_vs_oracle or test_ties_at_the_kth_boundary'
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 30 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$SEL" > gpurun_out/sanitizer_race.log 2>&1
tail -n 15 gpurun_out/sanitizer_race.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 30 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$SEL or test_barrier_free or test_fast_path_overflow" > gpurun_out/sanitizer_mem.log 2>&1
tail -n 8 gpurun_out/sanitizer_mem.log
