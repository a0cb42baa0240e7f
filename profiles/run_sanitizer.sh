#!/bin/bash
# compute-sanitizer passes over the small parity tests (racecheck: shared-memory hazards of the barrier-free sweep, the
# in-place table build, the collectors, the filtered argmins; memcheck: out-of-bounds), incl. the 4-thread concurrency test
mkdir -p gpurun_out
SEL='test_ivfpq_vs_oracle and fast or test_coarse_probes_vs_oracle or test_ties_at_the_kth_boundary or test_filtered_argmins'
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 30 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$SEL" > gpurun_out/sanitizer_race.log 2>&1
tail -n 6 gpurun_out/sanitizer_race.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 30 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -x -q -m gpu -k "$SEL or test_barrier_free or test_fast_path_overflow or test_concurrent or test_multi_step_on_one_gpu or test_vlad_multi or test_pca or test_random_rotation" > gpurun_out/sanitizer_mem.log 2>&1
tail -n 6 gpurun_out/sanitizer_mem.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 30 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "test_concurrent" > gpurun_out/sanitizer_race_threads.log 2>&1
tail -n 4 gpurun_out/sanitizer_race_threads.log
