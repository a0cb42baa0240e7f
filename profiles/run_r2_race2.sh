#!/bin/bash
# memcheck of the smoke search; racecheck of two small parity tests (ties at the k-th boundary: rank sort + tie replay;
# coarse probes: register-key verification, pre-split GEMM) -- final kernels of round 2
mkdir -p gpurun_out
timeout 60 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_memcheck_smoke.log 2>&1
grep -E "ERROR SUMMARY|smoke ok" gpurun_out/r2_memcheck_smoke.log | head -n 4
timeout 100 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ties_at_the_kth or coarse_probes_vs_oracle and fast" > gpurun_out/r2_racecheck_tests.log 2>&1
echo "rc=$?"
grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/r2_racecheck_tests.log | head -n 8
