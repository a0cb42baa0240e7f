#!/bin/bash
# compute-sanitizer racecheck of the smoke search (tiny IVFPQ index, the fused scan with descriptor staging, rank sort, 3-pass
# selection; the coarse GEMM with cp.async / ldmatrix; the rewritten prep) -- final kernels of round 2
mkdir -p gpurun_out
timeout 170 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_racecheck_smoke.log 2>&1
echo "rc=$?"
grep -E "RACECHECK SUMMARY|ERROR SUMMARY|smoke ok|hazard" gpurun_out/r2_racecheck_smoke.log | head -n 12
tail -n 3 gpurun_out/r2_racecheck_smoke.log | cut -c1-300
