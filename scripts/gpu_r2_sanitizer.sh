# compute-sanitizer over the round-2 kernels (device triangulation, graph build, tile / coop solvers, fb_update, pipelined step).
mkdir -p gpurun_out
SEL1='tests/test_gpu_delaunay.py -k "uniform-100 or tiny-3 or integer-dense or cocircular-20-perm0 or collinear or clustered or thin-strip or degree_overflow"'
SEL2='tests/test_gpu_nltgv2.py -k "tile or plan_free"'
SEL3='tests/test_gpu_update.py -k "not win8 and not 8-"'
SEL4='tests/test_gpu_hotpath_step.py'
( eval timeout 900 compute-sanitizer --tool memcheck --error-exitcode 0 python -m pytest $SEL1 $SEL2 -q -x ) > gpurun_out/r2_memcheck_a.log 2>&1; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_memcheck_a.log | tail -3
( eval timeout 900 compute-sanitizer --tool memcheck --error-exitcode 0 python -m pytest $SEL4 -q -x ) > gpurun_out/r2_memcheck_b.log 2>&1; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_memcheck_b.log | tail -3
( eval timeout 900 compute-sanitizer --tool racecheck --error-exitcode 0 python -m pytest $SEL1 $SEL2 -q -x ) > gpurun_out/r2_racecheck_a.log 2>&1; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r2_racecheck_a.log | tail -3
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 0 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2_memcheck_smoke.log 2>&1; grep -E "ERROR SUMMARY|smoke ok" gpurun_out/r2_memcheck_smoke.log | tail -3
( timeout 900 compute-sanitizer --tool racecheck --error-exitcode 0 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2_racecheck_smoke.log 2>&1; grep -E "RACECHECK SUMMARY|smoke ok" gpurun_out/r2_racecheck_smoke.log | tail -3
