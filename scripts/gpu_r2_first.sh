# Round 2, first GPU pass: device triangulation + plan-free solver + update pipeline parity, then timing.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_delaunay.py tests/test_gpu_nltgv2.py tests/test_gpu_update.py -q -x 2>&1 | tail -15 > gpurun_out/r2_first_tests.log
cat gpurun_out/r2_first_tests.log
(timeout 120 python scripts/profile_update.py 8 0; timeout 120 python scripts/profile_update.py 8 1; FB_COOP_DISABLE=1 timeout 120 python scripts/profile_update.py 8 0; timeout 120 python scripts/profile_update.py 16 0) > gpurun_out/r2_profile_update.log 2>&1
cat gpurun_out/r2_profile_update.log
