# Round 2 check: GPU parity of the triangulation / solvers / update pipeline, then the per-kernel table.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_delaunay.py tests/test_gpu_update.py tests/test_gpu_nltgv2.py -q -x 2>&1 | tail -15
(timeout 120 python scripts/profile_update.py 8 0; timeout 120 python scripts/profile_update.py 16 0; FB_UPDATE_GRAPH=0 timeout 120 python scripts/profile_update.py 8 0) 2>&1 | tee gpurun_out/r2_profile_update.log
bash scripts/gpu_r2_launches.sh 2>&1 | tail -30
