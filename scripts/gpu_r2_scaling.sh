mkdir -p gpurun_out
timeout 600 python scripts/update_scaling.py 2>&1 | tee gpurun_out/r2_update_scaling.log | tail -12
FB_UPDATE_GRAPH=0 timeout 600 python scripts/update_scaling.py 2>&1 | tee gpurun_out/r2_update_scaling_nograph.log | tail -10
timeout 900 python -m pytest tests/test_gpu_update.py tests/test_gpu_delaunay.py -q -x 2>&1 | tail -3
