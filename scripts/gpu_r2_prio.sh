mkdir -p gpurun_out
timeout 600 python scripts/update_scaling.py 2>&1 | tee gpurun_out/r2_update_scaling_prio.log | grep -v "^{" | tail -8
FB_TILE_NO_PRIORITY=1 timeout 600 python scripts/update_scaling.py 2>&1 | tee gpurun_out/r2_update_scaling_noprio.log | grep -v "^{" | tail -8
timeout 900 python -m pytest tests/test_gpu_update.py tests/test_gpu_delaunay.py -q -x 2>&1 | tail -3
