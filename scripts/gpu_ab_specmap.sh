# e2e_update with and without the filtered map rendered inside fb_update (FB_NO_SPEC_MAP=1 disables), same box.
timeout 600 python -m pytest tests/test_gpu_update.py tests/test_gpu_raster.py -q -x 2>&1 | tail -2
for i in 1 2; do
  UPD_SCALING_MODES=full UPD_SCALING_S=1,8 timeout 300 python scripts/update_scaling.py 2>&1 | grep "^full" | sed 's/^/spec    /'
  FB_NO_SPEC_MAP=1 UPD_SCALING_MODES=full UPD_SCALING_S=1,8 timeout 300 python scripts/update_scaling.py 2>&1 | grep "^full" | sed 's/^/no spec /'
done
