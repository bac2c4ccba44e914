# Star kernel: resident CTAs per SM (register budget 128 / 96 / 80 per thread): one camera and eight.
for m in 4 5 6; do
  FB_NVCC_EXTRA="-DDSG_STARS_MINB=$m" python -m flame_ros_b200.build > /dev/null 2>&1
  echo "== DSG_STARS_MINB=$m"
  for i in 1 2; do timeout 120 python scripts/profile_update.py 8 0 2>&1 | cut -c1-110; done
  UPD_SCALING_MODES=full UPD_SCALING_S=8 timeout 300 python scripts/update_scaling.py 2>&1 | grep "^full"
done
timeout 300 python -m pytest tests/test_gpu_delaunay.py -q -x 2>&1 | tail -1
