timeout 900 python -m pytest tests/test_gpu_nltgv2.py tests/test_gpu_update.py tests/test_gpu_delaunay.py -q -x 2>&1 | tail -4
bash scripts/gpu_r2_launches.sh 2>&1 | tail -26
bash scripts/gpu_r2_bench.sh 2>&1 | tail -60
