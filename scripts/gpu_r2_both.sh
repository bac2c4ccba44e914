timeout 900 python -m pytest tests/test_gpu_delaunay.py tests/test_gpu_update.py -q -x 2>&1 | tail -3
bash scripts/gpu_r2_launches.sh 2>&1 | tail -26 | head -12
(timeout 120 python scripts/profile_update.py 8 0; timeout 120 python scripts/profile_update.py 16 0) 2>&1 | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
