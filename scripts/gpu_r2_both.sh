timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
bash scripts/gpu_r2_e2e.sh 2>&1 | head -1
bash scripts/gpu_r2_ncu.sh 2>&1 | tail -6
