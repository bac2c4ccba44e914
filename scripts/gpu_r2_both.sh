bash scripts/gpu_r2_check.sh 2>&1 | head -12
bash scripts/gpu_r2_bench.sh 2>&1 | tail -45
