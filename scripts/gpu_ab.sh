set -x
mkdir -p gpurun_out
for cfg in "512 4 2" "1024 2 1" "256 8 4" "384 6 3"; do
set -- $cfg
FB_NVCC_EXTRA="-DFBC_THREADS=$1 -DFBC_EPT=$2 -DFBC_VPT=$3" python -c "from flame_ros_b200 import build; build.build(force=True, verbose=True)" 2>&1 | grep -A2 "k_nltgv2_cluster" | grep -E "registers|spill"
timeout 100 python -m pytest tests/test_gpu_nltgv2.py -q -x 2>&1 | tail -1
for S in 1 8; do
timeout 200 python bench.py --streams $S --steps 30 --warmup 5 --no-single --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('CFG $cfg S=$S solver_us', round(d['roofline']['launch_us'],1), 'value', round(d['value']))"
done
done
python -c "from flame_ros_b200 import build; build.build(force=True)"
