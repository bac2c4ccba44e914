# Phase trace of k_nltgv2_grid: rebuild with -DFBG_TRACE, print cycles per phase, restore the normal build.
FB_NVCC_EXTRA="-DFBG_TRACE" python -c "from flame_ros_b200 import build; build.build(force=True)"
timeout 120 python scripts/grid_trace.py 1 C2
timeout 120 python scripts/grid_trace.py 8 C2
timeout 120 python scripts/grid_trace.py 1 C4
python -c "from flame_ros_b200 import build; build.build(force=True)"
