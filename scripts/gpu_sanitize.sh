set -x
mkdir -p gpurun_out
cat > /tmp/san_driver.py <<'PY'
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from flame_ros_b200 import capi, synth
from flame_ros_b200 import workload as WL
from helpers import small_graph, gpu_load_graph
g = small_graph(24, 18, 192, 144, seed=4)
for variant in (1, 2):
    with capi.Context(2, 192, 144, 3, 512, 512, 2048) as ctx:
        gpu_load_graph(ctx, 0, g); gpu_load_graph(ctx, 1, g)
        ctx.nltgv2_solve(12, variant=variant)
        ctx.graph_state_get(1); ctx.costs(0)
# C2-sized graph through the 8-CTA cluster (DSMEM exchange)
g2 = synth.s_graph("C2")
with capi.Context(1, 640, 480, 2, 16, 5000, 15000) as ctx:
    gpu_load_graph(ctx, 0, g2)
    import os
    os.environ["FB_CLUSTER_NO16"] = "1"
    ctx.nltgv2_solve(6, variant=2)
    ctx.graph_state_get(0)
d = WL.StreamData("tiny", seed=0)
up = capi.default_update_params(); up.iters = 5; up.idepth_var_max_graph = 0.2
with capi.Context(1, d.W, d.H, 4, 512, 512, 2048) as ctx:
    ctx.set_intrinsics(0, d.K); ctx.set_update_params(up)
    for k in range(8):
        ctx.update(0, k / 30.0, k, d.poses[k], d.frames[k], k % 3 == 0)
    ctx.get_mesh(0); ctx.get_idepthmap(0, capi.default_tri_filter_params()); ctx.get_raw_idepths(0)
print("sanitizer driver done")
PY
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/san_driver.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/sanitizer_memcheck.log
tail -5 gpurun_out/sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python /tmp/san_driver.py > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/sanitizer_racecheck.log
tail -8 gpurun_out/sanitizer_racecheck.log
