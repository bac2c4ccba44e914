"""Phase trace of k_nltgv2_grid (FBG_TRACE build): cycles per phase seen by thread 0 of every CTA."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = sys.argv[2] if len(sys.argv) > 2 else "C2"
out = "/tmp/grid_trace.bin"
os.environ["FB_GRID_TRACE"] = out
from flame_ros_b200 import capi, synth
from helpers import gpu_load_graph
g = synth.s_graph(cfg)
V, E = len(g["pos"]), len(g["edges"])
iters = 50
with capi.Context(S, 1280, 720, 2, 16, V, E) as ctx:
    for s in range(S):
        gpu_load_graph(ctx, s, g)
    for _ in range(3):
        ctx.nltgv2_solve(iters, variant=3)
        ctx.sync()
    n = ctx.last_cluster_size() * S
    print("S", S, cfg, "ctas/stream", ctx.last_cluster_size(), "transport", ctx.last_solver_transport())
t = np.fromfile(out, dtype=np.int64).reshape(n, 64, 6)[:, :iters]
names = ["halo wait", "barrier A", "dual", "barrier B", "primal"]
d = np.diff(t, axis=2)[:, 5:45]          # [cta, it, 5]
loop = (t[:, 6:46, 0] - t[:, 5:45, 0])   # iteration period
print("iteration period: mean %.0f cycles (min cta %.0f, max cta %.0f)" % (loop.mean(), loop.mean(1).min(), loop.mean(1).max()))
for k, nm in enumerate(names):
    x = d[:, :, k]
    print("%-10s mean %6.0f  median %6.0f  p90 %6.0f  (per-CTA means %5.0f..%5.0f)" % (nm, x.mean(), np.median(x), np.percentile(x, 90), x.mean(1).min(), x.mean(1).max()))
