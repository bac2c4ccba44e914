"""Stage timings of the full fb_update pipeline (flame::Flame::update) on a synthetic VGA stream."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from flame_ros_b200 import capi, synth

W, H, win = 640, 480, int(sys.argv[1]) if len(sys.argv) > 1 else 8
tri = int(sys.argv[2]) if len(sys.argv) > 2 else 0   # 0 = device sync_graph + triangulate, 1 = host
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 50
getter = len(sys.argv) > 4 and sys.argv[4] == "getter"   # also fetch the filtered dense map every frame
n = 60
sc = synth.Scene(0, tex_size=1024)
poses = synth.stream_poses(n, step=0.01)
frames = [sc.render(synth.K_VGA, poses[k], W, H)[0] for k in range(n)]
up = capi.default_update_params()
up.detection_win_size = win
up.triangulator = tri
up.iters = iters
with capi.Context(1, W, H, 8, 8192, 8192, 24576) as ctx:
    ctx.set_intrinsics(0, synth.K_VGA)
    ctx.set_update_params(up)
    keys = ["update", "frame_creation", "update_idepths", "project_features", "triangulate", "sync_graph", "nltgv2",
            "interpolate", "detection", "num_vertices"]
    acc = {k: [] for k in keys}
    wall = []
    flt = capi.default_tri_filter_params()
    hmap = capi.PinnedBuffer((H, W), np.float32)
    for k in range(n):
        t0 = time.perf_counter()
        ok = ctx.update(0, k / 30.0, k, poses[k], frames[k], k % 6 == 0)
        if getter and ok:
            ctx.get_idepthmap(0, flt, out=hmap.array)
        wall.append(time.perf_counter() - t0)
        for key in keys:
            try:
                acc[key].append(ctx.get_stat(0, key))
            except capi.FlameError:
                pass
    out = {k: float(np.median(v[10:])) for k, v in acc.items() if len(v) > 10}
    out["wall_ms_median"] = 1e3 * float(np.median(wall[10:]))
    out["fps"] = 1.0 / float(np.median(wall[10:]))
    out["win"], out["triangulator"], out["solver_variant"] = win, tri, ctx.last_solver_variant()
    print(json.dumps(out))
