"""Scaling of the full-pipeline leg (fb_update + filtered dense map) with the number of concurrent streams
(one context + one host thread each): where the multi-stream throughput goes."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench as B
from flame_ros_b200 import capi, workload as WL

capi.load_library()
K = 36
n_frames = WL.UPD_WARMUP + K
datas = WL.update_streams("C2", [1000 + s for s in range(8)], n_frames)
out = {}
import os
MODES = os.environ.get("UPD_SCALING_MODES", "full,no_map").split(",")
STREAMS = [int(x) for x in os.environ.get("UPD_SCALING_S", "1,2,4,8").split(",")]
for mode in MODES:
    for S in STREAMS:
        run = B.UpdateRun(capi, datas[:S], 0)
        if mode == "no_map":
            for c in run.ctxs:
                c.get_idepthmap = lambda *a, **k: None
        run.run(0, WL.UPD_WARMUP)
        t = run.run(WL.UPD_WARMUP, n_frames)
        run.close()
        out["%s_S%d" % (mode, S)] = {"fps": S * K / t, "ms_per_frame_per_stream": 1e3 * t / K}
        print(mode, S, out["%s_S%d" % (mode, S)], flush=True)
print(json.dumps(out))
