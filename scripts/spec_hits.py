"""How often fb_get_idepthmap(filter) is served by the map fb_update rendered (diagnosis)."""
import sys
sys.path.insert(0, ".")
import bench as B
from flame_ros_b200 import capi, workload as WL
capi.load_library()
n_frames = WL.UPD_WARMUP + 36
datas = WL.update_streams("C2", [1000 + s for s in range(8)], n_frames)
run = B.UpdateRun(capi, datas, 0)
run.run(0, n_frames)
for s, c in enumerate(run.ctxs):
    try:
        print(s, "reused", c.get_stat(0, "filtered_maps_reused"), "of", n_frames)
    except Exception as e:
        print(s, "no counter", e)
run.close()
