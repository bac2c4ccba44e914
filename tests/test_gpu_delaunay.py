"""GPU parity of the device triangulation (csrc/delaunay_gpu.cuh, one warp per vertex) against the host
triangulator, and of the two fb_update graph paths (device / host sync_graph + triangulate) against
each other: identical meshes, identical solver state, frame by frame."""
import numpy as np
import pytest

from flame_ros_b200 import synth
from test_delaunay_star import point_sets

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,pts,cell", [p for p in point_sets() if not p[0].startswith("small-lattice") or p[0].endswith(("0", "5"))],
                         ids=lambda v: v if isinstance(v, str) else None)
def test_device_triangulation_equals_host(capi, name, pts, cell):
    pts = np.ascontiguousarray(pts, np.float32)
    n = len(pts)
    try:
        ht, he = capi.delaunay(pts)
    except capi.FlameError:
        ht = he = None
    W, H = (1280, 720) if n > 8000 else (640, 480)
    with capi.Context(1, W, H, 3, max(n, 16), max(n, 16), 3 * max(n, 16)) as ctx:
        if ht is None:
            with pytest.raises(capi.FlameError):
                ctx.delaunay_device(0, pts)
            return
        tris, edges = ctx.delaunay_device(0, pts)
        assert np.array_equal(tris, ht), name
        assert np.array_equal(edges, he), name
        # a second call on the same context (scratch reuse) gives the same answer
        tris2, edges2 = ctx.delaunay_device(0, pts[::-1].copy())
        ht2, he2 = capi.delaunay(pts[::-1].copy())
        assert np.array_equal(tris2, ht2) and np.array_equal(edges2, he2)


def test_device_triangulation_reports_degree_overflow(capi):
    pts = np.concatenate([np.stack([np.arange(50) * 3.0, np.full(50, 7.0)], 1), [[70., 60.]]]).astype(np.float32)
    with capi.Context(1, 640, 480, 3, 64, 64, 192) as ctx:
        with pytest.raises(capi.FlameError):
            ctx.delaunay_device(0, pts)


@pytest.mark.parametrize("win,n_frames", [(16, 14), (8, 20)])
def test_update_device_and_host_graph_paths_agree(capi, win, n_frames):
    """fb_update with triangulator = 0 (device) and = 1 (host) on the same VGA stream: same mesh, same
    vertex inverse depths (bit for bit), same dense map, same feature pool, every frame.  win = 8 is
    the C2 setting (~5k vertices)."""
    W, H, K = 640, 480, synth.K_VGA
    sc = synth.Scene(4, tex_size=1024)
    poses = synth.stream_poses(n_frames, step=0.01)
    frames = [sc.render(K, poses[k], W, H)[0] for k in range(n_frames)]
    up = capi.default_update_params()
    up.detection_win_size, up.iters = win, 50
    res = []
    for tri in (0, 1):
        up.triangulator = tri
        out = []
        with capi.Context(1, W, H, 6, 8192, 8192, 3 * 8192) as ctx:
            ctx.set_intrinsics(0, K)
            ctx.set_update_params(up)
            for k in range(n_frames):
                got = ctx.update(0, k / 30.0, k, poses[k], frames[k], k % 6 == 0)
                rec = dict(got=got, pool=ctx.get_feature_pool(0))
                if got:
                    rec["mesh"] = ctx.get_mesh(0)
                    rec["map"] = ctx.get_idepthmap(0)
                    rec["fmap"] = ctx.get_idepthmap(0, capi.default_tri_filter_params())
                    rec["nv"] = ctx.get_stat(0, "num_vtx")
                    rec["variant"] = ctx.last_solver_variant()
                out.append(rec)
        res.append(out)
    n_upd = 0
    for k in range(n_frames):
        a, b = res[0][k], res[1][k]
        assert a["got"] == b["got"], "frame %d" % k
        for key in ("alive", "mu", "var", "ref_slot", "dropouts"):
            assert np.array_equal(a["pool"][key], b["pool"][key]), "frame %d pool %s" % (k, key)
        if not a["got"]:
            continue
        n_upd += 1
        for key in ("tris", "edges", "vtx", "idepth", "normals"):
            assert np.array_equal(a["mesh"][key], b["mesh"][key]), "frame %d mesh %s" % (k, key)
        assert np.array_equal(np.nan_to_num(a["map"], nan=-1), np.nan_to_num(b["map"], nan=-1)), "frame %d map" % k
        assert np.array_equal(np.nan_to_num(a["fmap"], nan=-1), np.nan_to_num(b["fmap"], nan=-1)), "frame %d filtered" % k
        assert a["nv"] == b["nv"]
    assert n_upd >= n_frames - 5
    assert res[0][-1]["variant"] == 5      # device-planned tile-resident solver on the device-built graph
    if win == 8:
        assert res[0][-1]["nv"] > 3000
