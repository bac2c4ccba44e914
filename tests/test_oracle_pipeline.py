"""CPU checks of the oracle's whole-pipeline restatement (oracle/flame_pipeline.c): its own Delaunay
triangulator against the product's host triangulator and Qhull, and the C pipeline against the
Python mirror (tests/pipeline_mirror.py) that round 1's GPU tests were pinned to."""
import numpy as np
import pytest

from flame_ros_b200 import synth
from pipeline_mirror import MirrorFlame
from test_delaunay_star import point_sets


@pytest.mark.parametrize("name,pts,cell", [p for p in point_sets() if not p[0].startswith("small-lattice") or p[0].endswith(("1", "7"))],
                         ids=lambda v: v if isinstance(v, str) else None)
def test_oracle_triangulator_equals_product_host_triangulator(capi, oracle, name, pts, cell):
    """Three independently written triangulators (oracle: sorted sweep + flips; product host: incremental
    Bowyer-Watson with ghost triangles; product device: per-vertex stars) share one canonical output."""
    pts = np.ascontiguousarray(pts, np.float32)
    try:
        ht, he = capi.delaunay(pts)
    except capi.FlameError:
        with pytest.raises(ValueError):
            oracle.delaunay(pts)
        return
    ot, oe = oracle.delaunay(pts)
    assert np.array_equal(ot, ht) and np.array_equal(oe, he)


def test_oracle_triangulator_matches_qhull(oracle):
    scipy_spatial = pytest.importorskip("scipy.spatial")
    rng = np.random.default_rng(11)
    pts = (np.round(rng.uniform(0, [752, 480], (5000, 2)) * 64) / 64).astype(np.float32)
    tris, edges = oracle.delaunay(pts)
    q = scipy_spatial.Delaunay(pts.astype(np.float64)).simplices
    canon = lambda t: {tuple(sorted(map(int, r))) for r in t}
    assert canon(tris) == canon(q)
    a, b, c = pts[tris[:, 0]], pts[tris[:, 1]], pts[tris[:, 2]]
    assert np.all((b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0]) > 0)
    assert np.all(tris[:, 0] < tris[:, 1]) and np.all(tris[:, 0] < tris[:, 2])


@pytest.mark.parametrize("W,H,win,iters,pf_every,n_frames", [(320, 240, 16, 20, 3, 14), (640, 480, 8, 10, 6, 14)])
def test_c_pipeline_equals_python_mirror(capi, oracle, W, H, win, iters, pf_every, n_frames):
    K = (synth.K_VGA * np.array([[W / 640.0], [H / 480.0], [1.0]], np.float32)).astype(np.float32)
    sc = synth.Scene(1, tex_size=1024)
    poses = synth.stream_poses(n_frames, step=0.02)
    frames = [sc.render(K, poses[k], W, H)[0] for k in range(n_frames)]
    up = oracle.UpdateParams.default()
    up.detection_win_size, up.iters, up.idepth_var_max_graph = win, iters, 0.05
    n_slots, maxF, maxV = 4, 8192, 8192
    mir = MirrorFlame(oracle, capi, W, H, K, n_slots, maxF, maxV, up)
    n_upd = 0
    with oracle.Pipeline(W, H, K, n_slots, maxF, maxV, up) as pipe:
        for k in range(n_frames):
            is_pf = (k % pf_every) == 0
            got = pipe.update(k, poses[k], frames[k], is_pf)
            ref = mir.update(k / 30.0, k, poses[k], frames[k], is_pf)
            assert got == ref, "frame %d" % k
            f = pipe.features()
            assert np.array_equal(f["alive"], mir.alive)
            live = mir.alive == 1
            for key in ("mu", "var"):
                assert np.array_equal(f[key][live], getattr(mir, key)[live]), "frame %d %s" % (k, key)
            assert np.array_equal(f["ref_slot"][live], mir.ref_slot[live])
            if got:
                n_upd += 1
                m = pipe.mesh()
                assert np.array_equal(m["tris"], mir.tris) and np.array_equal(m["edges"], mir.edges)
                assert np.array_equal(m["vtx"], mir.pos) and np.array_equal(m["idepth"], mir.state["x"])
                dm = pipe.idepthmap()
                assert np.array_equal(np.nan_to_num(dm, nan=-1), np.nan_to_num(mir.idmap, nan=-1))
        assert n_upd >= n_frames - 5
        ms = pipe.stage_ms()
        assert ms["update"] > 0 and ms["triangulate"] > 0


def test_pipeline_options_semantics(oracle):
    """The options of SURVEY row a10 on the oracle pipeline: the height band removes the far plane's
    vertices, the letterbox confines detections to the middle third, rescale_data changes the
    transient (different effective step sizes) but not the scale of the answer."""
    W, H, K = 320, 240, (synth.K_VGA * np.array([[0.5], [0.5], [1.0]], np.float32)).astype(np.float32)
    n = 14
    sc = synth.Scene(5, tex_size=1024)
    poses = synth.stream_poses(n, step=0.02)
    frames = [sc.render(K, poses[k], W, H)[0] for k in range(n)]

    def run(**opts):
        up = oracle.UpdateParams.default()
        up.detection_win_size, up.iters, up.idepth_var_max_graph = 8, 30, 0.05
        for k, v in opts.items():
            setattr(up, k, v)
        with oracle.Pipeline(W, H, K, 5, 4096, 4096, up) as p:
            for k in range(n):
                p.update(k, poses[k], frames[k], k % 4 == 0)
            return p.mesh(), p.features(), p.idepthmap()

    base, fb_, mb = run()
    band, _, _ = run(min_height=0.0, max_height=3.0)
    assert 50 < len(band["vtx"]) < len(base["vtx"])
    assert (1.0 / band["idepth"]).max() < 3.6 < (1.0 / base["idepth"]).max()
    _, fl, _ = run(do_letterbox=1)
    live = fl["alive"] == 1
    assert live.sum() > 30 and np.all(fl["u_ref"][live][:, 1] >= H // 3) and np.all(fl["u_ref"][live][:, 1] < (2 * H) // 3)
    resc, _, mr = run(rescale_data=1)
    # a different iteration (other effective step sizes; the dense prediction also seeds new features,
    # so the feature sets drift apart), the same answer to first order
    assert abs(len(resc["vtx"]) - len(base["vtx"])) < 0.1 * len(base["vtx"])
    both = ~np.isnan(mr) & ~np.isnan(mb)
    assert both.mean() > 0.5 and not np.array_equal(mr[both], mb[both])
    assert np.median(np.abs(mr[both] - mb[both])) < 0.02
    up = oracle.UpdateParams.default()
    up.check_sticky_obstacles = 1
    with pytest.raises(ValueError):
        oracle.Pipeline(W, H, K, 5, 256, 256, up)
