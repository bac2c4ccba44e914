"""ROS-free offline frontends (flame_ros_b200/offline.py): on-disk formats, pose conventions, timestamp
association and the mesh output contract of the reference (CPU), and an end-to-end run of the TUM- and
ASL-shaped synthetic datasets through fb_update (GPU)."""
import numpy as np
import pytest

from flame_ros_b200 import offline as off
from flame_ros_b200 import synth


def R_of(q):
    return synth.quat_to_R(q)


def test_pose_frame_conventions_match_rotation_matrices():
    """tum_rgbd_offline_stream.cc:146-194 written with matrices instead of quaternions."""
    rng = np.random.default_rng(0)
    R_flu = np.array([[0, -1, 0], [0, 0, -1], [1, 0, 0]], float)   # FLU -> RDF: x_r=-y, y_r=-z, z_r=x
    R_frd = np.array([[0, 1, 0], [0, 0, 1], [1, 0, 0]], float)
    assert np.allclose(R_of(off.Q_FLU_TO_RDF), R_flu) and np.allclose(R_of(off.Q_FRD_TO_RDF), R_frd)
    assert np.allclose(R_of(off.Q_RFU_TO_RDF), [[1, 0, 0], [0, 0, -1], [0, 1, 0]])
    for _ in range(10):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        t = rng.normal(size=3)
        Rq = R_of(q)
        for frame, C, conj in (("FLU", R_flu, True), ("FRD", R_frd, True), ("RDF_IN_FLU", R_flu, False),
                               ("RDF_IN_FRD", R_frd, False), ("RDF", np.eye(3), False)):
            q2, t2 = off.tum_pose_to_rdf(q, t, frame)
            want = C @ Rq @ C.T if conj else C @ Rq
            assert np.allclose(R_of(q2), want, atol=1e-9), frame
            assert np.allclose(t2, C @ t, atol=1e-9)
    with pytest.raises(ValueError):
        off.tum_pose_to_rdf([0, 0, 0, 1], [0, 0, 0], "XYZ")


def test_parse_tum_line_with_and_without_depth():
    tm, t, q, rgb, depth = off.parse_tum_line("1.5 0.1 0.2 0.3 0 0 0 1 1.52 rgb/a.png 1.53 depth/a.png")
    assert (tm, t, q, rgb, depth) == (1.52, [0.1, 0.2, 0.3], [0, 0, 0, 1.0], "rgb/a.png", "depth/a.png")
    assert off.parse_tum_line("1.5 0 0 0 0 0 0 1 1.52 rgb/a.png")[4] is None
    with pytest.raises(ValueError):
        off.parse_tum_line("1.5 0 0 0")


def test_associate_matches_the_reference_semantics():
    """Brute-force restatement of dataset_utils::associate (utils.h:50-93) as the checker."""
    rng = np.random.default_rng(1)
    ta = np.sort(rng.uniform(0, 2, 40))
    tb = np.sort(np.concatenate([ta[::2] + rng.uniform(-0.015, 0.015, 20), rng.uniform(0, 2, 30)]))
    cand = sorted((float(np.float32(abs(x - y))), i, j) for i, x in enumerate(ta) for j, y in enumerate(tb)
                  if np.float32(abs(x - y)) < np.float32(0.02))
    ua, ub, ra, rb = set(), set(), [], []
    for _, i, j in sorted(cand, key=lambda c: c[0]):
        if i not in ua and j not in ub:
            ua.add(i); ub.add(j); ra.append(i); rb.append(j)
    ia, ib = off.associate(ta, tb)
    assert ia == sorted(ra) and ib == sorted(rb) and len(ia) >= 20


def test_depth_mesh_contract():
    """utils.cc:163-237: Kinv*(u,v,1)/idepth, NaN for invalid vertices, reversed winding, valid faces only."""
    K = synth.K_VGA
    mesh = dict(vtx=np.array([[319.5, 239.5], [419.5, 239.5], [319.5, 339.5], [10, 10]], np.float32),
                idepth=np.array([0.5, 0.25, np.nan, -1.0], np.float32), normals=np.zeros((4, 3), np.float32),
                tris=np.array([[0, 1, 2], [1, 2, 3]], np.int32), tri_valid=np.array([1, 0], np.uint8))
    dm = off.depth_mesh(K, 640, 480, mesh)
    assert np.allclose(dm["points"][0], [0, 0, 2.0], atol=1e-5)
    assert np.allclose(dm["points"][1], [100 / 525.0 * 4, 0, 4.0], atol=1e-4)
    assert np.all(np.isnan(dm["points"][2])) and np.all(np.isnan(dm["points"][3]))
    assert dm["faces"].tolist() == [[2, 1, 0]]
    assert np.allclose(dm["uv"][0], [319.5 / 639, 239.5 / 479])
    mesh["tri_valid"][:] = 0
    assert off.depth_mesh(K, 640, 480, mesh) is None   # nothing published without valid triangles
    d = off.idepth_to_depth(np.array([[0.5, np.nan, 0.0, -2.0]], np.float32))
    assert d[0, 0] == 2.0 and np.all(np.isnan(d[0, 1:]))


def _small_stream(n=10, W=320, H=240, seed=4):
    K = (synth.K_VGA * np.array([[0.5], [0.5], [1.0]], np.float32)).astype(np.float32)
    sc = synth.Scene(seed, tex_size=1024)
    poses = synth.stream_poses(n, step=0.02)
    # give the trajectory some rotation so the frame conventions matter
    for k in range(n):
        a = 0.004 * k
        poses[k, :4] = [0.0, np.sin(a / 2), 0.0, np.cos(a / 2)]
    rend = [sc.render(K, poses[k], W, H) for k in range(n)]
    return K, poses, [r[0] for r in rend], [r[1] for r in rend]


@pytest.mark.parametrize("frame", ["RDF_IN_FLU", "FLU", "RDF"])
def test_tum_reader_roundtrip(tmp_path, frame):
    K, poses, imgs, idepths = _small_stream(4)
    depth = [1.0 / d for d in idepths]
    assoc, calib = synth.write_tum_dataset(str(tmp_path), imgs, poses, K, frame, with_depth=depth)
    st = off.TUMStream(assoc, calib, frame)
    assert (st.width, st.height) == (320, 240) and np.allclose(st.K, K)
    for k in range(4):
        img_id, tm, gray, dep, q, t = st.get()
        assert img_id == k and np.array_equal(gray, imgs[k])
        assert np.allclose(R_of(q), R_of(poses[k, :4]), atol=1e-6) and np.allclose(t, poses[k, 4:7], atol=1e-6)
        assert np.max(np.abs(dep - depth[k])) < 1.0 / 5000 + 1e-6    # uint16 depth / 5000
    assert st.empty()


def test_asl_reader_roundtrip(tmp_path):
    K, poses, imgs, _ = _small_stream(5)
    gt, cam = synth.write_asl_dataset(str(tmp_path), imgs, poses, K, "RFU")
    st = off.ASLStream(gt, cam, "RFU")
    assert len(st.rgb_idxs) == 5 and np.allclose(st.K, K)
    for k in range(5):
        img_id, tm, gray, dep, q, t = st.get()
        assert np.array_equal(gray, imgs[k])
        # the associated pose is the exact sample, not one of the distractors 9 m away
        assert np.allclose(R_of(q), R_of(poses[k, :4]), atol=1e-6) and np.allclose(t, poses[k, 4:7], atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["tum", "asl"])
def test_offline_player_end_to_end(tmp_path, capi, fmt):
    K, poses, imgs, idepths = _small_stream(12)
    if fmt == "tum":
        assoc, calib = synth.write_tum_dataset(str(tmp_path), imgs, poses, K, "RDF_IN_FLU")
        stream = off.TUMStream(assoc, calib, "RDF_IN_FLU")
    else:
        gt, cam = synth.write_asl_dataset(str(tmp_path), imgs, poses, K, "RFU")
        stream = off.ASLStream(gt, cam, "RFU")
    up = capi.default_update_params()
    up.iters, up.idepth_var_max_graph = 20, 0.05
    errs = []

    def on_frame(img_id, tm, ctx, mesh, depth):
        dm = ctx.get_idepthmap(0)
        m = ~np.isnan(dm)
        if m.mean() > 0.2:
            errs.append(float(np.median(np.abs(dm[m] - idepths[img_id][m]))))

    out = tmp_path / "out"
    out.mkdir()
    st = off.run_offline(stream, capi, up, poseframe_subsample_factor=3, out_dir=str(out), on_frame=on_frame)
    assert st["frames"] == 12 and st["updates"] >= 8 and st["vertices_last"] > 50
    assert errs and errs[-1] < 0.05
    assert any(f.startswith("mesh_") for f in __import__("os").listdir(str(out)))
