"""GPU parity: epipolar inverse-depth update + feature projection vs the CPU oracle.

Statuses, dropout counters and alive flags are integers -> bit-exact.  mu / var / u_cmp are fp32
computed with the oracle's expression order -> asserted to TOL (and reported when bit-exact).
"""
import numpy as np
import pytest

from flame_ros_b200 import synth
from helpers import init_features, scene_frames

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _oracle_params(oracle, capi_params):
    p = oracle.EpiParams.default()
    for name, _ in p._fields_:
        setattr(p, name, getattr(capi_params, name))
    return p


def _run_both(capi, oracle, imgs, poses, K, feats, cmp_slots, params=None):
    n_slots, H, W = imgs.shape
    ep = params or capi.default_epi_params()
    op = _oracle_params(oracle, ep)
    ref = {k: v.copy() for k, v in feats.items()}
    out_ref, out_gpu = [], []
    with capi.Context(1, W, H, n_slots, len(feats["mu"]), 16, 16) as ctx:
        ctx.set_intrinsics(0, K)
        ctx.set_epi_params(ep)
        for s in range(n_slots):
            ctx.frame_set(0, s, imgs[s], poses[s])
        ctx.features_set(0, feats["u_ref"], feats["ref_slot"], feats["mu"], feats["var"],
                         feats["dropouts"], feats["alive"])
        for cs in cmp_slots:
            st, uc, cnt = oracle.idepth_update(imgs, poses, K, cs, ref["ref_slot"], ref["u_ref"], ref["mu"],
                                               ref["var"], ref["dropouts"], ref["alive"], op)
            out_ref.append(dict(status=st, u_cmp=uc, counters=cnt, mu=ref["mu"].copy(), var=ref["var"].copy(),
                                dropouts=ref["dropouts"].copy(), alive=ref["alive"].copy()))
            ctx.idepth_update(cs)
            g = ctx.features_get(0)
            g["counters"] = ctx.idepth_counters(0)
            out_gpu.append(g)
    return out_ref, out_gpu


def _compare(out_ref, out_gpu):
    for r, g in zip(out_ref, out_gpu):
        assert np.array_equal(r["status"], g["status"])
        assert np.array_equal(r["counters"], g["counters"])
        assert np.array_equal(r["dropouts"], g["dropouts"])
        assert np.array_equal(r["alive"], g["alive"])
        assert np.max(np.abs(r["mu"] - g["mu"])) < TOL
        assert np.max(np.abs(r["var"] - g["var"])) < TOL
        m = np.isfinite(r["u_cmp"][:, 0])
        assert np.array_equal(m, np.isfinite(g["u_cmp"][:, 0]))
        assert np.max(np.abs(r["u_cmp"][m] - g["u_cmp"][m]), initial=0.0) < TOL


def test_vga_stream_parity(capi, oracle):
    imgs, ids, poses = scene_frames(6)
    feats = init_features(640, 480, 8)
    out_ref, out_gpu = _run_both(capi, oracle, imgs, poses, synth.K_VGA, feats, [1, 2, 3, 4, 5])
    _compare(out_ref, out_gpu)
    assert out_gpu[-1]["counters"][0] > 3000  # most features measured
    # and the estimate approaches the rendered ground truth
    truth = ids[0][feats["u_ref"][:, 1].astype(int), feats["u_ref"][:, 0].astype(int)]
    ok = out_gpu[-1]["alive"] == 1
    assert np.median(np.abs(out_gpu[-1]["mu"][ok] - truth[ok])) < 0.02


def test_euroc_shape_parity(capi, oracle):
    imgs, ids, poses = scene_frames(4, 752, 480, synth.K_EUROC, seed=3)
    feats = init_features(752, 480, 16)
    out_ref, out_gpu = _run_both(capi, oracle, imgs, poses, synth.K_EUROC, feats, [1, 2, 3])
    _compare(out_ref, out_gpu)


def test_rotation_and_forward_motion_parity(capi, oracle):
    """Non-trivial relative pose: yaw + forward translation (epipole inside the image)."""
    W, H = 320, 240
    K = np.array([[260.0, 0, 159.5], [0, 260.0, 119.5], [0, 0, 1]], np.float32)
    sc = synth.Scene(5, tex_size=1024)
    poses = np.zeros((3, 7), np.float32)
    poses[:, 3] = 1.0
    a = 0.02
    poses[1] = [0, np.sin(a / 2), 0, np.cos(a / 2), 0.03, 0.0, 0.05]
    poses[2] = [np.sin(-a / 2), 0, 0, np.cos(a / 2), -0.02, 0.02, 0.10]
    imgs = np.stack([sc.render(K, p, W, H)[0] for p in poses])
    feats = init_features(W, H, 8)
    out_ref, out_gpu = _run_both(capi, oracle, imgs, poses, K, feats, [1, 2])
    _compare(out_ref, out_gpu)


def test_failure_modes_parity(capi, oracle):
    """Flat image -> FAIL_REF_PATCH_GRADIENT; repeated failures -> FAIL_MAX_DROPOUTS kills features;
    cmp == ref -> NO_PARALLAX leaves features untouched; dead features are SKIPPED."""
    W, H = 160, 120
    K = np.array([[130.0, 0, 79.5], [0, 130.0, 59.5], [0, 0, 1]], np.float32)
    imgs = np.full((3, H, W), 128, np.uint8)
    poses = synth.stream_poses(3, step=0.05)
    feats = init_features(W, H, 16)
    feats["alive"][::7] = 0
    ep = capi.default_epi_params()
    ep.max_dropouts = 2
    out_ref, out_gpu = _run_both(capi, oracle, imgs, poses, K, feats, [0, 1, 2, 1, 2], ep)
    _compare(out_ref, out_gpu)
    assert out_gpu[0]["counters"][7] == (feats["alive"] == 1).sum()       # NO_PARALLAX
    assert out_gpu[1]["counters"][1] == (feats["alive"] == 1).sum()       # gradient failures
    assert out_gpu[3]["counters"][5] == (feats["alive"] == 1).sum()       # killed on 3rd failure
    assert out_gpu[4]["counters"].sum() == 0 and np.all(out_gpu[4]["status"] == 8)


def test_out_of_image_and_wide_window(capi, oracle):
    """Features near the border with a wide window / long search exercise the inside tests."""
    imgs, ids, poses = scene_frames(3, 320, 240, np.array([[260.0, 0, 159.5], [0, 260.0, 119.5], [0, 0, 1]], np.float32), step=0.04)
    K = np.array([[260.0, 0, 159.5], [0, 260.0, 119.5], [0, 0, 1]], np.float32)
    rng = np.random.default_rng(3)
    N = 600
    u = np.stack([rng.uniform(0, 319, N), rng.uniform(0, 239, N)], axis=1).astype(np.float32)
    feats = dict(u_ref=u, ref_slot=np.zeros(N, np.int32), mu=np.full(N, 0.4, np.float32),
                 var=np.full(N, 1.0, np.float32), dropouts=np.zeros(N, np.int32), alive=np.ones(N, np.int32))
    ep = capi.default_epi_params()
    ep.win_size = 9
    ep.max_search_px = 200
    out_ref, out_gpu = _run_both(capi, oracle, imgs, poses, K, feats, [1, 2], ep)
    _compare(out_ref, out_gpu)
    assert out_gpu[0]["counters"][6] > 0  # some FAIL_OUT_OF_IMAGE


def test_project_features_parity(capi, oracle):
    imgs, ids, poses = scene_frames(3)
    feats = init_features(640, 480, 16)
    rng = np.random.default_rng(1)
    feats["mu"] = rng.uniform(0.2, 0.8, len(feats["mu"])).astype(np.float32)
    feats["alive"][::5] = 0
    with capi.Context(1, 640, 480, 3, len(feats["mu"]), 16, 16) as ctx:
        ctx.set_intrinsics(0, synth.K_VGA)
        for s in range(3):
            ctx.frame_set(0, s, imgs[s], poses[s])
        ctx.features_set(0, feats["u_ref"], feats["ref_slot"], feats["mu"], feats["var"], None, feats["alive"])
        u, mu, var, valid = ctx.project_features(0, 2)
    ru, rmu, rvar, rvalid = oracle.project_features(640, 480, poses, synth.K_VGA, 2, feats["ref_slot"],
                                                    feats["u_ref"], feats["mu"], feats["var"], feats["alive"])
    assert np.array_equal(valid, rvalid)
    m = valid == 1
    assert np.max(np.abs(u[m] - ru[m])) < TOL and np.max(np.abs(mu[m] - rmu[m])) < TOL
    assert np.max(np.abs(var[m] - rvar[m])) < TOL


def test_batched_streams_epipolar(capi, oracle):
    """Two streams with different images/poses/intrinsics in one launch."""
    W, H = 320, 240
    Ks = [np.array([[260.0, 0, 159.5], [0, 260.0, 119.5], [0, 0, 1]], np.float32),
          np.array([[300.0, 0, 150.0], [0, 290.0, 125.0], [0, 0, 1]], np.float32)]
    data = [scene_frames(3, W, H, Ks[s], seed=s, step=0.02 * (s + 1)) for s in range(2)]
    feats = [init_features(W, H, 8, seed=40 + s) for s in range(2)]
    with capi.Context(2, W, H, 3, 2048, 16, 16) as ctx:
        for s in range(2):
            ctx.set_intrinsics(s, Ks[s])
            for k in range(3):
                ctx.frame_set(s, k, data[s][0][k], data[s][2][k])
            f = feats[s]
            ctx.features_set(s, f["u_ref"], f["ref_slot"], f["mu"], f["var"])
        ctx.idepth_update([1, 2])
        got = [ctx.features_get(s) for s in range(2)]
    for s in range(2):
        f = {k: v.copy() for k, v in feats[s].items()}
        st, uc, cnt = oracle.idepth_update(data[s][0], data[s][2], Ks[s], [1, 2][s], f["ref_slot"], f["u_ref"],
                                           f["mu"], f["var"], f["dropouts"], f["alive"], oracle.EpiParams.default())
        assert np.array_equal(st, got[s]["status"])
        assert np.max(np.abs(f["mu"] - got[s]["mu"])) < TOL
