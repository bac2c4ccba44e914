"""CPU tests of the oracle's epipolar inverse-depth update: geometry against a
float64 restatement, the fronto-parallel known-answer test, the failure taxonomy and golden vectors."""
import os

import numpy as np

from flame_ros_b200 import synth
from helpers import init_features, scene_frames

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_geometry_matches_float64(oracle):
    rng = np.random.default_rng(0)
    K = synth.K_EUROC
    for _ in range(20):
        pr = np.concatenate([rng.normal(size=4), rng.normal(size=3)]).astype(np.float32)
        pc = np.concatenate([rng.normal(size=4), rng.normal(size=3)]).astype(np.float32)
        G = oracle.epi_geometry(K, pr, pc)
        Rr, Rc = synth.quat_to_R(pr[:4]), synth.quat_to_R(pc[:4])
        R = Rc.T @ Rr
        t = Rc.T @ (pr[4:].astype(np.float64) - pc[4:].astype(np.float64))
        K64 = K.astype(np.float64)
        A = K64 @ R @ np.linalg.inv(K64)
        scale = max(1.0, np.abs(A).max())
        assert np.max(np.abs(G[:9].reshape(3, 3) - A)) < 2e-4 * scale
        assert np.max(np.abs(G[9:12] - K64 @ t)) < 1e-3
        assert np.max(np.abs(G[12:15] - K64 @ (-R.T @ t))) < 1e-3


def test_identity_pose_gives_identity_geometry(oracle):
    p = np.array([0, 0, 0, 1, 0.3, -0.2, 0.1], np.float32)
    G = oracle.epi_geometry(synth.K_VGA, p, p)
    assert np.allclose(G[:9].reshape(3, 3), np.eye(3), atol=1e-5) and np.allclose(G[9:15], 0, atol=1e-6)


def test_fronto_parallel_plane_known_answer(oracle):
    """Plane at depth d, pure x translation b: disparity f*b/d, idepth -> 1/d."""
    W, H, d = 320, 240, 2.5
    K = np.array([[260.0, 0, 159.5], [0, 260.0, 119.5], [0, 0, 1]], np.float32)
    sc = synth.Scene(1, tex_size=1024, tilt_deg=0.0, near=d, far=d, x_split=1e9)
    poses = synth.stream_poses(5, step=0.03, wobble=0.0)
    imgs = np.stack([sc.render(K, p, W, H)[0] for p in poses])
    f = init_features(W, H, 16, mu0=0.5, var0=0.1)
    for cs in (1, 2, 3, 4):
        st, uc, cnt = oracle.idepth_update(imgs, poses, K, cs, f["ref_slot"], f["u_ref"], f["mu"], f["var"],
                                           f["dropouts"], f["alive"], oracle.EpiParams.default())
        ok = st == 0
        # matched pixel = u_ref - f*b/d along x (camera moves +x => scene moves -x)
        disp = 260.0 * poses[cs, 4] / d
        assert np.median(np.abs((f["u_ref"][ok, 0] - uc[ok, 0]) - disp)) < 0.1
        assert np.median(np.abs(uc[ok, 1] - f["u_ref"][ok, 1])) < 0.05
    good = (f["alive"] == 1) & (f["var"] < 0.05)
    assert good.sum() > 0.7 * len(good)
    assert np.median(np.abs(f["mu"][good] - 1.0 / d)) < 0.01
    # the filter's own uncertainty is honest: most errors fall within 3 sigma
    assert np.mean(np.abs(f["mu"][good] - 1.0 / d) < 3.0 * np.sqrt(f["var"][good])) > 0.9


def test_variance_shrinks_and_estimates_converge_on_vga_scene(oracle):
    imgs, ids, poses = scene_frames(6)
    f = init_features(640, 480, 16)
    truth = ids[0][f["u_ref"][:, 1].astype(int), f["u_ref"][:, 0].astype(int)]
    v_prev = f["var"].mean()
    for cs in range(1, 6):
        st, uc, cnt = oracle.idepth_update(imgs, poses, synth.K_VGA, cs, f["ref_slot"], f["u_ref"], f["mu"],
                                           f["var"], f["dropouts"], f["alive"], oracle.EpiParams.default())
        assert cnt.sum() == (st != 8).sum()
        assert f["var"].mean() < v_prev
        v_prev = f["var"].mean()
    ok = f["alive"] == 1
    assert np.median(np.abs(f["mu"][ok] - truth[ok])) < 0.02


def test_failure_taxonomy(oracle):
    W, H = 160, 120
    K = np.array([[130.0, 0, 79.5], [0, 130.0, 59.5], [0, 0, 1]], np.float32)
    flat = np.full((3, H, W), 100, np.uint8)
    poses = synth.stream_poses(3, step=0.05)
    f = init_features(W, H, 16)
    ep = oracle.EpiParams.default()
    ep.max_dropouts = 1
    n = len(f["mu"])
    mu0, var0 = f["mu"].copy(), f["var"].copy()
    st, _, cnt = oracle.idepth_update(flat, poses, K, 0, f["ref_slot"], f["u_ref"], f["mu"], f["var"],
                                      f["dropouts"], f["alive"], ep)
    assert cnt[7] == n and np.all(f["dropouts"] == 0)  # cmp == ref: NO_PARALLAX, untouched
    st, _, cnt = oracle.idepth_update(flat, poses, K, 1, f["ref_slot"], f["u_ref"], f["mu"], f["var"],
                                      f["dropouts"], f["alive"], ep)
    assert cnt[1] == n and np.all(f["dropouts"] == 1)  # flat image: no gradient
    st, _, cnt = oracle.idepth_update(flat, poses, K, 2, f["ref_slot"], f["u_ref"], f["mu"], f["var"],
                                      f["dropouts"], f["alive"], ep)
    assert cnt[5] == n and np.all(f["alive"] == 0)     # second failure > max_dropouts=1: killed
    st, _, cnt = oracle.idepth_update(flat, poses, K, 1, f["ref_slot"], f["u_ref"], f["mu"], f["var"],
                                      f["dropouts"], f["alive"], ep)
    assert cnt.sum() == 0 and np.all(st == 8)          # dead slots are skipped
    assert np.array_equal(f["mu"], mu0) and np.array_equal(f["var"], var0)  # failures never touch mu/var


def test_wrong_texture_fails_max_cost_or_ambiguous(oracle):
    """Comparison frame shows an unrelated texture: no feature may be accepted as a clean match
    with a tight variance."""
    a, _, poses = scene_frames(2, 320, 240, synth.K_VGA * np.array([[0.5], [0.5], [1]], np.float32), seed=1, step=0.03)
    b, _, _ = scene_frames(2, 320, 240, synth.K_VGA * np.array([[0.5], [0.5], [1]], np.float32), seed=77, step=0.03)
    K = (synth.K_VGA * np.array([[0.5], [0.5], [1]], np.float32)).astype(np.float32)
    imgs = np.stack([a[0], b[1]])
    f = init_features(320, 240, 16)
    st, _, cnt = oracle.idepth_update(imgs, poses, K, 1, f["ref_slot"], f["u_ref"], f["mu"], f["var"],
                                      f["dropouts"], f["alive"], oracle.EpiParams.default())
    assert cnt[0] < 0.35 * len(st)
    assert cnt[2] + cnt[3] > 0.5 * len(st)


def test_thread_count_does_not_change_results(oracle):
    imgs, ids, poses = scene_frames(3, 320, 240, synth.K_VGA, seed=4, step=0.02)
    fa, fb = init_features(320, 240, 8), init_features(320, 240, 8)
    ra = oracle.idepth_update(imgs, poses, synth.K_VGA, 2, fa["ref_slot"], fa["u_ref"], fa["mu"], fa["var"],
                              fa["dropouts"], fa["alive"], oracle.EpiParams.default(), nthreads=1)
    rb = oracle.idepth_update(imgs, poses, synth.K_VGA, 2, fb["ref_slot"], fb["u_ref"], fb["mu"], fb["var"],
                              fb["dropouts"], fb["alive"], oracle.EpiParams.default(), nthreads=4)
    assert np.array_equal(ra[0], rb[0]) and np.array_equal(fa["mu"], fb["mu"]) and np.array_equal(fa["var"], fb["var"])


def test_project_features_roundtrip(oracle):
    """Projecting into the feature's own frame is the identity; into another frame it matches a
    float64 pinhole projection."""
    poses = synth.stream_poses(3, step=0.05)
    f = init_features(640, 480, 32)
    rng = np.random.default_rng(2)
    f["mu"] = rng.uniform(0.2, 1.0, len(f["mu"])).astype(np.float32)
    u, mu, var, valid = oracle.project_features(640, 480, poses, synth.K_VGA, 0, f["ref_slot"], f["u_ref"],
                                                f["mu"], f["var"], f["alive"])
    assert np.all(valid == 1) and np.allclose(u, f["u_ref"], atol=1e-3) and np.allclose(mu, f["mu"], rtol=1e-5)
    u, mu, var, valid = oracle.project_features(640, 480, poses, synth.K_VGA, 2, f["ref_slot"], f["u_ref"],
                                                f["mu"], f["var"], f["alive"])
    K = synth.K_VGA.astype(np.float64)
    ray = np.stack([(f["u_ref"][:, 0] - K[0, 2]) / K[0, 0], (f["u_ref"][:, 1] - K[1, 2]) / K[1, 1],
                    np.ones(len(f["mu"]))], axis=1)
    X = ray / f["mu"][:, None] - (poses[2, 4:7] - poses[0, 4:7])[None, :]
    uu = np.stack([K[0, 0] * X[:, 0] / X[:, 2] + K[0, 2], K[1, 1] * X[:, 1] / X[:, 2] + K[1, 2]], axis=1)
    m = valid == 1
    assert m.sum() > 0 and np.max(np.abs(u[m] - uu[m])) < 2e-2
    assert np.max(np.abs(mu[m] - 1.0 / X[m, 2])) < 1e-4


def test_golden_epipolar(oracle):
    gd = np.load(os.path.join(GOLD, "epipolar_small.npz"))
    u = gd["u_ref"]
    N = len(u)
    mu, var = gd["mu0"].copy(), gd["var0"].copy()
    drop, alive, ref = np.zeros(N, np.int32), gd["alive0"].copy(), np.zeros(N, np.int32)
    for cs in (1, 2, 3):
        st, uc, cnt = oracle.idepth_update(gd["imgs"], gd["poses"], gd["K"], cs, ref, u, mu, var, drop, alive,
                                           oracle.EpiParams.default())
        assert np.array_equal(st, gd["status_%d" % cs].astype(np.int32))
        assert np.array_equal(cnt, gd["counters_%d" % cs])
        assert np.array_equal(mu, gd["mu_%d" % cs]) and np.array_equal(var, gd["var_%d" % cs])
        assert np.array_equal(np.nan_to_num(uc, nan=-1), np.nan_to_num(gd["ucmp_%d" % cs], nan=-1))
    pu, pmu, pvar, pvalid = oracle.project_features(160, 120, gd["poses"], gd["K"], 3, ref, u, mu, var, alive)
    assert np.array_equal(pvalid, gd["proj_valid"].astype(np.int32))
    assert np.array_equal(np.nan_to_num(pu, nan=-1), np.nan_to_num(gd["proj_u"], nan=-1))
