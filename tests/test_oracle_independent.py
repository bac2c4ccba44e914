"""The C oracle against an independent float64 restatement (oracle/independent.py, written from the
equations of SURVEY.md Appendix A / B, not from flame_oracle.c) and against long-horizon known
answers -- the oracle is the only thing the CUDA path is checked against, so it must not be
single-source (VERDICT r1 "parity: green but self-referential")."""
import numpy as np
import pytest

from flame_ros_b200 import synth
from helpers import init_features, run_oracle, scene_frames
from oracle import independent as I


def _ind_state(st):
    return dict(x=st["x"], w=np.stack([st["w1"], st["w2"]], 1), q=np.stack([st["q1"], st["q2"], st["q3"]], 1),
                xb=st["xb"], wb=np.stack([st["w1b"], st["w2b"]], 1))


@pytest.mark.parametrize("config,iters", [("tiny", 10), ("C2", 50), ("C4", 100)])
def test_solver_oracle_matches_independent_float64(oracle, config, iters):
    """fp32 C oracle vs float64 numpy after the same number of iterations: north_star's tolerance
    (vertex idepth L-inf < 1e-4) holds between two independently written implementations."""
    g = synth.s_graph(config)
    ref = run_oracle(oracle, g, iters, nthreads=4)
    ind = I.nltgv2_solve(g["pos"], g["edges"], g["alpha"], g["beta"], g["z"], g["wt"], iters)
    assert np.max(np.abs(ref["x"] - ind["x"])) < 1e-4
    assert np.max(np.abs(ref["xb"] - ind["xb"])) < 1e-4
    assert np.max(np.abs(np.stack([ref["w1"], ref["w2"]], 1) - ind["w"])) < 1e-5
    q = np.stack([ref["q1"], ref["q2"], ref["q3"]], 1)
    # the dual is clipped: away from the clip boundary it agrees tightly, at it both sit on +-1
    assert np.max(np.abs(q - ind["q"])) < 2e-3
    s_ref, d_ref = oracle.nltgv2_costs(g["pos"], g["edges"], g["alpha"], g["beta"], g["z"], g["wt"], ref["x"], ref["w1"], ref["w2"], 0.15)
    s_ind, d_ind = I.costs(g["pos"], g["edges"], g["alpha"], g["beta"], g["z"], g["wt"], ind["x"], ind["w"])
    assert abs(s_ref - s_ind) < 1e-3 * s_ind and abs(d_ref - d_ind) < 1e-3 * d_ind


def test_solver_warm_start_and_other_weights_match_independent(oracle):
    """alpha = beta = 1/|delta| (the other stable rule of SURVEY Appendix E1), non-default steps, a
    warm restart from a mid-run state."""
    g = synth.s_graph("C2")
    d = g["pos"][g["edges"][:, 0]] - g["pos"][g["edges"][:, 1]]
    g["alpha"] = g["beta"] = (1.0 / np.hypot(d[:, 0], d[:, 1])).astype(np.float32)
    p = oracle.NLTGV2Params(0.3, 0.002, 60.0, 0.5, 0.0, 10.0)
    st = run_oracle(oracle, g, 20, params=p)
    ind = I.nltgv2_solve(g["pos"], g["edges"], g["alpha"], g["beta"], g["z"], g["wt"], 20, 0.3, 0.002, 60.0, 0.5)
    assert np.max(np.abs(st["x"] - ind["x"])) < 1e-4
    st2 = run_oracle(oracle, g, 15, params=p, state={k: v.copy() for k, v in st.items()})
    ind2 = I.nltgv2_solve(g["pos"], g["edges"], g["alpha"], g["beta"], g["z"], g["wt"], 15, 0.3, 0.002, 60.0, 0.5, state=_ind_state(st))
    assert np.max(np.abs(st2["x"] - ind2["x"])) < 1e-4


def test_long_horizon_known_answers(oracle):
    """SURVEY Appendix E1 (probed with throw-away code during the survey, numbers recorded there):
    alpha = 1/|delta|, beta = 1 at the reference's default steps on plane + Laplace noise + 5 %
    outliers reaches RMSE 0.0186 vs the true plane at 2000 iterations from 0.16 in the raw data;
    the dual stays unsaturated (4 %), the smoothness cost falls.  Both implementations must
    reproduce that, and each other."""
    g = synth.s_graph("C2")
    raw_rmse = float(np.sqrt(np.mean((g["z"] - g["truth"]) ** 2)))
    assert 0.1 < raw_rmse < 0.25
    hist = []
    ind = I.nltgv2_solve(g["pos"], g["edges"], g["alpha"], g["beta"], g["z"], g["wt"], 2000, history=hist)
    st = run_oracle(oracle, g, 2000, nthreads=4)
    rmse_ind = float(np.sqrt(np.mean((ind["x"] - g["truth"]) ** 2)))
    rmse_ref = float(np.sqrt(np.mean((st["x"] - g["truth"]) ** 2)))
    assert rmse_ind <= 0.02 and rmse_ref <= 0.02, (rmse_ind, rmse_ref)
    assert abs(rmse_ind - rmse_ref) < 1e-3
    assert np.max(np.abs(st["x"] - ind["x"])) < 5e-3       # 2000 iterations of fp32 vs fp64 on a clipped dual
    smooth = np.array([h[0] for h in hist])
    sat = np.array([h[2] for h in hist])
    assert sat[-1] < 0.10                                  # not the saturated / non-convergent regime
    assert smooth[-1] < 0.25 * smooth[49]                  # far below the 50-iteration transient
    # monotone-ish: the running minimum keeps falling, no blow-up
    blocks = smooth[:2000].reshape(20, 100).mean(axis=1)
    assert np.all(np.diff(blocks[2:]) < 0.02 * blocks[2:-1])
    # 50 iterations from a cold start are a transient: not yet close to the plane
    st50 = run_oracle(oracle, g, 50)
    assert float(np.sqrt(np.mean((st50["x"] - g["truth"]) ** 2))) > 3 * rmse_ref


def test_unstable_weights_show_as_dual_saturation_not_nan(oracle):
    """alpha = beta = 1 with pixel-unit deltas violates tau sigma |D|^2 <= 1:
    nothing goes NaN (dual and x are boxed) -- it shows as a saturated dual and a flat cost."""
    g = synth.s_graph("C2")
    g["alpha"] = np.ones_like(g["alpha"])
    hist = []
    ind = I.nltgv2_solve(g["pos"], g["edges"], g["alpha"], g["beta"], g["z"], g["wt"], 300, history=hist)
    st = run_oracle(oracle, g, 300, nthreads=4)
    assert np.all(np.isfinite(ind["x"])) and np.all(np.isfinite(st["x"]))
    assert hist[-1][2] > 0.5
    sat_ref = float((np.abs(np.stack([st["q1"], st["q2"], st["q3"]], 1)) >= 0.999).mean())
    assert sat_ref > 0.5


def test_epipolar_oracle_matches_independent_search(oracle):
    """Geometry, sampling and triangulation of the epipolar search restated independently in float64:
    on features the oracle reports as successes, the implied measurement (un-fused from the prior)
    agrees with the independent dense search, and both agree with the rendered ground truth."""
    W, H, K = 640, 480, synth.K_VGA
    imgs, truth, poses = scene_frames(3, W, H, K, seed=2, step=0.03)
    f = init_features(W, H, 16)
    ep = oracle.EpiParams.default()
    mu0, var0 = f["mu"].copy(), f["var"].copy()
    frames = np.stack([imgs[0], imgs[2]])
    pp = np.stack([poses[0], poses[2]])
    status, u_cmp, counters = oracle.idepth_update(frames, pp, K, 1, f["ref_slot"], f["u_ref"], f["mu"], f["var"],
                                                   f["dropouts"], f["alive"], ep)
    ok = np.nonzero(status == 0)[0]
    assert len(ok) > 400
    err_pix, err_id, err_truth_ind, err_truth_ref = [], [], [], []
    for k in ok[::3]:
        m = I.epipolar_measurement(imgs[0], imgs[2], K, poses[0], poses[2], f["u_ref"][k].astype(np.float64),
                                   float(mu0[k]), float(var0[k]), win=ep.win_size, search_sigma=ep.search_sigma,
                                   idepth_min=ep.idepth_min, idepth_max=ep.idepth_max, max_search_px=ep.max_search_px)
        if m is None:
            continue
        # un-fuse the oracle's update: var1 = v0 vm / (v0 + vm), mu1 = (vm mu0 + v0 mum) / (v0 + vm)
        v0, v1 = float(var0[k]), float(f["var"][k])
        vm = v0 * v1 / (v0 - v1)
        mum = (float(f["mu"][k]) * (v0 + vm) - vm * float(mu0[k])) / v0
        t = float(truth[0][int(f["u_ref"][k][1]), int(f["u_ref"][k][0])])
        err_pix.append(float(np.hypot(*(m["u_cmp"] - u_cmp[k]))))
        err_id.append(abs(m["idepth"] - mum))
        err_truth_ind.append(abs(m["idepth"] - t))
        err_truth_ref.append(abs(mum - t))
        mu_f, var_f = I.gaussian_fuse(float(mu0[k]), v0, mum, vm)
        assert abs(mu_f - float(f["mu"][k])) < 1e-5 and abs(var_f - v1) < 1e-6 * max(v1, 1e-6) + 1e-9
    err_pix, err_id = np.array(err_pix), np.array(err_id)
    assert len(err_pix) > 100
    # matched pixel: same minimum of the same cost curve up to the candidate lattice's phase
    assert np.median(err_pix) < 0.15 and np.mean(err_pix < 0.5) > 0.9
    assert np.median(err_id) < 5e-3
    assert np.median(err_truth_ind) < 0.02 and np.median(err_truth_ref) < 0.02
