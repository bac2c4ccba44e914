"""CPU tests of the oracle's NLTGV2-L1 restatement: known-answer tests, an independent numpy
restatement, thread-count invariance and the committed golden vectors.  (PARITY UNPINNED: the
reference ships no vectors for this path -- SURVEY.md section 8c.)"""
import os

import numpy as np

from flame_ros_b200 import synth
from helpers import STATE_KEYS, run_oracle, small_graph

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def numpy_iteration(g, st, p, iters):
    """Independent vectorised float32 restatement of SURVEY.md Appendix A (scatter via np.add.at,
    so summation order differs from the oracle's CSR order: compare with a tolerance)."""
    f = np.float32
    i, j = g["edges"][:, 0], g["edges"][:, 1]
    dx = (g["pos"][i, 0] - g["pos"][j, 0]).astype(f)
    dy = (g["pos"][i, 1] - g["pos"][j, 1]).astype(f)
    a, b = g["alpha"], g["beta"]
    x, w1, w2 = st["x"].copy(), st["w1"].copy(), st["w2"].copy()
    xb, w1b, w2b = st["xb"].copy(), st["w1b"].copy(), st["w2b"].copy()
    q1, q2, q3 = st["q1"].copy(), st["q2"].copy(), st["q3"].copy()
    sig, tau, th, lam = f(p.step_q), f(p.step_x), f(p.theta), f(p.data_factor)
    for _ in range(iters):
        k1 = a * (xb[i] - xb[j] - dx * w1b[i] - dy * w2b[i])
        k2 = b * (w1b[i] - w1b[j])
        k3 = b * (w2b[i] - w2b[j])
        q1 = np.clip(q1 + sig * k1, -1, 1).astype(f)
        q2 = np.clip(q2 + sig * k2, -1, 1).astype(f)
        q3 = np.clip(q3 + sig * k3, -1, 1).astype(f)
        gx = np.zeros_like(x, dtype=np.float64)
        g1 = np.zeros_like(x, dtype=np.float64)
        g2 = np.zeros_like(x, dtype=np.float64)
        np.add.at(gx, i, a * q1)
        np.add.at(gx, j, -a * q1)
        np.add.at(g1, i, -a * dx * q1 + b * q2)
        np.add.at(g1, j, -b * q2)
        np.add.at(g2, i, -a * dy * q1 + b * q3)
        np.add.at(g2, j, -b * q3)
        xo, w1o, w2o = x, w1, w2
        xp = (x - tau * gx).astype(f)
        w1 = (w1 - tau * g1).astype(f)
        w2 = (w2 - tau * g2).astype(f)
        thr = tau * lam * g["wt"]
        d = xp - g["z"]
        x = np.where(d > thr, xp - thr, np.where(d < -thr, xp + thr, g["z"])).astype(f)
        x = np.clip(x, p.x_min, p.x_max).astype(f)
        xb = (x + th * (x - xo)).astype(f)
        w1b = (w1 + th * (w1 - w1o)).astype(f)
        w2b = (w2 + th * (w2 - w2o)).astype(f)
    return dict(x=x, w1=w1, w2=w2, xb=xb, w1b=w1b, w2b=w2b, q1=q1, q2=q2, q3=q3)


def test_matches_independent_numpy_restatement(oracle):
    g = small_graph()
    p = oracle.NLTGV2Params.default()
    st0 = oracle.new_state(g["z"], len(g["edges"]))
    ref = numpy_iteration(g, st0, p, 30)
    got = run_oracle(oracle, g, 30)
    for k in STATE_KEYS:
        assert np.max(np.abs(ref[k] - got[k])) < 2e-5, k


def test_exact_plane_is_a_fixed_point(oracle):
    g = small_graph(noise=0.0)
    g["z"] = g["truth"].copy()
    V, E = len(g["z"]), len(g["edges"])
    st = oracle.new_state(g["z"], E)
    st["w1"][:] = st["w1b"][:] = 0.2 / g["W"]
    st["w2"][:] = st["w2b"][:] = 0.1 / g["H"]
    run_oracle(oracle, g, 100, state=st)
    assert np.max(np.abs(st["x"] - g["z"])) < 1e-5
    s, d = oracle.nltgv2_costs(g["pos"], g["edges"], g["alpha"], g["beta"], g["z"], g["wt"], st["x"],
                               st["w1"], st["w2"], 0.15)
    assert s < 1e-2 and d < 1e-3


def test_l1_data_term_rejects_an_outlier(oracle):
    """Planar data with ONE gross outlier, warm-started on the plane: the outlier vertex is pulled
    back toward the plane (L1 data term), inliers stay on their data."""
    g = small_graph(noise=0.0)
    g["z"] = g["truth"].copy()
    k = len(g["z"]) // 2 + 3
    g["z"][k] += 1.0
    st = oracle.new_state(g["z"], len(g["edges"]))
    st["w1"][:] = st["w1b"][:] = 0.2 / g["W"]
    st["w2"][:] = st["w2b"][:] = 0.1 / g["H"]
    run_oracle(oracle, g, 3000, state=st)
    inl = np.ones(len(g["z"]), bool)
    inl[k] = False
    assert abs(st["x"][k] - g["truth"][k]) < 0.25 * 1.0          # pulled most of the way back
    assert np.max(np.abs(st["x"][inl] - g["truth"][inl])) < 0.02  # inliers barely move


def test_dual_saturates_and_everything_stays_boxed(oracle):
    g = small_graph()
    rng = np.random.default_rng(0)
    g["z"] = rng.uniform(-5.0, 15.0, len(g["z"])).astype(np.float32)
    st = run_oracle(oracle, g, 50)
    assert np.isclose(np.abs(st["q1"]).max(), 1.0)
    for k in ("q1", "q2", "q3"):
        assert np.all(np.abs(st[k]) <= 1.0)
    assert st["x"].min() >= 0.0 and st["x"].max() <= 10.0
    assert all(np.all(np.isfinite(st[k])) for k in STATE_KEYS)


def test_smoothness_cost_decreases_from_cold_start(oracle):
    g = synth.s_graph("C2")
    args = (g["pos"], g["edges"], g["alpha"], g["beta"], g["z"], g["wt"])
    st = oracle.new_state(g["z"], len(g["edges"]))
    s0, _ = oracle.nltgv2_costs(*args, st["x"], st["w1"], st["w2"], 0.15)
    run_oracle(oracle, g, 50, state=st)
    s50, _ = oracle.nltgv2_costs(*args, st["x"], st["w1"], st["w2"], 0.15)
    run_oracle(oracle, g, 450, state=st)
    s500, _ = oracle.nltgv2_costs(*args, st["x"], st["w1"], st["w2"], 0.15)
    assert s500 < s50 < s0
    assert np.sqrt(np.mean((st["x"] - g["truth"]) ** 2)) < np.sqrt(np.mean((g["z"] - g["truth"]) ** 2))


def test_thread_count_does_not_change_results(oracle):
    g = synth.s_graph("C2")
    a = run_oracle(oracle, g, 20, nthreads=1)
    b = run_oracle(oracle, g, 20, nthreads=4)
    assert all(np.array_equal(a[k], b[k]) for k in STATE_KEYS)


def test_split_runs_equal_one_run(oracle):
    g = small_graph()
    a = run_oracle(oracle, g, 50)
    b = run_oracle(oracle, g, 20)
    run_oracle(oracle, g, 30, state=b)
    assert all(np.array_equal(a[k], b[k]) for k in STATE_KEYS)


def test_costs_match_float64_numpy(oracle):
    g = small_graph()
    st = run_oracle(oracle, g, 10)
    s, d = oracle.nltgv2_costs(g["pos"], g["edges"], g["alpha"], g["beta"], g["z"], g["wt"], st["x"],
                               st["w1"], st["w2"], 0.15)
    i, j = g["edges"][:, 0], g["edges"][:, 1]
    P = g["pos"].astype(np.float64)
    x, w1, w2 = (st[k].astype(np.float64) for k in ("x", "w1", "w2"))
    k1 = g["alpha"] * (x[i] - x[j] - (P[i, 0] - P[j, 0]) * w1[i] - (P[i, 1] - P[j, 1]) * w2[i])
    k2 = g["beta"] * (w1[i] - w1[j])
    k3 = g["beta"] * (w2[i] - w2[j])
    s64 = np.sum(np.abs(k1) + np.abs(k2) + np.abs(k3))
    d64 = np.sum(0.15 * g["wt"] * np.abs(x - g["z"]))
    assert abs(s - s64) < 1e-4 * s64 and abs(d - d64) < 1e-4 * d64


def test_empty_graph_and_isolated_vertex(oracle):
    z = np.array([0.7], np.float32)
    st = oracle.new_state(z, 0)
    oracle.nltgv2_solve(np.zeros((1, 2), np.float32), np.zeros((0, 2), np.int32), np.zeros(0, np.float32),
                        np.zeros(0, np.float32), z, np.ones(1, np.float32), st, oracle.NLTGV2Params.default(), 5)
    assert st["x"][0] == np.float32(0.7)


def test_golden_small(oracle):
    gd = np.load(os.path.join(GOLD, "nltgv2_small.npz"))
    g = {k: gd[k] for k in ("pos", "edges", "alpha", "beta", "z", "wt")}
    st = oracle.new_state(g["z"], len(g["edges"]))
    done = 0
    for it in (1, 10, 50):
        run_oracle(oracle, g, it - done, state=st)
        done = it
        for k in ("x", "w1", "w2", "q1", "q2", "q3", "xb"):
            assert np.array_equal(st[k], gd["%s_it%d" % (k, it)]), (k, it)
    s, d = oracle.nltgv2_costs(g["pos"], g["edges"], g["alpha"], g["beta"], g["z"], g["wt"], st["x"],
                               st["w1"], st["w2"], 0.15)
    assert np.allclose([s, d], gd["costs_it50"], rtol=1e-12)


def test_golden_c2(oracle):
    gd = np.load(os.path.join(GOLD, "nltgv2_c2.npz"))
    pos, edges, z = gd["pos"], gd["edges"].astype(np.int32), gd["z"]
    alpha, beta = synth.edge_weights(pos, edges)
    g = dict(pos=pos, edges=edges, alpha=alpha, beta=beta, z=z, wt=np.ones(len(z), np.float32))
    st = run_oracle(oracle, g, 50)
    assert np.array_equal(st["x"], gd["x_it50"]) and np.array_equal(st["q1"], gd["q1_it50"])
