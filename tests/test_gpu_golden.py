"""GPU path against the COMMITTED golden vectors (tests/golden/*.npz, produced by
tests/golden/make_golden.py from the CPU oracle; parity unpinned -- see that script's header)."""
import os

import numpy as np
import pytest

from flame_ros_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-4


@pytest.mark.parametrize("variant", [1, 2, 3])
def test_golden_nltgv2_small(capi, variant):
    gd = np.load(os.path.join(GOLD, "nltgv2_small.npz"))
    with capi.Context(1, 96, 72, 2, 16, 256, 1024) as ctx:
        ctx.graph_set(0, gd["pos"], gd["edges"], gd["alpha"], gd["beta"])
        ctx.graph_data_set(0, gd["z"], gd["wt"])
        ctx.graph_state_set(0)
        done = 0
        for it in (1, 10, 50):
            ctx.nltgv2_solve(it - done, variant=variant)
            done = it
            st = ctx.graph_state_get(0)
            for k in ("x", "w1", "w2", "q1", "q2", "q3", "xb"):
                assert np.max(np.abs(st[k] - gd["%s_it%d" % (k, it)])) < TOL, (k, it)
            assert np.array_equal(st["x"], gd["x_it%d" % it])
        s, d = ctx.costs(0, 0.15)
        assert np.allclose([s, d], gd["costs_it50"], rtol=1e-6)


@pytest.mark.parametrize("variant", [1, 2, 3])
def test_golden_nltgv2_c2(capi, variant):
    gd = np.load(os.path.join(GOLD, "nltgv2_c2.npz"))
    pos, edges, z = gd["pos"], gd["edges"].astype(np.int32), gd["z"]
    alpha, beta = synth.edge_weights(pos, edges)
    with capi.Context(1, 640, 480, 2, 16, 5000, 15000) as ctx:
        ctx.graph_set(0, pos, edges, alpha, beta)
        ctx.graph_data_set(0, z)
        ctx.graph_state_set(0)
        ctx.nltgv2_solve(50, variant=variant)
        st = ctx.graph_state_get(0)
    assert np.max(np.abs(st["x"] - gd["x_it50"])) < TOL and np.max(np.abs(st["q1"] - gd["q1_it50"])) < TOL
    assert np.array_equal(st["x"], gd["x_it50"])


def test_golden_epipolar(capi):
    gd = np.load(os.path.join(GOLD, "epipolar_small.npz"))
    u = gd["u_ref"]
    N = len(u)
    with capi.Context(1, 160, 120, 4, N, 16, 16) as ctx:
        ctx.set_intrinsics(0, gd["K"])
        for s in range(4):
            ctx.frame_set(0, s, gd["imgs"][s], gd["poses"][s])
        ctx.features_set(0, u, np.zeros(N, np.int32), gd["mu0"], gd["var0"], None, gd["alive0"])
        for cs in (1, 2, 3):
            ctx.idepth_update(cs)
            f = ctx.features_get(0)
            assert np.array_equal(f["status"], gd["status_%d" % cs].astype(np.int32))
            assert np.array_equal(ctx.idepth_counters(0), gd["counters_%d" % cs])
            assert np.max(np.abs(f["mu"] - gd["mu_%d" % cs])) < TOL
            assert np.max(np.abs(f["var"] - gd["var_%d" % cs])) < TOL
            m = np.isfinite(gd["ucmp_%d" % cs][:, 0])
            assert np.array_equal(m, np.isfinite(f["u_cmp"][:, 0]))
            assert np.max(np.abs(f["u_cmp"][m] - gd["ucmp_%d" % cs][m]), initial=0) < TOL
        pu, pmu, pvar, pvalid = ctx.project_features(0, 3)
        assert np.array_equal(pvalid, gd["proj_valid"].astype(np.int32))
        m = pvalid == 1
        assert np.max(np.abs(pu[m] - gd["proj_u"][m])) < TOL and np.max(np.abs(pmu[m] - gd["proj_mu"][m])) < TOL


def test_golden_raster(capi):
    gd = np.load(os.path.join(GOLD, "raster_small.npz"))
    pos, tris, x = gd["pos"], gd["tris"], gd["x"]
    edges = synth.canonical_edges(tris)
    fp = capi.default_tri_filter_params()
    fp.oblique_normal_thresh, fp.oblique_idepth_diff_factor, fp.oblique_idepth_diff_abs, \
        fp.edge_length_thresh, fp.min_triangle_idepth = [float(v) for v in gd["filt"]]
    with capi.Context(1, 160, 120, 2, 16, len(x), len(edges)) as ctx:
        ctx.set_intrinsics(0, gd["K"])
        ctx.graph_set(0, pos, edges, np.ones(len(edges), np.float32), np.ones(len(edges), np.float32))
        ctx.graph_data_set(0, x)
        ctx.graph_state_set(0, x, None, None)
        ctx.mesh_set(0, tris)
        a, _ = ctx.interpolate(0, None)
        b, valid = ctx.interpolate(0, fp)
    assert np.array_equal(valid, gd["valid"])
    for got, ref in ((a, gd["map_all"]), (b, gd["map_filtered"])):
        assert np.array_equal(np.isnan(got), np.isnan(ref))
        m = ~np.isnan(ref)
        assert np.max(np.abs(got[m] - ref[m])) < TOL
