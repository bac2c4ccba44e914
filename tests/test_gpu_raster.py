"""GPU parity: triangle validity filters + barycentric rasterisation vs the CPU oracle."""
import numpy as np
import pytest

from flame_ros_b200 import synth
from helpers import gpu_load_graph, small_graph

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _setup(capi, g, x):
    ctx = capi.Context(1, g["W"], g["H"], 2, 16, len(x) + 8, len(g["edges"]) + 8)
    gpu_load_graph(ctx, 0, g)
    ctx.graph_state_set(0, x, None, None)
    ctx.mesh_set(0, g["tris"])
    return ctx


def test_interpolate_unfiltered_parity(capi, oracle):
    g = synth.s_graph("C2")
    x = g["truth"]
    K = synth.K_VGA
    with _setup(capi, g, x) as ctx:
        ctx.set_intrinsics(0, K)
        got, valid = ctx.interpolate(0, None)
    ref = oracle.rasterize_idepth(640, 480, g["pos"], x, g["tris"], None)
    assert np.array_equal(np.isnan(ref), np.isnan(got))
    m = ~np.isnan(ref)
    assert m.sum() > 0.9 * 640 * 480
    assert np.max(np.abs(ref[m] - got[m])) < TOL
    # planar idepth is reproduced by barycentric interpolation
    u, v = np.meshgrid(np.arange(640), np.arange(480))
    plane = 0.5 + 0.2 * u / 640 + 0.1 * v / 480
    assert np.max(np.abs(got[m] - plane[m])) < 1e-4


def test_interpolate_filtered_parity(capi, oracle):
    g = small_graph(24, 18, 320, 240, seed=8)
    rng = np.random.default_rng(2)
    x = g["truth"].copy()
    x[rng.choice(len(x), 20, replace=False)] = rng.uniform(0.001, 3.0, 20).astype(np.float32)  # oblique / far
    x[5] = np.nan
    K = np.array([[260.0, 0, 159.5], [0, 260.0, 119.5], [0, 0, 1]], np.float32)
    fp = capi.default_tri_filter_params()
    fp.oblique_normal_thresh = 1.2
    fp.edge_length_thresh = 0.06
    op = oracle.TriFilterParams.default()
    op.oblique_normal_thresh = 1.2
    op.edge_length_thresh = 0.06
    with _setup(capi, g, x) as ctx:
        ctx.set_intrinsics(0, K)
        got, valid = ctx.interpolate(0, fp)
    rvalid = oracle.triangle_validity(320, 240, K, g["pos"], x, g["tris"], op)
    assert np.array_equal(valid, rvalid)
    assert 0 < rvalid.sum() < len(rvalid)
    ref = oracle.rasterize_idepth(320, 240, g["pos"], x, g["tris"], rvalid)
    assert np.array_equal(np.isnan(ref), np.isnan(got))
    m = ~np.isnan(ref)
    assert np.max(np.abs(ref[m] - got[m])) < TOL
