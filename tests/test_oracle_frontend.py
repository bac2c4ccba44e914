"""CPU tests of the oracle's "next rows" (SURVEY.md section 8f): gradient / pyramid / grid detector and
mesh -> dense inverse-depth interpolation with the display filters."""
import os

import numpy as np

from flame_ros_b200 import synth
from helpers import small_graph

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_gradient_mag_of_a_ramp(oracle):
    img = np.tile(np.arange(0, 128, 2, dtype=np.uint8), (32, 1))  # I = 2x  -> |grad| = 2
    mag = oracle.gradient_mag(img)
    assert np.all(mag[1:-1, 1:-1] == 2.0) and np.all(mag[0] == 0) and np.all(mag[:, 0] == 0)


def test_pyr_down_is_rounded_box_filter(oracle):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(48, 64), dtype=np.uint8)
    out = oracle.pyr_down(img)
    ref = (img[0::2, 0::2].astype(int) + img[0::2, 1::2] + img[1::2, 0::2] + img[1::2, 1::2] + 2) >> 2
    assert np.array_equal(out, ref.astype(np.uint8))


def test_grid_detector_picks_cell_maxima_and_respects_occupancy(oracle):
    rng = np.random.default_rng(1)
    mag = rng.uniform(0, 4, size=(48, 64)).astype(np.float32)
    mag[20, 37] = 50.0   # cell (2, 4) for win 8
    mag[3, 3] = 60.0     # inside the 4-px border: must be ignored
    n, xy, ok = oracle.detect_features(mag, 8, 4, 5.0)
    cells_x = 64 // 8
    c = (20 // 8) * cells_x + 37 // 8
    assert ok[c] == 1 and tuple(xy[c]) == (37.0, 20.0)
    assert n == ok.sum() == 1
    occ = np.zeros(len(ok), np.uint8)
    occ[c] = 1
    n2, _, ok2 = oracle.detect_features(mag, 8, 4, 5.0, occ)
    assert n2 == 0 and ok2.sum() == 0


def test_rasterised_plane_is_exact_and_covers_the_hull(oracle):
    g = small_graph(16, 12, 160, 120, seed=12, noise=0.0)
    m = oracle.rasterize_idepth(160, 120, g["pos"], g["truth"], g["tris"], None)
    u, v = np.meshgrid(np.arange(160), np.arange(120))
    plane = 0.5 + 0.2 * u / 160 + 0.1 * v / 120
    cov = ~np.isnan(m)
    assert cov.mean() > 0.8
    assert np.max(np.abs(m[cov] - plane[cov])) < 1e-5


def test_triangle_filters(oracle):
    g = small_graph(16, 12, 160, 120, seed=12, noise=0.0)
    K = np.array([[130.0, 0, 79.5], [0, 130.0, 59.5], [0, 0, 1]], np.float32)
    x = g["truth"].copy()
    fp = oracle.TriFilterParams.default()
    base = oracle.triangle_validity(160, 120, K, g["pos"], x, g["tris"], fp)
    assert base.mean() > 0.95  # only long hull slivers are rejected by the edge-length filter
    x2 = x.copy()
    x2[40] = 0.001  # far vertex: min_triangle_idepth + idepth-difference filters
    v2 = oracle.triangle_validity(160, 120, K, g["pos"], x2, g["tris"], fp)
    touched = (g["tris"] == 40).any(axis=1)
    assert not v2[touched].any() and np.array_equal(v2[~touched], base[~touched])
    x3 = x.copy()
    x3[41] = np.nan
    v3 = oracle.triangle_validity(160, 120, K, g["pos"], x3, g["tris"], fp)
    assert not v3[(g["tris"] == 41).any(axis=1)].any()
    fp.edge_length_thresh = 0.05  # 8 px: kills most triangles of a 10-px grid
    v4 = oracle.triangle_validity(160, 120, K, g["pos"], x, g["tris"], fp)
    assert v4.sum() < 0.2 * len(v4)


def test_golden_raster(oracle):
    gd = np.load(os.path.join(GOLD, "raster_small.npz"))
    fp = oracle.TriFilterParams.default()
    fp.oblique_normal_thresh, fp.oblique_idepth_diff_factor, fp.oblique_idepth_diff_abs, \
        fp.edge_length_thresh, fp.min_triangle_idepth = [float(v) for v in gd["filt"]]
    valid = oracle.triangle_validity(160, 120, gd["K"], gd["pos"], gd["x"], gd["tris"], fp)
    assert np.array_equal(valid, gd["valid"])
    a = oracle.rasterize_idepth(160, 120, gd["pos"], gd["x"], gd["tris"], None)
    b = oracle.rasterize_idepth(160, 120, gd["pos"], gd["x"], gd["tris"], valid)
    assert np.array_equal(np.nan_to_num(a, nan=-1), np.nan_to_num(gd["map_all"], nan=-1))
    assert np.array_equal(np.nan_to_num(b, nan=-1), np.nan_to_num(gd["map_filtered"], nan=-1))
