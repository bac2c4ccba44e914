"""CPU tests of the host-side triangulator behind the `triangulate` stage (fb_delaunay in the C-ABI;
no GPU needed).  scipy's Qhull Delaunay is the checker: on points in general position the Delaunay
triangulation is unique, so the edge sets must be identical; on degenerate (co-circular / collinear)
inputs any valid triangulation is accepted and checked through its properties."""
import numpy as np
import pytest
from scipy.spatial import Delaunay

from flame_ros_b200 import synth


def edge_set(tris):
    return {tuple(e) for e in synth.canonical_edges(tris)}


def signed_areas(pts, tris):
    a, b, c = pts[tris[:, 0]], pts[tris[:, 1]], pts[tris[:, 2]]
    return 0.5 * ((b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0]))


def hull_area(pts):
    from scipy.spatial import ConvexHull
    return ConvexHull(pts).volume


def empty_circle_violations(pts, tris, tol=1e-7):
    """Number of (triangle, vertex) pairs with the vertex strictly inside the circumcircle (float64)."""
    P = pts.astype(np.float64)
    bad = 0
    for t in tris:
        a, b, c = P[t[0]], P[t[1]], P[t[2]]
        d = 2 * (a[0] * (b[1] - c[1]) + b[0] * (c[1] - a[1]) + c[0] * (a[1] - b[1]))
        ux = ((a @ a) * (b[1] - c[1]) + (b @ b) * (c[1] - a[1]) + (c @ c) * (a[1] - b[1])) / d
        uy = ((a @ a) * (c[0] - b[0]) + (b @ b) * (a[0] - c[0]) + (c @ c) * (b[0] - a[0])) / d
        r2 = (a[0] - ux) ** 2 + (a[1] - uy) ** 2
        d2 = (P[:, 0] - ux) ** 2 + (P[:, 1] - uy) ** 2
        bad += int(np.sum(d2 < r2 * (1 - tol)))
    return bad


def snap(p):
    return (np.round(np.asarray(p, np.float32) * 64) / 64).astype(np.float32)


@pytest.mark.parametrize("n,seed", [(10, 0), (300, 1), (5000, 2)])
def test_random_points_match_qhull(capi, n, seed):
    rng = np.random.default_rng(seed)
    pts = snap(rng.uniform([0, 0], [639, 479], (n, 2)))
    tris, edges = capi.delaunay(pts)
    ref = Delaunay(pts.astype(np.float64)).simplices
    assert len(tris) == len(ref)
    assert edge_set(tris) == edge_set(ref)
    assert np.all(signed_areas(pts, tris) > 0)
    # canonical edge list: i<j, strictly sorted by (i,j)
    assert np.all(edges[:, 0] < edges[:, 1])
    key = edges[:, 0].astype(np.int64) * (n + 1) + edges[:, 1]
    assert np.all(np.diff(key) > 0)
    assert {tuple(e) for e in edges} == edge_set(tris)


def test_bench_graph_matches_qhull(capi):
    pts = snap(synth.jittered_grid(640, 480, 80, 60, 3.0, seed=1))
    tris, edges = capi.delaunay(pts)
    assert edge_set(tris) == edge_set(Delaunay(pts.astype(np.float64)).simplices)


def test_cocircular_grid_is_a_valid_delaunay_triangulation(capi):
    """Detections on a regular lattice: every cell is co-circular, any diagonal is acceptable."""
    gx, gy = np.meshgrid(np.arange(0, 96, 8.0), np.arange(0, 72, 8.0))
    pts = np.stack([gx.ravel(), gy.ravel()], axis=1).astype(np.float32)
    tris, edges = capi.delaunay(pts)
    n = len(pts)
    assert len(tris) == 2 * (12 - 1) * (9 - 1)
    areas = signed_areas(pts, tris)
    assert np.all(areas > 0) and abs(areas.sum() - 88.0 * 64.0) < 1e-3
    assert empty_circle_violations(pts, tris) == 0
    # Euler: V - E + T = 1 for a triangulated disk
    assert n - len(edges) + len(tris) == 1


def test_integer_pixel_features_like_the_detector(capi):
    pts = synth.grid_features(640, 480, 8)
    tris, edges = capi.delaunay(pts)
    areas = signed_areas(pts, tris)
    assert np.all(areas > 0) and abs(areas.sum() - hull_area(pts.astype(np.float64))) < 1e-2
    assert len(np.unique(pts, axis=0)) - len(edges) + len(tris) == 1   # clipped border cells collide
    rng = np.random.default_rng(0)
    sub = rng.choice(len(tris), 300, replace=False)
    assert empty_circle_violations(pts, tris[sub]) == 0


def test_collinear_prefix_and_duplicates(capi):
    line = np.stack([np.arange(10) * 5.0, np.full(10, 7.0)], axis=1)
    pts = np.concatenate([line, [[20.0, 30.0], [20.0, 30.0], [3.0, 3.0], [0.0, 7.0]]]).astype(np.float32)
    tris, edges = capi.delaunay(pts)
    used = set(tris.ravel().tolist())
    assert 11 not in used or 10 not in used      # one of the duplicate pair is unreferenced
    assert 13 not in used or 0 not in used       # (0,7) duplicates line[0]
    uniq = np.unique(pts, axis=0)
    areas = signed_areas(pts, tris)
    assert np.all(areas > 0) and abs(areas.sum() - hull_area(uniq.astype(np.float64))) < 1e-3
    assert empty_circle_violations(pts, tris) == 0


def test_degenerate_inputs_are_rejected(capi):
    with pytest.raises(capi.FlameError):
        capi.delaunay(np.zeros((2, 2), np.float32))
    with pytest.raises(capi.FlameError):
        capi.delaunay(np.stack([np.arange(6.0), 2 * np.arange(6.0)], axis=1).astype(np.float32))


def test_deterministic_and_independent_of_earlier_calls(capi):
    """The insertion order is a fixed pseudo-random permutation (no clock, no global state): the same
    points give the same triangles in the same order, whatever was triangulated before."""
    rng = np.random.default_rng(3)
    pts = snap(rng.uniform([0, 0], [639, 479], (2000, 2)))
    t1, e1 = capi.delaunay(pts)
    capi.delaunay(snap(rng.uniform([0, 0], [100, 100], (50, 2))))
    t2, e2 = capi.delaunay(pts)
    assert np.array_equal(t1, t2) and np.array_equal(e1, e2)


def test_c4_sized_grid_matches_qhull(capi):
    pts = snap(synth.jittered_grid(1280, 720, 200, 100, 2.5, seed=3))
    tris, edges = capi.delaunay(pts)
    assert edge_set(tris) == edge_set(Delaunay(pts.astype(np.float64)).simplices)


def test_coordinates_outside_the_lattice_range_are_rejected(capi):
    pts = np.array([[0, 0], [1e8, 0], [0, 1e8], [5, 5]], np.float32)
    with pytest.raises(capi.FlameError):
        capi.delaunay(pts)
