"""CPU check of the GPU triangulation's algorithm (flame_ros_b200/csrc/delaunay_star.h): the same
source the device kernels instantiate with 32 lanes is compiled here for the host with a one-lane
"warp" (tests/cpp/star_sim.cc) and compared with the host triangulator (fb_delaunay) -- triangle
list and edge list must be IDENTICAL (both are canonical: co-circular points fan out from their
smallest index, duplicates keep the smallest index) -- and with Qhull on inputs in general position."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def star():
    src = os.path.join(ROOT, "tests", "cpp", "star_sim.cc")
    out_dir = os.path.join(ROOT, "tests", "cpp", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libstar_sim.so")
    hdr = os.path.join(ROOT, "flame_ros_b200", "csrc", "delaunay_star.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src], check=True)
    lib = C.CDLL(so)

    def run(pts, cell=16):
        pts = np.ascontiguousarray(pts, np.float32)
        n = len(pts)
        tris = np.zeros((2 * n + 8) * 3, np.int32)
        edges = np.zeros((3 * n + 8) * 2, np.int32)
        nt, ne, md = C.c_int32(), C.c_int32(), C.c_int32()
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = lib.star_sim_delaunay(n, vp(pts), cell, vp(tris), C.byref(nt), vp(edges), C.byref(ne), C.byref(md))
        return rc, tris[:3 * nt.value].reshape(-1, 3), edges[:2 * ne.value].reshape(-1, 2), md.value
    return run


def point_sets():
    rng = np.random.default_rng(0)
    yield "uniform-6k", rng.uniform(0, [640, 480], (6000, 2)), 16
    yield "uniform-100", rng.uniform(0, [640, 480], (100, 2)), 16
    yield "tiny-3", rng.uniform(0, [64, 48], (3, 2)), 16
    g = np.stack(np.meshgrid(np.arange(8, 640, 8), np.arange(8, 480, 8)), -1).reshape(-1, 2).astype(np.float32)
    yield "integer-grid", g, 16                       # every cell square is co-circular
    yield "integer-grid-shuffled", g[rng.permutation(len(g))], 16
    yield "integer-random", rng.integers(8, 300, (3000, 2)).astype(np.float32), 16
    yield "integer-dense-duplicates", rng.integers(0, 40, (2000, 2)).astype(np.float32), 8
    yield "clustered", np.concatenate([rng.normal([100, 100], 5, (500, 2)), rng.normal([500, 300], 40, (1500, 2)),
                                       rng.uniform(0, [640, 480], (500, 2))]), 16
    yield "thin-strip", rng.uniform(0, [640, 3], (800, 2)), 16   # hull slivers, long thin triangles
    pc = np.array([(25, 0), (0, 25), (-25, 0), (0, -25), (15, 20), (20, 15), (-15, 20), (-20, 15), (15, -20),
                   (20, -15), (-15, -20), (-20, -15), (7, 24), (24, 7), (-7, 24), (-24, 7), (7, -24), (24, -7),
                   (-7, -24), (-24, -7)], np.float32) + 100
    for k in range(4):
        yield "cocircular-20-perm%d" % k, pc[rng.permutation(len(pc))], 16
    yield "cocircular-20+centre", np.concatenate([pc, [[100, 100]]]).astype(np.float32), 16
    yield "collinear+1", np.concatenate([np.stack([np.arange(25) * 3.0, np.full(25, 7.0)], 1), [[20., 30.]]]), 16
    yield "C4-20k", rng.uniform(0, [1280, 720], (20000, 2)), 16
    for k in range(40):
        n = int(rng.integers(3, 60))
        yield "small-lattice-%d" % k, rng.integers(0, 8, (n, 2)).astype(np.float32), 4


@pytest.mark.parametrize("name,pts,cell", list(point_sets()), ids=[n for n, _, _ in point_sets()])
def test_star_algorithm_equals_host_triangulator(capi, star, name, pts, cell):
    pts = np.ascontiguousarray(pts, np.float32)
    rc, tris, edges, maxdeg = star(pts, cell)
    try:
        ht, he = capi.delaunay(pts)
    except capi.FlameError:
        ht, he = np.zeros((0, 3), np.int32), np.zeros((0, 2), np.int32)
    if len(ht) == 0:   # degenerate (all collinear): no mesh on either side
        assert len(tris) == 0
        return
    assert rc == 0 and maxdeg <= 32
    assert np.array_equal(tris, ht), name
    assert np.array_equal(edges, he), name


def test_star_algorithm_matches_qhull(star):
    scipy_spatial = pytest.importorskip("scipy.spatial")
    rng = np.random.default_rng(3)
    pts = (np.round(rng.uniform(0, [640, 480], (4000, 2)) * 64) / 64).astype(np.float32)
    rc, tris, edges, _ = star(pts)
    assert rc == 0
    q = scipy_spatial.Delaunay(pts.astype(np.float64)).simplices
    canon = lambda t: {tuple(sorted(map(int, r))) for r in t}
    assert canon(tris) == canon(q)


def test_degree_overflow_is_reported(star):
    """A vertex with more than 32 Delaunay neighbours cannot be held: reported, never silently wrong."""
    pts = np.concatenate([np.stack([np.arange(50) * 3.0, np.full(50, 7.0)], 1), [[70., 60.]]]).astype(np.float32)
    rc, _, _, _ = star(pts)
    assert rc == 1   # DS_E_DEGREE
