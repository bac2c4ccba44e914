"""Generates the committed golden fixtures under tests/golden/.

PARITY UNPINNED: /root/reference ships neither the algorithm's source nor any test vectors (SURVEY.md
sections 0, 4, 8c), so these vectors are produced by THIS repository's CPU oracle (oracle/flame_oracle.c)
on seeded synthetic inputs.  They pin the oracle (and through it the CUDA kernels) against
regressions; they are not outputs of robustrobotics/flame.

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from flame_ros_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402
from helpers import small_graph  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def nltgv2_small():
    g = small_graph()
    out = dict(pos=g["pos"], edges=g["edges"], alpha=g["alpha"], beta=g["beta"], z=g["z"], wt=g["wt"])
    st = O.new_state(g["z"], len(g["edges"]))
    done = 0
    for it in (1, 10, 50):
        O.nltgv2_solve(g["pos"], g["edges"], g["alpha"], g["beta"], g["z"], g["wt"], st,
                       O.NLTGV2Params.default(), it - done)
        done = it
        for k in ("x", "w1", "w2", "q1", "q2", "q3", "xb"):
            out["%s_it%d" % (k, it)] = st[k].copy()
    s, d = O.nltgv2_costs(g["pos"], g["edges"], g["alpha"], g["beta"], g["z"], g["wt"], st["x"], st["w1"],
                          st["w2"], 0.15)
    out["costs_it50"] = np.array([s, d])
    np.savez_compressed(os.path.join(OUT, "nltgv2_small.npz"), **out)


def nltgv2_c2():
    g = synth.s_graph("C2")
    st = O.new_state(g["z"], len(g["edges"]))
    O.nltgv2_solve(g["pos"], g["edges"], g["alpha"], g["beta"], g["z"], g["wt"], st, O.NLTGV2Params.default(), 50)
    np.savez_compressed(os.path.join(OUT, "nltgv2_c2.npz"), pos=g["pos"], edges=g["edges"].astype(np.int16),
                        z=g["z"], x_it50=st["x"], q1_it50=st["q1"].astype(np.float32))


def epipolar_small():
    W, H = 160, 120
    K = np.array([[130.0, 0, 79.5], [0, 130.0, 59.5], [0, 0, 1]], np.float32)
    sc = synth.Scene(2, tex_size=512)
    poses = np.zeros((4, 7), np.float32)
    poses[:, 3] = 1.0
    poses[1, 4:7] = [0.03, 0.001, 0.0]
    a = 0.015
    poses[2] = [0, np.sin(a / 2), 0, np.cos(a / 2), 0.06, -0.002, 0.02]
    poses[3] = [np.sin(a / 2), 0, 0, np.cos(a / 2), 0.09, 0.004, 0.05]
    imgs = np.stack([sc.render(K, p, W, H)[0] for p in poses])
    u = synth.grid_features(W, H, 8, border=4, seed=5)
    N = len(u)
    mu = np.full(N, 0.4, np.float32)
    var = np.full(N, 0.3, np.float32)
    drop = np.zeros(N, np.int32)
    alive = np.ones(N, np.int32)
    alive[::11] = 0
    ref = np.zeros(N, np.int32)
    out = dict(K=K, poses=poses, imgs=imgs, u_ref=u, mu0=mu.copy(), var0=var.copy(), alive0=alive.copy())
    for cs in (1, 2, 3):
        st, uc, cnt = O.idepth_update(imgs, poses, K, cs, ref, u, mu, var, drop, alive, O.EpiParams.default())
        out["status_%d" % cs] = st.astype(np.int8)
        out["mu_%d" % cs] = mu.copy()
        out["var_%d" % cs] = var.copy()
        out["ucmp_%d" % cs] = uc.copy()
        out["counters_%d" % cs] = cnt.copy()
    pu, pmu, pvar, pvalid = O.project_features(W, H, poses, K, 3, ref, u, mu, var, alive)
    out.update(proj_u=pu, proj_mu=pmu, proj_var=pvar, proj_valid=pvalid.astype(np.int8))
    np.savez_compressed(os.path.join(OUT, "epipolar_small.npz"), **out)


def raster_small():
    g = small_graph(16, 12, 160, 120, seed=12)
    rng = np.random.default_rng(3)
    x = g["truth"].copy()
    x[rng.choice(len(x), 10, replace=False)] = rng.uniform(0.002, 2.5, 10).astype(np.float32)
    K = np.array([[130.0, 0, 79.5], [0, 130.0, 59.5], [0, 0, 1]], np.float32)
    fp = O.TriFilterParams.default()
    fp.oblique_normal_thresh = 1.3
    fp.edge_length_thresh = 0.12
    valid = O.triangle_validity(160, 120, K, g["pos"], x, g["tris"], fp)
    m_all = O.rasterize_idepth(160, 120, g["pos"], x, g["tris"], None)
    m_f = O.rasterize_idepth(160, 120, g["pos"], x, g["tris"], valid)
    np.savez_compressed(os.path.join(OUT, "raster_small.npz"), pos=g["pos"], tris=g["tris"], x=x, K=K,
                        valid=valid, map_all=m_all, map_filtered=m_f,
                        filt=np.array([1.3, 0.35, 0.1, 0.12, 0.01], np.float32))


if __name__ == "__main__":
    nltgv2_small()
    nltgv2_c2()
    epipolar_small()
    raster_small()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))
