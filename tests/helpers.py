"""Shared builders for tests (synthetic inputs + running the oracle / the CUDA path side by side)."""
import numpy as np

from flame_ros_b200 import synth

STATE_KEYS = ("x", "w1", "w2", "xb", "w1b", "w2b", "q1", "q2", "q3")


def small_graph(nx=12, ny=9, W=96, H=72, seed=5, noise=0.02):
    pos = synth.jittered_grid(W, H, nx, ny, 2.0, seed=seed)
    tris, edges = synth.delaunay(pos)
    alpha, beta = synth.edge_weights(pos, edges)
    z, truth = synth.plane_data(pos, W, H, seed=seed + 1, noise=noise)
    return dict(W=W, H=H, pos=pos, tris=tris, edges=edges, alpha=alpha, beta=beta, z=z,
                wt=np.ones(len(z), np.float32), truth=truth)


def run_oracle(O, g, iters, params=None, state=None, nthreads=1):
    p = params or O.NLTGV2Params.default()
    st = state if state is not None else O.new_state(g["z"], len(g["edges"]))
    O.nltgv2_solve(g["pos"], g["edges"], g["alpha"], g["beta"], g["z"], g["wt"], st, p, iters, nthreads)
    return st


def gpu_load_graph(ctx, stream, g, state=None):
    ctx.graph_set(stream, g["pos"], g["edges"], g["alpha"], g["beta"])
    ctx.graph_data_set(stream, g["z"], g["wt"])
    if state is None:
        ctx.graph_state_set(stream)
    else:
        w = np.stack([state["w1"], state["w2"]], axis=1)
        q = np.stack([state["q1"], state["q2"], state["q3"]], axis=1)
        ctx.graph_state_set(stream, state["x"], w, q)


def scene_frames(n_frames=4, W=640, H=480, K=None, seed=0, step=0.01):
    K = synth.K_VGA if K is None else K
    sc = synth.Scene(seed, tex_size=1024)
    poses = synth.stream_poses(n_frames, step=step)
    imgs, ids = [], []
    for k in range(n_frames):
        im, idp = sc.render(K, poses[k], W, H)
        imgs.append(im)
        ids.append(idp)
    return np.stack(imgs), np.stack(ids), poses


def init_features(W, H, win, mu0=0.5, var0=0.25, seed=11):
    u = synth.grid_features(W, H, win, seed=seed)
    N = len(u)
    return dict(u_ref=u, ref_slot=np.zeros(N, np.int32), mu=np.full(N, mu0, np.float32),
                var=np.full(N, var0, np.float32), dropouts=np.zeros(N, np.int32),
                alive=np.ones(N, np.int32))
