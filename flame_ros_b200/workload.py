"""The benchmark workload (BASELINE.json configs[1], "C2"): per stream a synthetic 640x480
random-texture image+pose stream, ~5k features = 5k Delaunay-graph vertices, and per frame one
epipolar inverse-depth update of every feature followed by 50 warm-started NLTGV2-L1 primal-dual
iterations on the graph.  Shared by bench.py's GPU arm and its CPU (`--impl reference`) arm so both
see identical inputs and the identical per-step schedule.  Nothing here touches the oracle.
"""
import numpy as np

from . import synth

EPOCH = 5          # comparison frames per poseframe (poseframe_subsample_factor 6:
                   # /root/reference/cfg/flame_nodelet.yaml:6 -> 1 poseframe + 5 updates)
POOL_FRAMES = 12   # two poseframe epochs: [P0 c1..c5 | P1 c1..c5]
MU0, VAR0 = 0.5, 0.25   # fresh-feature prior (inverse metres, variance)

CONFIGS = {
    # name: (W, H, K, detection win, n_vertices, iters)
    "C2": (640, 480, synth.K_VGA, 8, 5000, 50),
    "C3": (752, 480, synth.K_EUROC, 8, 5000, 50),
    "C4": (1280, 720, synth.K_720P, 7, 20000, 100),
    "tiny": (160, 120, np.array([[130.0, 0, 79.5], [0, 130.0, 59.5], [0, 0, 1]], np.float32), 8, 300, 10),
}


class StreamData:
    """Everything one stream needs: frame pool, poses, features (= graph vertices), Delaunay graph."""

    def __init__(self, config, seed):
        W, H, K, win, nv, iters = CONFIGS[config]
        self.W, self.H, self.K, self.iters = W, H, np.asarray(K, np.float32), iters
        sc = synth.Scene(seed, tex_size=1024)
        self.poses = synth.stream_poses(POOL_FRAMES)
        # second epoch restarts from its own poseframe: same relative motion, different viewpoint
        self.frames = np.stack([sc.render(self.K, self.poses[k], W, H)[0] for k in range(POOL_FRAMES)])
        # A poseframe is the current frame flagged is_poseframe (/root/reference/src/flame_nodelet.cc:634):
        # each epoch's poseframe is the LAST frame of the previous epoch (the stream cycles), so a new
        # poseframe never needs its own upload
        half = POOL_FRAMES // 2
        for dst, src in ((half, half - 1), (0, POOL_FRAMES - 1)):
            self.frames[dst] = self.frames[src]
            self.poses[dst] = self.poses[src]
        feats = synth.grid_features(W, H, win, border=8, seed=100 + seed)
        if len(feats) < nv:
            rng = np.random.default_rng(200 + seed)
            have = {(int(x), int(y)) for x, y in feats}
            extra = []
            while len(feats) + len(extra) < nv:
                p = (int(rng.integers(8, W - 8)), int(rng.integers(8, H - 8)))
                if p not in have:
                    have.add(p)
                    extra.append(p)
            feats = np.concatenate([feats, np.array(extra, np.float32).reshape(-1, 2)], axis=0)
        self.u_ref = feats[:nv].astype(np.float32)
        self.tris, self.edges = synth.delaunay(self.u_ref)
        self.alpha, self.beta = synth.edge_weights(self.u_ref, self.edges)
        self.V, self.E = len(self.u_ref), len(self.edges)

    def algorithmic_bytes_per_iter(self):
        """B_iter = 40 E + 64 V."""
        return 40 * self.E + 64 * self.V


def schedule(step):
    """Per-step plan: (new_poseframe, ref_slot, ref_pool_idx, cmp_pool_idx). Slots 0/1 hold the two
    alternating poseframes, slot 2 the current frame."""
    epoch, j = divmod(step, EPOCH)
    half = epoch % 2
    base = half * (POOL_FRAMES // 2)
    return (j == 0, half, base, base + 1 + j)


CMP_SLOT = 2
N_SLOTS = 4   # 2 alternating poseframe slots + 2 alternating current-frame slots (pipelined uploads)


# ---- the full-pipeline workload: flame::Flame::update driven frame by frame ------------------------
UPD_WARMUP = 18        # frames until the graph is populated (3 poseframes at subsample factor 6)
UPD_POSEFRAME_EVERY = 6   # poseframe_subsample_factor, /root/reference/cfg/flame_nodelet.yaml:6
UPD_N_SLOTS = 8        # poseframe ring of 7 + the current frame
UPD_MAX_FEATURES = 8192
UPD_MAX_VERTICES = 8192


def _render_stream(args):
    config, seed, n_frames = args
    W, H, K, win, nv, iters = CONFIGS[config]
    sc = synth.Scene(seed, tex_size=1024)
    poses = synth.stream_poses(n_frames)
    return np.stack([sc.render(np.asarray(K, np.float32), poses[k], W, H)[0] for k in range(n_frames)]), poses


class UpdateStreamData:
    """Image + pose stream for the whole per-frame pipeline (detection .. dense map): the same scene,
    camera and detection window as the hot-path workload, n_frames consecutive frames."""

    def __init__(self, config, seed, n_frames, rendered=None):
        W, H, K, win, nv, iters = CONFIGS[config]
        self.W, self.H, self.K, self.win, self.iters = W, H, np.asarray(K, np.float32), win, iters
        self.frames, self.poses = rendered if rendered is not None else _render_stream((config, seed, n_frames))
        self.n_frames = n_frames


def update_streams(config, seeds, n_frames, workers=None):
    """Render several streams in parallel worker processes (the renderer is single-threaded numpy)."""
    import concurrent.futures as cf
    import os
    jobs = [(config, s, n_frames) for s in seeds]
    workers = workers or min(len(jobs), len(os.sched_getaffinity(0)))
    if workers <= 1 or len(jobs) == 1:
        res = [_render_stream(j) for j in jobs]
    else:
        with cf.ProcessPoolExecutor(workers) as ex:
            res = list(ex.map(_render_stream, jobs))
    return [UpdateStreamData(config, s, n_frames, rendered=r) for s, r in zip(seeds, res)]
