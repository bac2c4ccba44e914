"""Seeded synthetic inputs for the FLaME hot path (S-graph and S-stream).

No dataset (TUM / EuRoC) is on disk and there is no network, so every test and bench input is
generated here.  Nothing in this module touches the CPU oracle.
"""
import math
import os

import numpy as np

# cfg/kinect.yaml:7 (VGA Kinect pinhole), EuRoC cam0 pinhole, 720p synthetic.
K_VGA = np.array([[525.0, 0.0, 319.5], [0.0, 525.0, 239.5], [0.0, 0.0, 1.0]], np.float32)
K_EUROC = np.array([[458.654, 0.0, 367.215], [0.0, 457.296, 248.375], [0.0, 0.0, 1.0]], np.float32)
K_720P = np.array([[1050.0, 0.0, 639.5], [0.0, 1050.0, 359.5], [0.0, 0.0, 1.0]], np.float32)


# --------------------------------------------------------------------------- S-graph

def canonical_edges(tris):
    """Unique undirected edges of a triangle list, oriented i<j, sorted by (i,j)."""
    t = np.asarray(tris, np.int64)
    e = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]], axis=0)
    e.sort(axis=1)
    e = np.unique(e, axis=0)
    return e.astype(np.int32)


def delaunay(pos):
    """Delaunay triangulation of pixel positions -> (tris [T,3] int32, edges [E,2] int32).

    scipy is used only to produce *inputs* (fixtures / synthetic graphs); the solver is agnostic to
    where the edge list comes from.
    """
    from scipy.spatial import Delaunay
    tri = Delaunay(np.asarray(pos, np.float64)).simplices.astype(np.int32)
    return tri, canonical_edges(tri)


def jittered_grid(W, H, cells_x, cells_y, jitter, seed, pad_to=None, pad_seed=2):
    """One vertex per cell at the cell centre + U(-jitter, jitter) px, row-major cell order."""
    rng = np.random.default_rng(seed)
    cw, ch = W / cells_x, H / cells_y
    gx, gy = np.meshgrid(np.arange(cells_x), np.arange(cells_y))
    pos = np.stack([(gx.ravel() + 0.5) * cw, (gy.ravel() + 0.5) * ch], axis=1)
    pos += rng.uniform(-jitter, jitter, size=pos.shape)
    if pad_to is not None and pad_to > pos.shape[0]:
        prng = np.random.default_rng(pad_seed)
        extra = prng.uniform([1.0, 1.0], [W - 2.0, H - 2.0], size=(pad_to - pos.shape[0], 2))
        pos = np.concatenate([pos, extra], axis=0)
    return pos.astype(np.float32)


def edge_weights(pos, edges, rule="inv_len"):
    """alpha, beta per edge. Primary rule: alpha = 1/|delta|, beta = 1."""
    d = pos[edges[:, 0]] - pos[edges[:, 1]]
    ln = np.sqrt((d.astype(np.float64) ** 2).sum(axis=1))
    if rule == "inv_len":
        return (1.0 / ln).astype(np.float32), np.ones(len(edges), np.float32)
    if rule == "inv_len_both":
        return (1.0 / ln).astype(np.float32), (1.0 / ln).astype(np.float32)
    if rule == "ones":
        return np.ones(len(edges), np.float32), np.ones(len(edges), np.float32)
    raise ValueError(rule)


def plane_data(pos, W, H, seed=4, noise=0.02, outlier_frac=0.05):
    """z = 0.5 + 0.2 u/W + 0.1 v/H + Laplace(0, noise), outlier_frac of vertices ~ U(0,2)."""
    rng = np.random.default_rng(seed)
    truth = 0.5 + 0.2 * pos[:, 0] / W + 0.1 * pos[:, 1] / H
    z = truth + rng.laplace(0.0, noise, size=truth.shape) if noise > 0 else truth.copy()
    n_out = int(round(outlier_frac * len(z)))
    if n_out:
        idx = rng.choice(len(z), n_out, replace=False)
        z[idx] = rng.uniform(0.0, 2.0, size=n_out)
    return z.astype(np.float32), truth.astype(np.float32)


def s_graph(config="C2"):
    """The S-graph of SURVEY.md section 8(d). Returns a dict of float32/int32 arrays."""
    if config == "C2":
        W, H = 640, 480
        pos = jittered_grid(W, H, 80, 60, 3.0, seed=1, pad_to=5000, pad_seed=2)
        iters = 50
    elif config == "C4":
        W, H = 1280, 720
        pos = jittered_grid(W, H, 200, 100, 2.5, seed=3)
        iters = 100
    elif config == "tiny":
        W, H = 64, 48
        pos = jittered_grid(W, H, 8, 6, 1.5, seed=7)
        iters = 10
    else:
        raise ValueError(config)
    tris, edges = delaunay(pos)
    alpha, beta = edge_weights(pos, edges)
    z, truth = plane_data(pos, W, H)
    return dict(W=W, H=H, pos=pos, tris=tris, edges=edges, alpha=alpha, beta=beta, z=z,
                wt=np.ones(len(z), np.float32), truth=truth, iters=iters)


# --------------------------------------------------------------------------- S-stream

def make_texture(size=2048, seed=0, blur=1.5):
    """Random uint8 texture, Gaussian-blurred and contrast-stretched (gradient >> min_grad_mag)."""
    from scipy.ndimage import gaussian_filter
    rng = np.random.default_rng(seed)
    t = rng.uniform(0.0, 255.0, size=(size, size)).astype(np.float32)
    t = gaussian_filter(t, blur, mode="wrap")
    lo, hi = np.percentile(t, [1.0, 99.0])
    t = np.clip((t - lo) / (hi - lo), 0.0, 1.0) * 255.0
    return t.astype(np.float32)


def _sample_wrap(tex, u, v):
    n = tex.shape[0]
    u0 = np.floor(u)
    v0 = np.floor(v)
    fu = (u - u0).astype(np.float32)
    fv = (v - v0).astype(np.float32)
    iu = np.mod(u0.astype(np.int64), n)
    iv = np.mod(v0.astype(np.int64), n)
    iu1 = np.mod(iu + 1, n)
    iv1 = np.mod(iv + 1, n)
    a = tex[iv, iu] * (1 - fu) + tex[iv, iu1] * fu
    b = tex[iv1, iu] * (1 - fu) + tex[iv1, iu1] * fu
    return a * (1 - fv) + b * fv


def quat_to_R(q):
    x, y, z, w = [float(c) for c in q]
    n = x * x + y * y + z * z + w * w
    s = 2.0 / n
    return np.array([[1 - s * (y * y + z * z), s * (x * y - w * z), s * (x * z + w * y)],
                     [s * (x * y + w * z), 1 - s * (x * x + z * z), s * (y * z - w * x)],
                     [s * (x * z - w * y), s * (y * z + w * x), 1 - s * (x * x + y * y)]], np.float64)


class Scene:
    """Textured plane at 2 m tilted 20 deg about y for world x < x_split, fronto-parallel plane at
    4 m elsewhere (a depth discontinuity), as in SURVEY.md section 8(d) S-stream."""

    def __init__(self, seed=0, tex_size=2048, tilt_deg=20.0, near=2.0, far=4.0, x_split=0.45):
        self.tex_a = make_texture(tex_size, seed, 1.5)
        self.tex_b = make_texture(tex_size, seed + 1000, 1.5)
        th = math.radians(tilt_deg)
        self.n_a = np.array([math.sin(th), 0.0, math.cos(th)])  # plane A normal (world)
        self.d_a = near * math.cos(th)  # n.X = d, passes through (0,0,near)
        self.ax_u = np.array([math.cos(th), 0.0, -math.sin(th)])  # in-plane axes of A
        self.ax_v = np.array([0.0, 1.0, 0.0])
        self.far = far
        self.x_split = x_split
        self.texels_per_m_a = 262.0
        self.texels_per_m_b = 131.0

    def render(self, K, pose, W, H):
        """pose = (qx,qy,qz,qw,tx,ty,tz) camera-in-world, RDF. Returns (uint8 [H,W], idepth [H,W])."""
        K = np.asarray(K, np.float64)
        R = quat_to_R(pose[:4])
        t = np.asarray(pose[4:7], np.float64)
        u, v = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
        rc = np.stack([(u - K[0, 2]) / K[0, 0], (v - K[1, 2]) / K[1, 1], np.ones_like(u)], axis=-1)
        rw = rc @ R.T  # world ray directions (camera z component = 1)
        # plane A
        den_a = rw @ self.n_a
        lam_a = (self.d_a - t @ self.n_a) / np.where(np.abs(den_a) < 1e-12, 1e-12, den_a)
        Xa = t + lam_a[..., None] * rw
        hit_a = (lam_a > 0) & (Xa[..., 0] < self.x_split)
        # plane B: world z = far
        lam_b = (self.far - t[2]) / np.where(np.abs(rw[..., 2]) < 1e-12, 1e-12, rw[..., 2])
        Xb = t + lam_b[..., None] * rw
        ua = (Xa @ self.ax_u) * self.texels_per_m_a + 1000.0
        va = (Xa @ self.ax_v) * self.texels_per_m_a + 1000.0
        ub = Xb[..., 0] * self.texels_per_m_b + 500.0
        vb = Xb[..., 1] * self.texels_per_m_b + 500.0
        ia = _sample_wrap(self.tex_a, ua, va)
        ib = _sample_wrap(self.tex_b, ub, vb)
        img = np.where(hit_a, ia, ib)
        lam = np.where(hit_a, lam_a, lam_b)  # depth along camera z because rc.z = 1
        idepth = (1.0 / lam).astype(np.float32)
        return np.clip(np.rint(img), 0, 255).astype(np.uint8), idepth


def stream_poses(n_frames, step=0.01, wobble=0.001, period=30.0):
    """Identity rotation, +x `step` m/frame with a `wobble` m sinusoidal y (RDF camera-in-world)."""
    p = np.zeros((n_frames, 7), np.float32)
    p[:, 3] = 1.0
    k = np.arange(n_frames)
    p[:, 4] = step * k
    p[:, 5] = wobble * np.sin(2.0 * math.pi * k / period)
    return p


def grid_features(W, H, win, border=8, seed=11):
    """One feature per win x win cell at an integer pixel (emulates the grid detector's output)."""
    rng = np.random.default_rng(seed)
    cx, cy = W // win, H // win
    gx, gy = np.meshgrid(np.arange(cx), np.arange(cy))
    x = gx.ravel() * win + rng.integers(0, win, size=cx * cy)
    y = gy.ravel() * win + rng.integers(0, win, size=cx * cy)
    x = np.clip(x, border, W - 1 - border)
    y = np.clip(y, border, H - 1 - border)
    return np.stack([x, y], axis=1).astype(np.float32)


# --------------------------------------------------------------------------- on-disk datasets (tests)

def write_tum_dataset(root, frames, poses_rdf, K, input_frame="RDF_IN_FLU", with_depth=None, t0=1000.0, dt=1.0 / 30):
    """Write a TUM-RGB-D-shaped dataset: rgb/*.png, optional depth/*.png (uint16, depth*5000),
    `associations.txt` (t tx ty tz qx qy qz qw t_rgb rgb.png [t_d depth.png]) and a camera_info YAML.
    poses_rdf are camera-in-world RDF poses; they are stored in `input_frame` so that the reader's
    conversion (flame_ros_b200/offline.py:tum_pose_to_rdf) recovers them."""
    import cv2
    import yaml
    from . import offline as off
    os.makedirs(os.path.join(root, "rgb"), exist_ok=True)
    if with_depth is not None:
        os.makedirs(os.path.join(root, "depth"), exist_ok=True)
    lines = []
    for k, (img, p) in enumerate(zip(frames, poses_rdf)):
        q, t = np.asarray(p[:4], np.float64), np.asarray(p[4:7], np.float64)
        if input_frame == "RDF":
            qs, ts = q, t
        elif input_frame == "RDF_IN_FLU":
            ci = off.q_inv(off.Q_FLU_TO_RDF)
            qs, ts = off.q_mul(ci, q), off.q_rot(ci, t)
        elif input_frame == "RDF_IN_FRD":
            ci = off.q_inv(off.Q_FRD_TO_RDF)
            qs, ts = off.q_mul(ci, q), off.q_rot(ci, t)
        elif input_frame in ("FLU", "FRD"):
            c = off.Q_FLU_TO_RDF if input_frame == "FLU" else off.Q_FRD_TO_RDF
            ci = off.q_inv(c)
            qs, ts = off.q_mul(off.q_mul(ci, q), c), off.q_rot(ci, t)
        else:
            raise ValueError(input_frame)
        tm = t0 + k * dt
        name = "rgb/%.6f.png" % tm
        cv2.imwrite(os.path.join(root, name), np.stack([img] * 3, axis=-1))
        line = "%.6f %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.6f %s" % (tm, ts[0], ts[1], ts[2], qs[0], qs[1], qs[2], qs[3], tm, name)
        if with_depth is not None:
            dname = "depth/%.6f.png" % tm
            cv2.imwrite(os.path.join(root, dname), np.clip(with_depth[k] * 5000.0, 0, 65535).astype(np.uint16))
            line += " %.6f %s" % (tm, dname)
        lines.append(line)
    open(os.path.join(root, "associations.txt"), "w").write("\n".join(lines) + "\n")
    H, W = frames[0].shape
    K = np.asarray(K, np.float64)
    P = np.concatenate([K, np.zeros((3, 1))], axis=1)
    calib = dict(image_height=int(H), image_width=int(W), camera_name="synthetic",
                 camera_matrix=dict(rows=3, cols=3, data=[float(v) for v in K.ravel()]),
                 distortion_model="plumb_bob", distortion_coefficients=dict(rows=1, cols=5, data=[0.0] * 5),
                 rectification_matrix=dict(rows=3, cols=3, data=[1.0, 0, 0, 0, 1.0, 0, 0, 0, 1.0]),
                 projection_matrix=dict(rows=3, cols=4, data=[float(v) for v in P.ravel()]))
    yaml.safe_dump(calib, open(os.path.join(root, "calib.yaml"), "w"))
    return os.path.join(root, "associations.txt"), os.path.join(root, "calib.yaml")


def write_asl_dataset(root, frames, poses_rdf, K, world_frame="RFU", t0_ns=1403715273262142976, dt_ns=33333333,
                      pose_rate_mult=4):
    """Write an EuRoC/ASL-shaped dataset: cam0/{sensor.yaml,data.csv,data/*.png} and
    state_groundtruth_estimate0/{sensor.yaml,data.csv} (pose rows t,tx,ty,tz,qw,qx,qy,qz at a higher
    rate than the images, identity T_BS for the pose sensor, a non-trivial camera T_BS)."""
    import cv2
    import yaml
    from . import offline as off
    cam, gt = os.path.join(root, "cam0"), os.path.join(root, "state_groundtruth_estimate0")
    os.makedirs(os.path.join(cam, "data"), exist_ok=True)
    os.makedirs(gt, exist_ok=True)
    c = {"RDF": None, "FLU": off.Q_FLU_TO_RDF, "FRD": off.Q_FRD_TO_RDF, "RFU": off.Q_RFU_TO_RDF}[world_frame]
    # camera mounted rotated 90 deg about body z and offset: T_BS (camera in body)
    q_cb = off.q_from_R([[0, -1, 0], [1, 0, 0], [0, 0, 1]])
    t_cb = np.array([0.05, -0.02, 0.01])
    Tc = np.eye(4)
    Tc[:3, :3] = quat_to_R(q_cb)
    Tc[:3, 3] = t_cb
    H, W = frames[0].shape
    yaml.safe_dump(dict(sensor_type="camera", comment="synthetic cam0", T_BS=dict(rows=4, cols=4, data=[float(v) for v in Tc.ravel()]),
                        rate_hz=30, resolution=[int(W), int(H)], camera_model="pinhole",
                        intrinsics=[float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2])],
                        distortion_model="radial-tangential", distortion_coefficients=[0.0, 0.0, 0.0, 0.0]),
                   open(os.path.join(cam, "sensor.yaml"), "w"))
    yaml.safe_dump(dict(sensor_type="visual-inertial", comment="synthetic ground truth",
                        T_BS=dict(rows=4, cols=4, data=[float(v) for v in np.eye(4).ravel()])),
                   open(os.path.join(gt, "sensor.yaml"), "w"))
    rows_img, rows_pose = ["#timestamp [ns],filename"], ["#timestamp,p_x,p_y,p_z,q_w,q_x,q_y,q_z"]
    for k, (img, p) in enumerate(zip(frames, poses_rdf)):
        ts = t0_ns + k * dt_ns
        cv2.imwrite(os.path.join(cam, "data", "%d.png" % ts), img)
        rows_img.append("%d,%d.png" % (ts, ts))
        # camera-in-world (world_frame axes) from the RDF pose, then body-in-world
        q, t = np.asarray(p[:4], np.float64), np.asarray(p[4:7], np.float64)
        if c is not None:
            ci = off.q_inv(c)
            q, t = off.q_mul(ci, q), off.q_rot(ci, t)
        q_bw = off.q_mul(q, off.q_inv(q_cb))
        t_bw = t - off.q_rot(q_bw, t_cb)
        # pose samples: the exact one 1 ms after the image stamp plus distractors between frames
        rows_pose.append("%d,%.9f,%.9f,%.9f,%.9f,%.9f,%.9f,%.9f" % (ts + 1000000, t_bw[0], t_bw[1], t_bw[2], q_bw[3], q_bw[0], q_bw[1], q_bw[2]))
        for j in range(1, pose_rate_mult):
            tj = ts + j * dt_ns // pose_rate_mult
            rows_pose.append("%d,%.9f,%.9f,%.9f,%.9f,%.9f,%.9f,%.9f" % (tj, t_bw[0] + 9.0, t_bw[1], t_bw[2], q_bw[3], q_bw[0], q_bw[1], q_bw[2]))
    open(os.path.join(cam, "data.csv"), "w").write("\n".join(rows_img) + "\n")
    open(os.path.join(gt, "data.csv"), "w").write("\n".join(rows_pose) + "\n")
    return gt, cam
