"""ROS-free offline frontends: the TUM RGB-D and EuRoC/ASL file players of
the reference, re-created around libflame_b200 without ROS, PCL or catkin.

Mirrors, for the data path only:
  * TUM association file `t tx ty tz qx qy qz qw t_rgb rgb.png [t_d depth.png]`, depth/5000 and the five
    pose-frame conventions (/root/reference/src/ros_sensor_streams/tum_rgbd_offline_stream.cc:124-300,
    tum_rgbd_offline_stream.h:53-80);
  * ASL `sensor.yaml` + `data.csv` (ns timestamps; pose rows `t,tx,ty,tz,qw,qx,qy,qz`), the T_BS extrinsics
    chain, greedy nearest-timestamp association with max_diff 0.02 s and the world-frame conventions
    (/root/reference/src/ros_sensor_streams/asl_rgbd_offline_stream.cc:135-275,
    src/dataset_utils/asl/types.h:60-78, src/dataset_utils/utils.h:50-93);
  * the offline loop get() -> gray -> is_poseframe = id % poseframe_subsample_factor == 0 -> update()
    (/root/reference/src/flame_offline_tum.cc:417-436,565-601), WITHOUT the 30 Hz ros::Rate sleep;
  * the mesh output contract of publishDepthMesh (/root/reference/src/utils.cc:163-237): points
    Kinv*(u,v,1)/idepth, NaN for invalid vertices, texture uv, reversed-winding faces of valid triangles,
    nothing emitted when no triangle is valid.
Images are decoded with OpenCV's Python module (I/O glue, like the reference's cv::imread).
"""
import argparse
import json
import os
import time

import numpy as np

POSE_FRAMES_TUM = ("RDF", "FLU", "FRD", "RDF_IN_FLU", "RDF_IN_FRD")
WORLD_FRAMES_ASL = ("RDF", "FLU", "FRD", "RFU")


# ------------------------------------------------------------------------------------- quaternions
def q_mul(a, b):
    """Hamilton product, (x, y, z, w) order."""
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw,
                     aw * bw - ax * bx - ay * by - az * bz], np.float64)


def q_inv(q):
    q = np.asarray(q, np.float64)
    return np.array([-q[0], -q[1], -q[2], q[3]]) / np.dot(q, q)


def q_rot(q, v):
    """Rotate vector v by unit quaternion q (x, y, z, w)."""
    qv = np.array([v[0], v[1], v[2], 0.0])
    return q_mul(q_mul(q, qv), q_inv(q))[:3]


def q_from_R(R):
    R = np.asarray(R, np.float64)
    tr = np.trace(R)
    if tr > 0:
        s = np.sqrt(tr + 1.0) * 2
        q = [(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s]
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = [0.0, 0.0, 0.0, 0.0]
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        q[3] = (R[k, j] - R[j, k]) / s
    q = np.array(q, np.float64)
    return q / np.linalg.norm(q)


# Eigen's constructor order is (w, x, y, z): Quaternionf(-0.5, -0.5, 0.5, -0.5)
# (/root/reference/src/ros_sensor_streams/tum_rgbd_offline_stream.cc:172-178)
Q_FLU_TO_RDF = np.array([-0.5, 0.5, -0.5, -0.5])  # (x, y, z, w)
Q_FRD_TO_RDF = q_from_R([[0, 1, 0], [0, 0, 1], [1, 0, 0]])
Q_RFU_TO_RDF = q_from_R([[1, 0, 0], [0, 0, -1], [0, 1, 0]])


def tum_pose_to_rdf(q, t, frame):
    """The five input-frame conventions of TUMRGBDOfflineStream::get."""
    q = np.asarray(q, np.float64)
    q = q / np.linalg.norm(q)
    t = np.asarray(t, np.float64)
    if frame == "RDF":
        return q, t
    if frame == "FLU":
        return q_mul(q_mul(Q_FLU_TO_RDF, q), q_inv(Q_FLU_TO_RDF)), q_rot(Q_FLU_TO_RDF, t)
    if frame == "FRD":
        return q_mul(q_mul(Q_FRD_TO_RDF, q), q_inv(Q_FRD_TO_RDF)), q_rot(Q_FRD_TO_RDF, t)
    if frame == "RDF_IN_FLU":
        return q_mul(Q_FLU_TO_RDF, q), q_rot(Q_FLU_TO_RDF, t)
    if frame == "RDF_IN_FRD":
        return q_mul(Q_FRD_TO_RDF, q), q_rot(Q_FRD_TO_RDF, t)
    raise ValueError("unknown input frame %r" % frame)


def asl_world_to_rdf(q, t, world_frame):
    """World-frame conventions of ASLRGBDOfflineStream::get (local RDF camera in the given world)."""
    if world_frame == "RDF":
        return q, t
    c = {"FLU": Q_FLU_TO_RDF, "FRD": Q_FRD_TO_RDF, "RFU": Q_RFU_TO_RDF}.get(world_frame)
    if c is None:
        raise ValueError("unknown world frame %r" % world_frame)
    return q_mul(c, q), q_rot(c, t)


# ------------------------------------------------------------------------------------- calibration
def load_camera_info(path):
    """ROS camera_info YAML (cfg/kinect.yaml). Intrinsics come from P, not K
    (/root/reference/src/ros_sensor_streams/tum_rgbd_offline_stream.cc:92-103)."""
    import yaml
    d = yaml.safe_load(open(path))
    P = np.array(d["projection_matrix"]["data"], np.float64).reshape(3, 4)
    D = np.array(d["distortion_coefficients"]["data"], np.float64)
    Kraw = np.array(d["camera_matrix"]["data"], np.float64).reshape(3, 3)
    return dict(width=int(d["image_width"]), height=int(d["image_height"]), K=P[:, :3].astype(np.float32),
                K_raw=Kraw, D=D)


def _rectify(img, K_raw, D, K_new=None):
    if D is None or not np.any(np.abs(D) > 0):
        return img
    import cv2
    return cv2.undistort(img, K_raw, D, None, K_new)


def _to_gray(img):
    if img.ndim == 2:
        return np.ascontiguousarray(img, np.uint8)
    import cv2
    return cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)  # the frontends convert before update() (tum:572-573)


# ------------------------------------------------------------------------------------- TUM
def parse_tum_line(line):
    """-> (time, trans[3], quat_xyzw[4], rgb_file, depth_file or None); time = rgb time (as the reference)."""
    tok = line.split()
    if len(tok) < 10:
        raise ValueError("TUM line needs >= 10 tokens: %r" % line)
    t = [float(v) for v in tok[1:4]]
    q = [float(v) for v in tok[4:8]]
    depth = tok[11] if len(tok) >= 12 else None
    return float(tok[8]), t, q, tok[9], depth


class TUMStream:
    def __init__(self, input_file, calib, input_frame="RDF_IN_FLU", depth_scale_factor=5000.0):
        if input_frame not in POSE_FRAMES_TUM:
            raise ValueError("unknown input frame %r" % input_frame)
        self.lines = [l for l in open(input_file).read().splitlines() if l.strip() and not l.startswith("#")]
        self.base = os.path.dirname(os.path.abspath(input_file))
        self.calib = calib if isinstance(calib, dict) else load_camera_info(calib)
        self.frame, self.depth_scale, self.idx = input_frame, depth_scale_factor, 0
        self.width, self.height, self.K = self.calib["width"], self.calib["height"], self.calib["K"]

    def empty(self):
        return self.idx >= len(self.lines)

    def get(self):
        """-> (id, time, gray uint8 [H,W], depth float32 [H,W] (0 = none), quat_xyzw, trans) in RDF."""
        import cv2
        tm, t, q, rgb_file, depth_file = parse_tum_line(self.lines[self.idx])
        quat, trans = tum_pose_to_rdf(q, t, self.frame)
        rgb = cv2.imread(os.path.join(self.base, rgb_file), cv2.IMREAD_UNCHANGED)
        if rgb is None:
            raise IOError("cannot read %s" % rgb_file)
        rgb = _rectify(rgb, self.calib["K_raw"], self.calib["D"])
        if depth_file is None:
            depth = np.zeros(rgb.shape[:2], np.float32)  # dummy depthmap, as the reference
        else:
            raw = cv2.imread(os.path.join(self.base, depth_file), cv2.IMREAD_ANYDEPTH)
            depth = raw.astype(np.float32) / np.float32(self.depth_scale)
        out = (self.idx, tm, _to_gray(rgb), depth, quat, trans)
        self.idx += 1
        return out


# ------------------------------------------------------------------------------------- ASL
def associate(ta, tb, max_diff=0.02):
    """Greedy nearest-timestamp association without repeats (dataset_utils::associate): all pairs closer
    than max_diff sorted by distance, taken greedily; returns sorted index lists into a and b."""
    ta, tb = np.asarray(ta, np.float64), np.asarray(tb, np.float64)
    order_b = np.argsort(tb)
    sb = tb[order_b]
    cand = []
    for ia, t in enumerate(ta):
        lo, hi = np.searchsorted(sb, t - max_diff, "left"), np.searchsorted(sb, t + max_diff, "right")
        for k in range(lo, hi):
            d = np.float32(abs(t - sb[k]))  # the reference compares float differences
            if d < np.float32(max_diff):
                cand.append((float(d), ia, int(order_b[k])))
    cand.sort(key=lambda c: c[0])  # stable, like the reference's sort on distance only
    used_a, used_b, ia_out, ib_out = set(), set(), [], []
    for _, ia, ib in cand:
        if ia not in used_a and ib not in used_b:
            used_a.add(ia)
            used_b.add(ib)
            ia_out.append(ia)
            ib_out.append(ib)
    return sorted(ia_out), sorted(ib_out)


def _read_asl_sensor(path):
    import yaml
    meta = yaml.safe_load(open(os.path.join(path, "sensor.yaml")))
    rows = [l.strip() for l in open(os.path.join(path, "data.csv")).read().splitlines()
            if l.strip() and not l.startswith("#")]
    return meta, rows


def _T_BS(meta):
    return np.array(meta["T_BS"]["data"], np.float64).reshape(4, 4)


class ASLStream:
    def __init__(self, pose_path, rgb_path, world_frame="RFU", depth_path=None, depth_scale_factor=1000.0):
        if world_frame not in WORLD_FRAMES_ASL:
            raise ValueError("unknown world frame %r" % world_frame)
        self.world_frame, self.rgb_path, self.depth_path = world_frame, rgb_path, depth_path
        self.depth_scale = depth_scale_factor
        pmeta, prow = _read_asl_sensor(pose_path)
        cmeta, crow = _read_asl_sensor(rgb_path)
        self.pose_t, self.pose_q, self.pose_ts = [], [], []
        for r in prow:
            v = r.split(",")
            self.pose_ts.append(int(v[0]))
            self.pose_t.append([float(x) for x in v[1:4]])
            qw, qx, qy, qz = [float(x) for x in v[4:8]]  # CSV is w,x,y,z; stored x,y,z,w (types.h:66-77)
            self.pose_q.append([qx, qy, qz, qw])
        self.img_ts = [int(r.split(",")[0]) for r in crow]
        self.img_files = [r.split(",")[1].strip() for r in crow]
        ia, ib = associate(np.array(self.img_ts) * 1e-9, np.array(self.pose_ts) * 1e-9)
        self.rgb_idxs, self.pose_idxs = ia, ib
        Tp, Tc = _T_BS(pmeta), _T_BS(cmeta)
        self.q_pose_in_body, self.t_pose_in_body = q_from_R(Tp[:3, :3]), Tp[:3, 3]
        self.q_cam_in_body, self.t_cam_in_body = q_from_R(Tc[:3, :3]), Tc[:3, 3]
        fu, fv, cu, cv_ = cmeta["intrinsics"]
        self.width, self.height = [int(v) for v in cmeta["resolution"]]
        self.K = np.array([[fu, 0, cu], [0, fv, cv_], [0, 0, 1]], np.float32)
        self.D = np.array(cmeta.get("distortion_coefficients", [0, 0, 0, 0]), np.float64)
        self.idx = 0

    def empty(self):
        return self.idx >= len(self.rgb_idxs)

    def get(self):
        import cv2
        ri, pi = self.rgb_idxs[self.idx], self.pose_idxs[self.idx]
        tm = self.img_ts[ri] * 1e-9
        q_pw = np.array(self.pose_q[pi], np.float64)
        q_pw /= np.linalg.norm(q_pw)
        t_pw = np.array(self.pose_t[pi], np.float64)
        # pose sensor -> body -> camera (asl_rgbd_offline_stream.cc:215-225)
        q_bp = q_inv(self.q_pose_in_body)
        t_bp = -q_rot(q_bp, self.t_pose_in_body)
        q_bw = q_mul(q_pw, q_bp)
        t_bw = q_rot(q_pw, t_bp) + t_pw
        q_cw = q_mul(q_bw, self.q_cam_in_body)
        t_cw = q_rot(q_bw, self.t_cam_in_body) + t_bw
        quat, trans = asl_world_to_rdf(q_cw, t_cw, self.world_frame)
        rgb = cv2.imread(os.path.join(self.rgb_path, "data", self.img_files[ri]), cv2.IMREAD_UNCHANGED)
        if rgb is None:
            raise IOError("cannot read %s" % self.img_files[ri])
        rgb = _rectify(rgb, self.K.astype(np.float64), self.D, self.K.astype(np.float64))
        out = (self.idx, tm, _to_gray(rgb), np.zeros(rgb.shape[:2], np.float32), quat, trans)
        self.idx += 1
        return out


# ------------------------------------------------------------------------------------- outputs
def depth_mesh(K, W, H, mesh):
    """publishDepthMesh's payload: points [V,3] (NaN = invalid vertex), normals, uv, faces [F,3]
    (reversed winding, valid triangles only). Returns None when no triangle is valid."""
    Kinv = np.linalg.inv(np.asarray(K, np.float64))
    vtx, idp = mesh["vtx"].astype(np.float64), mesh["idepth"].astype(np.float64)
    ok = np.isfinite(idp) & (idp > 0)
    uhom = np.concatenate([vtx, np.ones((len(vtx), 1))], axis=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        pts = (uhom / idp[:, None]) @ Kinv.T
    pts[~ok] = np.nan
    uv = vtx / np.array([W - 1.0, H - 1.0])
    faces = mesh["tris"][mesh["tri_valid"].astype(bool)][:, ::-1]
    if len(faces) == 0:
        return None
    return dict(points=pts.astype(np.float32), normals=mesh["normals"], uv=uv.astype(np.float32),
                faces=np.ascontiguousarray(faces, np.int32))


def write_ply(path, dm):
    pts, faces = dm["points"], dm["faces"]
    with open(path, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
                "property float nx\nproperty float ny\nproperty float nz\nproperty float s\nproperty float t\n"
                "element face %d\nproperty list uchar int vertex_indices\nend_header\n" % (len(pts), len(faces)))
        for p, n, uv in zip(pts, dm["normals"], dm["uv"]):
            f.write("%g %g %g %g %g %g %g %g\n" % (p[0], p[1], p[2], n[0], n[1], n[2], uv[0], uv[1]))
        for t in faces:
            f.write("3 %d %d %d\n" % (t[0], t[1], t[2]))


def idepth_to_depth(idm):
    """1/idepth, NaN where idepth is NaN or <= 0 (/root/reference/src/flame_nodelet.cc:694-700)."""
    d = np.full(idm.shape, np.nan, np.float32)
    ok = np.isfinite(idm) & (idm > 0)
    d[ok] = 1.0 / idm[ok]
    return d


def run_offline(stream, capi, update_params=None, poseframe_subsample_factor=6, subsample_factor=1,
                max_frames=None, out_dir=None, on_frame=None):
    """The offline loop of FlameOffline{TUM}::main/processFrame without ROS. Returns a stats dict."""
    up = update_params or capi.default_update_params()
    cells = (stream.width // up.detection_win_size) * (stream.height // up.detection_win_size)
    maxF = max(1024, 2 * cells)
    stats = dict(frames=0, updates=0, vertices=[], ms=[])
    with capi.Context(1, stream.width, stream.height, 8, maxF, maxF, 3 * maxF) as ctx:
        ctx.set_intrinsics(0, stream.K)
        ctx.set_update_params(up)
        filt = capi.default_tri_filter_params()
        while not stream.empty() and (max_frames is None or stats["frames"] < max_frames):
            img_id, tm, gray, depth, quat, trans = stream.get()
            if img_id % subsample_factor != 0:
                continue
            pose = np.concatenate([quat, trans]).astype(np.float32)
            is_pf = (img_id % poseframe_subsample_factor) == 0  # flame_offline_tum.cc:575
            t0 = time.perf_counter()
            ok = ctx.update(0, tm, img_id, pose, gray, is_pf)
            stats["ms"].append(1e3 * (time.perf_counter() - t0))
            stats["frames"] += 1
            if not ok:
                continue  # "update_success false -> warn, return" (flame_offline_tum.cc:597-601)
            stats["updates"] += 1
            mesh = ctx.get_mesh(0, filt)
            stats["vertices"].append(len(mesh["idepth"]))
            if on_frame is not None:
                on_frame(img_id, tm, ctx, mesh, depth)
            if out_dir is not None:
                dm = depth_mesh(stream.K, stream.width, stream.height, mesh)
                if dm is not None:
                    write_ply(os.path.join(out_dir, "mesh_%06d.ply" % img_id), dm)
                np.save(os.path.join(out_dir, "depth_%06d.npy" % img_id), idepth_to_depth(ctx.get_idepthmap(0, filt)))
    ms = np.array(stats["ms"][2:] or stats["ms"] or [0.0])
    stats.update(fps=float(1e3 / ms.mean()) if ms.mean() > 0 else 0.0, ms_median=float(np.median(ms)),
                 vertices_last=stats["vertices"][-1] if stats["vertices"] else 0)
    return stats


def main(argv=None):
    ap = argparse.ArgumentParser(description="ROS-free FLaME offline player (TUM RGB-D / EuRoC-ASL formats)")
    ap.add_argument("--tum", help="TUM association file")
    ap.add_argument("--calib", help="camera_info YAML (TUM)")
    ap.add_argument("--input-frame", default="RDF_IN_FLU", choices=POSE_FRAMES_TUM)
    ap.add_argument("--asl-pose", help="ASL pose sensor directory (sensor.yaml + data.csv)")
    ap.add_argument("--asl-rgb", help="ASL camera directory")
    ap.add_argument("--world-frame", default="RFU", choices=WORLD_FRAMES_ASL)
    ap.add_argument("--poseframe-subsample-factor", type=int, default=6)
    ap.add_argument("--max-frames", type=int)
    ap.add_argument("--out", help="directory for mesh_*.ply / depth_*.npy")
    a = ap.parse_args(argv)
    from . import capi
    if a.tum:
        stream = TUMStream(a.tum, a.calib, a.input_frame)
    elif a.asl_pose and a.asl_rgb:
        stream = ASLStream(a.asl_pose, a.asl_rgb, a.world_frame)
    else:
        ap.error("give --tum/--calib or --asl-pose/--asl-rgb")
    if a.out:
        os.makedirs(a.out, exist_ok=True)
    st = run_offline(stream, capi, None, a.poseframe_subsample_factor, 1, a.max_frames, a.out)
    print(json.dumps({k: v for k, v in st.items() if k not in ("ms", "vertices")}))


if __name__ == "__main__":
    main()
