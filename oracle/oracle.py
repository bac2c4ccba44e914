"""ctypes binding of the CPU oracle (oracle/flame_oracle.{h,c}).

TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED (see flame_oracle.h).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; nothing under flame_ros_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_LIB_NATIVE = None

STATUS_NAMES = ["SUCCESS", "FAIL_REF_PATCH_GRADIENT", "FAIL_AMBIGUOUS_MATCH", "FAIL_MAX_COST",
                "FAIL_MAX_VAR", "FAIL_MAX_DROPOUTS", "FAIL_OUT_OF_IMAGE", "NO_PARALLAX", "SKIPPED"]
NUM_COUNTERS = 8


class NLTGV2Params(C.Structure):
    _fields_ = [("data_factor", C.c_float), ("step_x", C.c_float), ("step_q", C.c_float),
                ("theta", C.c_float), ("x_min", C.c_float), ("x_max", C.c_float)]

    @classmethod
    def default(cls):
        # /root/reference/cfg/flame_nodelet.yaml:86-89
        return cls(0.15, 0.001, 125.0, 0.25, 0.0, 10.0)


class EpiParams(C.Structure):
    _fields_ = [("win_size", C.c_int), ("min_grad_mag", C.c_float), ("epipolar_line_var", C.c_float),
                ("max_dropouts", C.c_int), ("search_sigma", C.c_float), ("max_cost", C.c_float),
                ("ambiguity_ratio", C.c_float), ("ambiguity_radius", C.c_int),
                ("pixel_noise_var", C.c_float), ("meas_var_max", C.c_float),
                ("idepth_min", C.c_float), ("idepth_max", C.c_float), ("max_search_px", C.c_int),
                ("min_parallax", C.c_float)]

    @classmethod
    def default(cls):
        # win 5, min_grad 5, line var 4, dropouts 5: /root/reference/cfg/flame_nodelet.yaml:69-75
        return cls(5, 5.0, 4.0, 5, 2.0, 400.0, 1.5, 2, 4.0, 1.0, 0.0, 10.0, 64, 0.5)


class TriFilterParams(C.Structure):
    _fields_ = [("do_oblique", C.c_int), ("oblique_normal_thresh", C.c_float),
                ("oblique_idepth_diff_factor", C.c_float), ("oblique_idepth_diff_abs", C.c_float),
                ("do_edge_length", C.c_int), ("edge_length_thresh", C.c_float),
                ("do_idepth", C.c_int), ("min_triangle_idepth", C.c_float)]

    @classmethod
    def default(cls):
        # /root/reference/cfg/flame_nodelet.yaml:31-46
        return cls(1, 1.57, 0.35, 0.1, 1, 0.333, 1, 0.01)


def build(native=False):
    """Compile the oracle with gcc (make -C oracle [native])."""
    env = dict(os.environ)
    env.pop("CC", None)
    cmd = ["make", "-C", _HERE] + (["native"] if native else [])
    subprocess.run(cmd, check=True, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _bp(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def load(native=False):
    """Load (building if necessary) the oracle shared library."""
    global _LIB, _LIB_NATIVE
    if native and _LIB_NATIVE is not None:
        return _LIB_NATIVE
    if not native and _LIB is not None:
        return _LIB
    name = "libflame_oracle_native.so" if native else "libflame_oracle.so"
    path = os.path.join(_HERE, "_build", name)
    srcs = [os.path.join(_HERE, f) for f in ("flame_oracle.c", "flame_pipeline.c", "flame_oracle.h")]
    if native or not os.path.exists(path) or os.path.getmtime(path) < max(os.path.getmtime(f) for f in srcs):
        build(native=native)
    lib = C.CDLL(path)
    for fn in ("fo_nltgv2_solve", "fo_nltgv2_costs", "fo_idepth_update", "fo_epi_geometry",
               "fo_project_features", "fo_gradient_mag", "fo_pyr_down", "fo_triangle_validity",
               "fo_rasterize_idepth"):
        getattr(lib, fn).restype = None
    lib.fo_detect_features.restype = C.c_int
    for fn in ("fo_default_update_params", "fo_pipeline_destroy", "fo_pipeline_sizes", "fo_pipeline_mesh",
               "fo_pipeline_idepthmap", "fo_pipeline_features", "fo_pipeline_stage_ms"):
        getattr(lib, fn).restype = None
    lib.fo_delaunay.restype = C.c_int
    lib.fo_pipeline_create.restype = C.c_void_p
    lib.fo_pipeline_update.restype = C.c_int
    if native:
        _LIB_NATIVE = lib
    else:
        _LIB = lib
    return lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def nltgv2_solve(pos, edge_ij, alpha, beta, z, wt, state, params, iters, nthreads=1, native=False):
    """state: dict with x,w1,w2,xb,w1b,w2b [V], q1,q2,q3 [E] float32 arrays (modified in place)."""
    lib = load(native)
    pos, edge_ij = _f32(pos), _i32(edge_ij)
    alpha, beta, z, wt = _f32(alpha), _f32(beta), _f32(z), _f32(wt)
    V, E = z.shape[0], alpha.shape[0]
    for k in ("x", "w1", "w2", "xb", "w1b", "w2b"):
        assert state[k].dtype == np.float32 and state[k].flags.c_contiguous and state[k].shape == (V,)
    for k in ("q1", "q2", "q3"):
        assert state[k].dtype == np.float32 and state[k].flags.c_contiguous and state[k].shape == (E,)
    lib.fo_nltgv2_solve(C.c_int(V), C.c_int(E), _fp(pos), _ip(edge_ij), _fp(alpha), _fp(beta),
                        _fp(z), _fp(wt), _fp(state["x"]), _fp(state["w1"]), _fp(state["w2"]),
                        _fp(state["xb"]), _fp(state["w1b"]), _fp(state["w2b"]), _fp(state["q1"]),
                        _fp(state["q2"]), _fp(state["q3"]), C.byref(params), C.c_int(iters),
                        C.c_int(nthreads))
    return state


def nltgv2_costs(pos, edge_ij, alpha, beta, z, wt, x, w1, w2, data_factor):
    lib = load()
    pos, edge_ij = _f32(pos), _i32(edge_ij)
    alpha, beta, z, wt = _f32(alpha), _f32(beta), _f32(z), _f32(wt)
    x, w1, w2 = _f32(x), _f32(w1), _f32(w2)
    s, d = C.c_double(0), C.c_double(0)
    lib.fo_nltgv2_costs(C.c_int(z.shape[0]), C.c_int(alpha.shape[0]), _fp(pos), _ip(edge_ij),
                        _fp(alpha), _fp(beta), _fp(z), _fp(wt), _fp(x), _fp(w1), _fp(w2),
                        C.c_float(data_factor), C.byref(s), C.byref(d))
    return s.value, d.value


def new_state(z, E):
    """Cold start used by every test: x = xb = z, w = wb = 0, q = 0."""
    z = _f32(z)
    V = z.shape[0]
    st = {k: np.zeros(V, np.float32) for k in ("w1", "w2", "w1b", "w2b")}
    st["x"] = z.copy()
    st["xb"] = z.copy()
    for k in ("q1", "q2", "q3"):
        st[k] = np.zeros(E, np.float32)
    return st


def epi_geometry(K, pose_ref, pose_cmp):
    lib = load()
    G = np.zeros(15, np.float32)
    lib.fo_epi_geometry(_fp(_f32(K).ravel()), _fp(_f32(pose_ref)), _fp(_f32(pose_cmp)), _fp(G))
    return G


def idepth_update(imgs, poses, K, cmp_slot, ref_slot, u_ref, mu, var, dropouts, alive, params,
                  nthreads=1, native=False):
    """imgs [n_slots,H,W] uint8; mu,var,dropouts,alive are modified in place.
    Returns (status [N] int32, u_cmp [N,2] float32, counters [8] int32)."""
    lib = load(native)
    imgs = np.ascontiguousarray(imgs, dtype=np.uint8)
    n_slots, H, W = imgs.shape
    poses = _f32(poses)
    Kf = _f32(K).ravel()
    ref_slot, u_ref = _i32(ref_slot), _f32(u_ref)
    N = ref_slot.shape[0]
    for a, dt in ((mu, np.float32), (var, np.float32), (dropouts, np.int32), (alive, np.int32)):
        assert a.dtype == dt and a.flags.c_contiguous and a.shape == (N,)
    status = np.zeros(N, np.int32)
    u_cmp = np.zeros((N, 2), np.float32)
    counters = np.zeros(NUM_COUNTERS, np.int32)
    lib.fo_idepth_update(C.c_int(W), C.c_int(H), C.c_int(n_slots), _bp(imgs), _fp(poses), _fp(Kf),
                         C.c_int(cmp_slot), C.c_int(N), _ip(ref_slot), _fp(u_ref), _fp(mu), _fp(var),
                         _ip(dropouts), _ip(alive), _ip(status), _fp(u_cmp), _ip(counters),
                         C.byref(params), C.c_int(nthreads))
    return status, u_cmp, counters


def project_features(W, H, poses, K, cur_slot, ref_slot, u_ref, mu, var, alive):
    lib = load()
    poses = _f32(poses)
    n_slots = poses.shape[0]
    ref_slot, u_ref, mu, var, alive = _i32(ref_slot), _f32(u_ref), _f32(mu), _f32(var), _i32(alive)
    N = ref_slot.shape[0]
    u_cur = np.zeros((N, 2), np.float32)
    mu_cur = np.zeros(N, np.float32)
    var_cur = np.zeros(N, np.float32)
    valid = np.zeros(N, np.int32)
    lib.fo_project_features(C.c_int(W), C.c_int(H), C.c_int(n_slots), _fp(poses), _fp(_f32(K).ravel()),
                            C.c_int(cur_slot), C.c_int(N), _ip(ref_slot), _fp(u_ref), _fp(mu),
                            _fp(var), _ip(alive), _fp(u_cur), _fp(mu_cur), _fp(var_cur), _ip(valid))
    return u_cur, mu_cur, var_cur, valid


def gradient_mag(img):
    lib = load()
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape
    mag = np.zeros((H, W), np.float32)
    lib.fo_gradient_mag(C.c_int(W), C.c_int(H), _bp(img), _fp(mag))
    return mag


def pyr_down(img):
    lib = load()
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape
    out = np.zeros((H // 2, W // 2), np.uint8)
    lib.fo_pyr_down(C.c_int(W), C.c_int(H), _bp(img), _bp(out))
    return out


def detect_features(mag, win, border, min_grad_mag, occupied=None):
    lib = load()
    mag = _f32(mag)
    H, W = mag.shape
    cells = (W // win) * (H // win)
    det_xy = np.zeros((cells, 2), np.float32)
    det_ok = np.zeros(cells, np.int32)
    occ = None
    if occupied is not None:
        occ = np.ascontiguousarray(occupied, dtype=np.uint8).ravel()
        assert occ.shape[0] == cells
    n = lib.fo_detect_features(C.c_int(W), C.c_int(H), _fp(mag), C.c_int(win), C.c_int(border),
                               C.c_float(min_grad_mag), _bp(occ) if occ is not None else None,
                               _fp(det_xy), _ip(det_ok))
    return n, det_xy, det_ok


def triangle_validity(W, H, K, vtx, idepth, tri, fparams):
    lib = load()
    vtx, idepth, tri = _f32(vtx), _f32(idepth), _i32(tri)
    T = tri.shape[0]
    valid = np.zeros(T, np.uint8)
    lib.fo_triangle_validity(C.c_int(W), C.c_int(H), _fp(_f32(K).ravel()), C.c_int(idepth.shape[0]),
                             _fp(vtx), _fp(idepth), C.c_int(T), _ip(tri), C.byref(fparams), _bp(valid))
    return valid


def rasterize_idepth(W, H, vtx, idepth, tri, valid=None):
    lib = load()
    vtx, idepth, tri = _f32(vtx), _f32(idepth), _i32(tri)
    T = tri.shape[0]
    out = np.zeros((H, W), np.float32)
    v = None
    if valid is not None:
        v = np.ascontiguousarray(valid, dtype=np.uint8)
    lib.fo_rasterize_idepth(C.c_int(W), C.c_int(H), C.c_int(idepth.shape[0]), _fp(vtx), _fp(idepth),
                            C.c_int(T), _ip(tri), _bp(v) if v is not None else None, _fp(out))
    return out


class UpdateParams(C.Structure):
    _fields_ = [("detection_win_size", C.c_int), ("min_grad_mag", C.c_float), ("detection_border", C.c_int),
                ("idepth_init", C.c_float), ("idepth_var_init", C.c_float), ("idepth_var_max_graph", C.c_float),
                ("adaptive_data_weights", C.c_int), ("init_with_prediction", C.c_int), ("do_nltgv2", C.c_int),
                ("iters", C.c_int), ("rparams", NLTGV2Params), ("rescale_data", C.c_int), ("min_height", C.c_float),
                ("max_height", C.c_float), ("check_sticky_obstacles", C.c_int), ("min_error", C.c_float),
                ("do_letterbox", C.c_int)]

    @classmethod
    def default(cls):
        p = cls()
        load().fo_default_update_params(C.byref(p))
        return p

    @classmethod
    def like(cls, other):
        """Copy the fields this struct shares with another parameter struct (the product's fb_update_params)."""
        p = cls.default()
        for n, _ in cls._fields_:
            if n == "rparams":
                for m, _ in NLTGV2Params._fields_:
                    setattr(p.rparams, m, getattr(other.rparams, m))
            elif hasattr(other, n):
                setattr(p, n, getattr(other, n))
        return p


STAGES = ["update", "frame_creation", "update_idepths", "project_features", "sync_graph", "triangulate", "nltgv2",
          "interpolate", "detection"]


def delaunay(pts):
    """The oracle's own triangulator (sorted sweep + Lawson flips, exact predicates).  Returns
    (tris [T,3], edges [E,2]) in the canonical order, or raises ValueError when degenerate."""
    lib = load()
    pts = _f32(pts)
    n = pts.shape[0]
    tris = np.zeros((max(2 * n, 1), 3), np.int32)
    edges = np.zeros((max(3 * n, 1), 2), np.int32)
    nt, ne = C.c_int32(0), C.c_int32(0)
    rc = lib.fo_delaunay(C.c_int(n), _fp(pts), _ip(tris), C.byref(nt), _ip(edges), C.byref(ne))
    if rc != 0:
        raise ValueError("fo_delaunay: degenerate input")
    return tris[:nt.value].copy(), edges[:ne.value].copy()


class Pipeline:
    """fo_pipeline_*: the whole flame::Flame::update pipeline on the CPU."""

    def __init__(self, W, H, K, n_slots, max_features, max_vertices, up=None, ep=None, nthreads=1, native=False):
        self._lib = load(native)
        self._lib.fo_pipeline_create.restype = C.c_void_p
        self._lib.fo_pipeline_update.restype = C.c_int
        self.W, self.H, self.maxF, self.maxV = W, H, max_features, max_vertices
        self.up = up or UpdateParams.default()
        self.ep = ep or EpiParams.default()
        Kf = _f32(K).ravel()
        self._h = C.c_void_p(self._lib.fo_pipeline_create(C.c_int(W), C.c_int(H), _fp(Kf), C.c_int(n_slots),
                                                          C.c_int(max_features), C.c_int(max_vertices),
                                                          C.byref(self.up), C.byref(self.ep), C.c_int(nthreads)))
        if not self._h:
            raise ValueError("fo_pipeline_create failed")

    def close(self):
        if self._h:
            self._lib.fo_pipeline_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def update(self, img_id, pose, gray, is_poseframe):
        gray = np.ascontiguousarray(gray, np.uint8)
        assert gray.shape == (self.H, self.W)
        return bool(self._lib.fo_pipeline_update(self._h, C.c_int(img_id), _fp(_f32(pose)), _bp(gray),
                                                 C.c_int(1 if is_poseframe else 0)))

    def mesh(self):
        V, T, E = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        self._lib.fo_pipeline_sizes(self._h, C.byref(V), C.byref(T), C.byref(E))
        out = dict(vtx=np.zeros((V.value, 2), np.float32), idepth=np.zeros(V.value, np.float32),
                   tris=np.zeros((T.value, 3), np.int32), edges=np.zeros((E.value, 2), np.int32),
                   vert_feat=np.zeros(V.value, np.int32))
        self._lib.fo_pipeline_mesh(self._h, _fp(out["vtx"]), _fp(out["idepth"]), _ip(out["tris"]), _ip(out["edges"]),
                                   _ip(out["vert_feat"]))
        return out

    def idepthmap(self, filter_params=None):
        out = np.zeros((self.H, self.W), np.float32)
        self._lib.fo_pipeline_idepthmap(self._h, C.byref(filter_params) if filter_params is not None else None, _fp(out))
        return out

    def features(self):
        F = self.maxF
        out = dict(u_ref=np.zeros((F, 2), np.float32), ref_slot=np.zeros(F, np.int32), mu=np.zeros(F, np.float32),
                   var=np.zeros(F, np.float32), dropouts=np.zeros(F, np.int32), alive=np.zeros(F, np.int32),
                   valid=np.zeros(F, np.int32))
        self._lib.fo_pipeline_features(self._h, _fp(out["u_ref"]), _ip(out["ref_slot"]), _fp(out["mu"]), _fp(out["var"]),
                                       _ip(out["dropouts"]), _ip(out["alive"]), _ip(out["valid"]))
        return out

    def stage_ms(self):
        ms = (C.c_double * len(STAGES))()
        self._lib.fo_pipeline_stage_ms(self._h, ms)
        return dict(zip(STAGES, list(ms)))
