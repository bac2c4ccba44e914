"""ctypes binding of libflame_b200.so (include/flame_b200.h) and a thin numpy-facing wrapper.

This is the host-side mirror used by the tests and bench.py; every call goes through the C-ABI that
the C++ `flame::Flame` shim (include/flame/flame.h) uses.  There is no CPU fallback: if the shared
library is missing or no CUDA device is present, construction raises.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

NUM_COUNTERS = 8
STATUS_NAMES = ["SUCCESS", "FAIL_REF_PATCH_GRADIENT", "FAIL_AMBIGUOUS_MATCH", "FAIL_MAX_COST",
                "FAIL_MAX_VAR", "FAIL_MAX_DROPOUTS", "FAIL_OUT_OF_IMAGE", "NO_PARALLAX", "SKIPPED"]
PROF_SOLVE, PROF_IDEPTH, PROF_UPLOAD, PROF_ASSEMBLY, PROF_INTERP = range(5)

# every symbol include/flame_b200.h declares (tests/test_capi_symbols.py checks this list against
# the header and the built library)
SYMBOLS = [
    "fb_create", "fb_destroy", "fb_last_error", "fb_sync", "fb_version", "fb_host_alloc",
    "fb_host_free", "fb_set_intrinsics", "fb_set_epi_params", "fb_default_epi_params",
    "fb_default_nltgv2_params", "fb_default_tri_filter_params", "fb_graph_set", "fb_graph_data_set",
    "fb_graph_state_set", "fb_graph_state_get", "fb_graph_x_get_all", "fb_nltgv2_solve", "fb_costs",
    "fb_frame_set", "fb_frame_pose_set", "fb_pool_reserve", "fb_pool_upload", "fb_frame_from_pool",
    "fb_features_set", "fb_features_get", "fb_features_reinit", "fb_idepth_update", "fb_idepth_counters",
    "fb_project_features", "fb_graph_bind_features", "fb_graph_data_from_features", "fb_mesh_set",
    "fb_interpolate", "fb_profile_enable", "fb_profile_reset", "fb_profile_get", "fb_launch_count",
    "fb_last_solver_variant", "fb_last_cluster_size", "fb_last_solver_transport", "fb_grid_plan_verify", "fb_delaunay", "fb_hotpath_step", "fb_results_wait", "fb_pipeline_join", "fb_default_update_params",
    "fb_set_update_params", "fb_update", "fb_update_run", "fb_get_mesh_sizes", "fb_get_mesh", "fb_get_idepthmap",
    "fb_get_raw_idepths", "fb_get_stat", "fb_update_poseframe_poses", "fb_prune_poseframes",
    "fb_frame_gradient", "fb_frame_pyr_down", "fb_detect", "fb_get_feature_pool", "fb_delaunay_device",
]


class FlameError(RuntimeError):
    pass


class NLTGV2Params(C.Structure):
    """flame::Params::rparams (/root/reference/src/flame_nodelet.cc:256-259)."""
    _fields_ = [("data_factor", C.c_float), ("step_x", C.c_float), ("step_q", C.c_float),
                ("theta", C.c_float), ("x_min", C.c_float), ("x_max", C.c_float)]


class EpiParams(C.Structure):
    """flame::Params::{fparams,zparams,max_dropouts} (/root/reference/src/flame_nodelet.cc:227-245)."""
    _fields_ = [("win_size", C.c_int), ("min_grad_mag", C.c_float), ("epipolar_line_var", C.c_float),
                ("max_dropouts", C.c_int), ("search_sigma", C.c_float), ("max_cost", C.c_float),
                ("ambiguity_ratio", C.c_float), ("ambiguity_radius", C.c_int),
                ("pixel_noise_var", C.c_float), ("meas_var_max", C.c_float),
                ("idepth_min", C.c_float), ("idepth_max", C.c_float), ("max_search_px", C.c_int),
                ("min_parallax", C.c_float)]


class TriFilterParams(C.Structure):
    """output/filter_* (/root/reference/src/flame_nodelet.cc:182-206)."""
    _fields_ = [("do_oblique", C.c_int), ("oblique_normal_thresh", C.c_float),
                ("oblique_idepth_diff_factor", C.c_float), ("oblique_idepth_diff_abs", C.c_float),
                ("do_edge_length", C.c_int), ("edge_length_thresh", C.c_float),
                ("do_idepth", C.c_int), ("min_triangle_idepth", C.c_float)]


class UpdateParams(C.Structure):
    """fb_update_params: the subset of flame::Params that drives flame::Flame::update
    (/root/reference/src/flame_nodelet.cc:225-259)."""
    _fields_ = [("detection_win_size", C.c_int), ("min_grad_mag", C.c_float), ("detection_border", C.c_int),
                ("idepth_init", C.c_float), ("idepth_var_init", C.c_float), ("idepth_var_max_graph", C.c_float),
                ("adaptive_data_weights", C.c_int), ("init_with_prediction", C.c_int), ("do_nltgv2", C.c_int),
                ("iters", C.c_int), ("rparams", NLTGV2Params), ("triangulator", C.c_int),
                ("rescale_data", C.c_int), ("min_height", C.c_float), ("max_height", C.c_float),
                ("check_sticky_obstacles", C.c_int), ("min_error", C.c_float), ("do_letterbox", C.c_int)]


class StepDesc(C.Structure):
    """fb_step_desc: one frame of every stream through the hot path in a single call."""
    _fields_ = [("new_poseframe", C.c_int), ("ref_slot", C.c_int), ("cmp_slot", C.c_int),
                ("ref_images", C.POINTER(C.c_void_p)), ("cmp_images", C.POINTER(C.c_void_p)),
                ("ref_pool_idx", C.POINTER(C.c_int32)), ("cmp_pool_idx", C.POINTER(C.c_int32)),
                ("ref_poses", C.POINTER(C.c_float)), ("cmp_poses", C.POINTER(C.c_float)),
                ("mu0", C.c_float), ("var0", C.c_float), ("adaptive_weights", C.c_int),
                ("iters", C.c_int), ("variant", C.c_int), ("rparams", NLTGV2Params),
                ("x_out", C.POINTER(C.c_float)), ("pipelined", C.c_int), ("ref_from_slot", C.c_int)]


_LIB = None


def lib_path():
    return _build.LIB_PATH


def load_library(build_if_missing=True):
    """dlopen libflame_b200.so (building it in-tree with nvcc when absent/stale and allowed)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB_PATH
    if build_if_missing and _build.is_stale():
        try:
            _build.build()
        except Exception as exc:  # stale-but-present library is still usable on a box without nvcc
            if not os.path.exists(path):
                raise FlameError("libflame_b200.so is missing and could not be built: %s" % exc)
    if not os.path.exists(path):
        raise FlameError("libflame_b200.so not found at %s (run __graft_entry__.build())" % path)
    lib = C.CDLL(path)
    lib.fb_create.restype = C.c_void_p
    lib.fb_create.argtypes = [C.c_int] * 8 + [C.c_void_p]
    lib.fb_destroy.restype = None
    lib.fb_destroy.argtypes = [C.c_void_p]
    lib.fb_last_error.restype = C.c_char_p
    lib.fb_last_error.argtypes = [C.c_void_p]
    lib.fb_host_alloc.restype = C.c_void_p
    lib.fb_host_alloc.argtypes = [C.c_size_t]
    lib.fb_host_free.restype = None
    lib.fb_host_free.argtypes = [C.c_void_p]
    lib.fb_launch_count.restype = C.c_int64
    lib.fb_launch_count.argtypes = [C.c_void_p]
    for name in ("fb_default_epi_params", "fb_default_nltgv2_params", "fb_default_tri_filter_params",
                 "fb_default_update_params"):
        getattr(lib, name).restype = None
    P = C.c_void_p
    I = C.c_int
    sigs = {
        "fb_sync": [P], "fb_set_intrinsics": [P, I, P], "fb_set_epi_params": [P, P],
        "fb_graph_set": [P, I, I, I, P, P, P, P], "fb_graph_data_set": [P, I, P, P],
        "fb_graph_state_set": [P, I, P, P, P], "fb_graph_state_get": [P, I, P, P, P, P],
        "fb_graph_x_get_all": [P, P], "fb_nltgv2_solve": [P, I, P, I],
        "fb_costs": [P, I, C.c_float, P, P], "fb_frame_set": [P, I, I, P, I, P],
        "fb_frame_pose_set": [P, I, I, P], "fb_pool_reserve": [P, I], "fb_pool_upload": [P, I, P, I],
        "fb_frame_from_pool": [P, I, I, I, P], "fb_features_set": [P, I, I, P, P, P, P, P, P],
        "fb_features_get": [P, I, P, P, P, P, P, P], "fb_idepth_update": [P, P], "fb_features_reinit": [P, P, C.c_float, C.c_float],
        "fb_idepth_counters": [P, I, P], "fb_project_features": [P, I, I, P, P, P, P],
        "fb_graph_bind_features": [P, I, P], "fb_graph_data_from_features": [P, I],
        "fb_mesh_set": [P, I, I, P], "fb_interpolate": [P, I, P, P, P],
        "fb_profile_enable": [P, I], "fb_profile_reset": [P], "fb_profile_get": [P, I, P, P, P],
        "fb_last_solver_variant": [P], "fb_last_cluster_size": [P], "fb_last_solver_transport": [P], "fb_grid_plan_verify": [I, I, P, P, I, I, P], "fb_version": [], "fb_delaunay": [I, P, P, P, P, P], "fb_hotpath_step": [P, P], "fb_results_wait": [P, I], "fb_pipeline_join": [P],
        "fb_set_update_params": [P, P], "fb_update": [P, I, C.c_double, I, P, P, I, I],
        "fb_get_mesh_sizes": [P, I, P, P, P], "fb_get_mesh": [P, I, P, P, P, P, P, P, P],
        "fb_get_idepthmap": [P, I, P, P], "fb_update_run": [P, I, I, I, P, C.c_size_t, P, I, P, P], "fb_get_raw_idepths": [P, I, P, P, P, P],
        "fb_get_stat": [P, I, C.c_char_p, P], "fb_update_poseframe_poses": [P, I, I, P, P],
        "fb_prune_poseframes": [P, I, I, P], "fb_frame_gradient": [P, I, I, P],
        "fb_frame_pyr_down": [P, I, I, P], "fb_detect": [P, I, I, I, I, C.c_float, P, P, P, P],
        "fb_get_feature_pool": [P, I, P, P, P, P, P, P],
        "fb_delaunay_device": [P, I, I, P, P, P, P, P],
    }
    for name, args in sigs.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = args
    _LIB = lib
    return lib


def default_nltgv2_params():
    p = NLTGV2Params()
    load_library().fb_default_nltgv2_params(C.byref(p))
    return p


def default_epi_params():
    p = EpiParams()
    load_library().fb_default_epi_params(C.byref(p))
    return p


def default_update_params():
    p = UpdateParams()
    load_library().fb_default_update_params(C.byref(p))
    return p


def default_tri_filter_params():
    p = TriFilterParams()
    load_library().fb_default_tri_filter_params(C.byref(p))
    return p


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def delaunay(pts):
    """Exact-predicate Delaunay triangulation (host code of libflame_b200; no GPU needed).
    Returns (tris [T,3] int32, edges [E,2] int32 canonical)."""
    lib = load_library()
    pts = _f32(pts)
    n = pts.shape[0]
    tris = np.zeros((max(2 * n, 1), 3), np.int32)
    edges = np.zeros((max(3 * n, 1), 2), np.int32)
    nt, ne = C.c_int32(0), C.c_int32(0)
    rc = lib.fb_delaunay(n, _ptr(pts), _ptr(tris), C.byref(nt), _ptr(edges), C.byref(ne))
    if rc != 0:
        raise FlameError("fb_delaunay: degenerate input (rc %d)" % rc)
    return tris[:nt.value].copy(), edges[:ne.value].copy()


def grid_plan_verify(pos, edges, parts, cluster=False):
    """Host-only self-check of the grid-resident solver's partition tables (no GPU needed).
    Returns (rc, stats dict); rc 0 = invariants hold, 1 = does not fit `parts` CTAs."""
    lib = load_library()
    pos = _f32(pos)
    edges = np.ascontiguousarray(edges, np.int32).reshape(-1, 2)
    stats = np.zeros(12, np.int32)
    rc = lib.fb_grid_plan_verify(pos.shape[0], edges.shape[0], _ptr(pos), _ptr(edges), int(parts), int(bool(cluster)), _ptr(stats))
    if rc < 0:
        raise FlameError("fb_grid_plan_verify: bad argument")
    if rc > 1:
        raise FlameError(lib.fb_last_error(None).decode())
    keys = ("max_own", "max_generic", "max_halo", "cut_edges", "max_slots", "smem_bytes", "boundary", "overflow_edges",
            "wavefronts_ideal", "wavefronts_load", "wavefronts_store", "reserved")
    return rc, dict(zip(keys, stats.tolist()))


class PinnedBuffer:
    """Page-locked host memory from fb_host_alloc exposed as a numpy array."""

    def __init__(self, shape, dtype):
        self._lib = load_library()
        self.shape = tuple(int(s) for s in np.atleast_1d(shape))
        self.dtype = np.dtype(dtype)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self._p = self._lib.fb_host_alloc(nbytes)
        if not self._p:
            raise FlameError("fb_host_alloc(%d) failed" % nbytes)
        buf = (C.c_uint8 * nbytes).from_address(self._p)
        self.array = np.frombuffer(buf, dtype=self.dtype).reshape(self.shape)

    def free(self):
        if self._p:
            self.array = None
            self._lib.fb_host_free(self._p)
            self._p = None


class Context:
    """A batch of `n_streams` FLaME hot-path states on one GPU (wraps fb_ctx)."""

    def __init__(self, n_streams=1, width=640, height=480, n_slots=8, max_features=8192,
                 max_vertices=8192, max_edges=24576, device=0, cuda_stream=None):
        self._lib = load_library()
        self.S, self.W, self.H, self.n_slots = n_streams, width, height, n_slots
        self.max_features, self.max_vertices, self.max_edges = max_features, max_vertices, max_edges
        self._h = self._lib.fb_create(device, n_streams, width, height, n_slots, max_features,
                                      max_vertices, max_edges, cuda_stream)
        if not self._h:
            raise FlameError(self._lib.fb_last_error(None).decode())
        self._nV = [0] * n_streams
        self._nE = [0] * n_streams
        self._nF = [0] * n_streams
        self._nT = [0] * n_streams

    def close(self):
        if getattr(self, "_h", None):
            self._lib.fb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, rc):
        if rc != 0:
            raise FlameError("libflame_b200 error %d: %s" % (rc, self._lib.fb_last_error(self._h).decode()))

    # ------------------------------------------------------------------ configuration
    def sync(self):
        self._ck(self._lib.fb_sync(self._h))

    def set_intrinsics(self, stream, K):
        K = _f32(np.asarray(K).reshape(9))
        self._ck(self._lib.fb_set_intrinsics(self._h, stream, _ptr(K)))

    def set_epi_params(self, p):
        self._ck(self._lib.fb_set_epi_params(self._h, C.byref(p)))

    # ------------------------------------------------------------------ graph / solver
    def graph_set(self, stream, pos, edges, alpha, beta):
        pos, edges, alpha, beta = _f32(pos), _i32(edges), _f32(alpha), _f32(beta)
        V, E = pos.shape[0], edges.shape[0]
        self._ck(self._lib.fb_graph_set(self._h, stream, V, E, _ptr(pos), _ptr(edges), _ptr(alpha), _ptr(beta)))
        self._nV[stream], self._nE[stream] = V, E

    def graph_data_set(self, stream, z, wt=None):
        z, wt = _f32(z), _f32(wt)
        assert z.shape[0] == self._nV[stream]
        self._ck(self._lib.fb_graph_data_set(self._h, stream, _ptr(z), _ptr(wt)))

    def graph_state_set(self, stream, x=None, w=None, q=None):
        x, w, q = _f32(x), _f32(w), _f32(q)
        self._ck(self._lib.fb_graph_state_set(self._h, stream, _ptr(x), _ptr(w), _ptr(q)))

    def graph_state_get(self, stream):
        V, E = self._nV[stream], self._nE[stream]
        x = np.zeros(V, np.float32)
        w = np.zeros((V, 2), np.float32)
        q = np.zeros((E, 3), np.float32)
        xb = np.zeros((V, 3), np.float32)
        self._ck(self._lib.fb_graph_state_get(self._h, stream, _ptr(x), _ptr(w), _ptr(q), _ptr(xb)))
        return dict(x=x, w1=w[:, 0].copy(), w2=w[:, 1].copy(), q1=q[:, 0].copy(), q2=q[:, 1].copy(),
                    q3=q[:, 2].copy(), xb=xb[:, 0].copy(), w1b=xb[:, 1].copy(), w2b=xb[:, 2].copy())

    def graph_x_get_all(self, out=None):
        if out is None:
            out = np.zeros((self.S, self.max_vertices), np.float32)
        self._ck(self._lib.fb_graph_x_get_all(self._h, _ptr(out)))
        return out

    def nltgv2_solve(self, iters, params=None, variant=0):
        p = params if params is not None else default_nltgv2_params()
        self._ck(self._lib.fb_nltgv2_solve(self._h, iters, C.byref(p), variant))

    def last_solver_variant(self):
        return self._lib.fb_last_solver_variant(self._h)

    def last_solver_transport(self):
        """Variant 3: 1 = cluster (DSMEM), 2 = L2 mailboxes."""
        return self._lib.fb_last_solver_transport(self._h)

    def last_cluster_size(self):
        return self._lib.fb_last_cluster_size(self._h)

    def costs(self, stream, data_factor=0.15):
        s, d = C.c_double(0), C.c_double(0)
        self._ck(self._lib.fb_costs(self._h, stream, data_factor, C.byref(s), C.byref(d)))
        return s.value, d.value

    # ------------------------------------------------------------------ frames / features
    def frame_set(self, stream, slot, gray, pose):
        assert gray.dtype == np.uint8 and gray.ndim == 2 and gray.shape == (self.H, self.W)
        assert gray.strides[1] == 1
        pose = _f32(pose)
        self._ck(self._lib.fb_frame_set(self._h, stream, slot, _ptr(gray), gray.strides[0], _ptr(pose)))

    def frame_pose_set(self, stream, slot, pose):
        pose = _f32(pose)
        self._ck(self._lib.fb_frame_pose_set(self._h, stream, slot, _ptr(pose)))

    def pool_reserve(self, n):
        self._ck(self._lib.fb_pool_reserve(self._h, n))

    def pool_upload(self, idx, gray):
        gray = np.ascontiguousarray(gray, dtype=np.uint8)
        self._ck(self._lib.fb_pool_upload(self._h, idx, _ptr(gray), gray.strides[0]))

    def frame_from_pool(self, stream, slot, idx, pose):
        pose = _f32(pose)
        self._ck(self._lib.fb_frame_from_pool(self._h, stream, slot, idx, _ptr(pose)))

    def features_set(self, stream, u_ref, ref_slot, mu, var, dropouts=None, alive=None):
        u_ref, ref_slot, mu, var = _f32(u_ref), _i32(ref_slot), _f32(mu), _f32(var)
        dropouts, alive = _i32(dropouts), _i32(alive)
        N = ref_slot.shape[0]
        self._ck(self._lib.fb_features_set(self._h, stream, N, _ptr(u_ref), _ptr(ref_slot), _ptr(mu),
                                           _ptr(var), _ptr(dropouts), _ptr(alive)))
        self._nF[stream] = N

    def features_get(self, stream):
        N = self._nF[stream]
        out = dict(mu=np.zeros(N, np.float32), var=np.zeros(N, np.float32),
                   dropouts=np.zeros(N, np.int32), alive=np.zeros(N, np.int32),
                   status=np.zeros(N, np.int32), u_cmp=np.zeros((N, 2), np.float32))
        self._ck(self._lib.fb_features_get(self._h, stream, _ptr(out["mu"]), _ptr(out["var"]),
                                           _ptr(out["dropouts"]), _ptr(out["alive"]),
                                           _ptr(out["status"]), _ptr(out["u_cmp"])))
        return out

    def features_reinit(self, ref_slot, mu0, var0):
        ref_slot = _i32(np.broadcast_to(np.asarray(ref_slot, np.int32), (self.S,)))
        self._ck(self._lib.fb_features_reinit(self._h, _ptr(ref_slot), mu0, var0))

    def idepth_update(self, cmp_slot):
        cmp_slot = _i32(np.broadcast_to(np.asarray(cmp_slot, np.int32), (self.S,)))
        self._ck(self._lib.fb_idepth_update(self._h, _ptr(cmp_slot)))

    def idepth_counters(self, stream):
        c = np.zeros(NUM_COUNTERS, np.int32)
        self._ck(self._lib.fb_idepth_counters(self._h, stream, _ptr(c)))
        return c

    def project_features(self, stream, cur_slot):
        N = self._nF[stream]
        u = np.zeros((N, 2), np.float32)
        mu = np.zeros(N, np.float32)
        var = np.zeros(N, np.float32)
        valid = np.zeros(N, np.int32)
        self._ck(self._lib.fb_project_features(self._h, stream, cur_slot, _ptr(u), _ptr(mu), _ptr(var), _ptr(valid)))
        return u, mu, var, valid

    # ------------------------------------------------------------------ assembly / interpolation
    def graph_bind_features(self, stream, vertex_feature):
        vf = _i32(vertex_feature)
        assert vf.shape[0] == self._nV[stream]
        self._ck(self._lib.fb_graph_bind_features(self._h, stream, _ptr(vf)))

    def graph_data_from_features(self, adaptive_weights=False):
        self._ck(self._lib.fb_graph_data_from_features(self._h, 1 if adaptive_weights else 0))

    def mesh_set(self, stream, tris):
        tris = _i32(tris)
        self._ck(self._lib.fb_mesh_set(self._h, stream, tris.shape[0], _ptr(tris)))
        self._nT[stream] = tris.shape[0]

    def interpolate(self, stream, filter_params=None):
        out = np.zeros((self.H, self.W), np.float32)
        # sized for the context's triangle capacity (2 * max_vertices): after fb_update the triangle count lives
        # in the library, not in this mirror
        valid = np.zeros(2 * self.max_vertices, np.uint8)
        fp = C.byref(filter_params) if filter_params is not None else None
        self._ck(self._lib.fb_interpolate(self._h, stream, fp, _ptr(out), _ptr(valid)))
        nt = self._nT[stream]
        if nt == 0:
            try:
                nt = int(self.get_stat(stream, "num_tris"))
            except Exception:
                nt = 0
        return out, valid[:nt]

    # ------------------------------------------------------------------ flame::Flame::update + getters
    def set_update_params(self, p):
        self._ck(self._lib.fb_set_update_params(self._h, C.byref(p)))

    def update(self, stream, time, img_id, pose, gray, is_poseframe):
        """flame::Flame::update: returns True when the mesh / depth outputs were refreshed."""
        assert gray.dtype == np.uint8 and gray.shape == (self.H, self.W) and gray.strides[1] == 1
        pose = _f32(pose)
        rc = self._lib.fb_update(self._h, stream, float(time), int(img_id), _ptr(pose), _ptr(gray),
                                 gray.strides[0], 1 if is_poseframe else 0)
        if rc < 0:
            self._ck(rc)
        return rc == 1

    def get_mesh(self, stream, filter_params=None):
        V, T, E = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        self._ck(self._lib.fb_get_mesh_sizes(self._h, stream, C.byref(V), C.byref(T), C.byref(E)))
        V, T, E = V.value, T.value, E.value
        out = dict(vtx=np.zeros((V, 2), np.float32), idepth=np.zeros(V, np.float32),
                   normals=np.zeros((V, 3), np.float32), tris=np.zeros((T, 3), np.int32),
                   tri_valid=np.zeros(T, np.uint8), edges=np.zeros((E, 2), np.int32))
        if V == 0:
            return out
        fp = C.byref(filter_params) if filter_params is not None else None
        self._ck(self._lib.fb_get_mesh(self._h, stream, fp, _ptr(out["vtx"]), _ptr(out["idepth"]),
                                       _ptr(out["normals"]), _ptr(out["tris"]), _ptr(out["tri_valid"]),
                                       _ptr(out["edges"])))
        return out

    def delaunay_device(self, stream, pts):
        """fb_delaunay_device: the update pipeline's device triangulation of an arbitrary point set.
        Returns (tris [T,3], edges [E,2]) like delaunay()."""
        pts = _f32(pts)
        n = pts.shape[0]
        tris = np.zeros((max(2 * n, 1), 3), np.int32)
        edges = np.zeros((max(3 * n, 1), 2), np.int32)
        nt, ne = C.c_int32(0), C.c_int32(0)
        rc = self._lib.fb_delaunay_device(self._h, stream, n, _ptr(pts), _ptr(tris), C.byref(nt), _ptr(edges), C.byref(ne))
        if rc != 0:
            raise FlameError("fb_delaunay_device: rc %d (%s)" % (rc, self._lib.fb_last_error(self._h).decode()))
        return tris[:nt.value].copy(), edges[:ne.value].copy()

    def get_idepthmap(self, stream, filter_params=None, out=None):
        """getInverseDepthMap / getFilteredInverseDepthMap; `out` may be a (pinned) float32 [H,W] array."""
        if out is None:
            out = np.zeros((self.H, self.W), np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.size == self.H * self.W
        fp = C.byref(filter_params) if filter_params is not None else None
        self._ck(self._lib.fb_get_idepthmap(self._h, stream, fp, _ptr(out)))
        return out

    def get_raw_idepths(self, stream):
        n = C.c_int32(0)
        xy = np.zeros((self.max_features, 2), np.float32)
        mu = np.zeros(self.max_features, np.float32)
        var = np.zeros(self.max_features, np.float32)
        self._ck(self._lib.fb_get_raw_idepths(self._h, stream, C.byref(n), _ptr(xy), _ptr(mu), _ptr(var)))
        return xy[:n.value].copy(), mu[:n.value].copy(), var[:n.value].copy()

    def update_run(self, stream, k0, k1, frames, poses, poseframe_every, filter_params, out_map):
        """fb_update_run: frames [k0, k1) of a (n, H, W) uint8 array through update + filtered map, looped in C."""
        assert frames.dtype == np.uint8 and frames.flags["C_CONTIGUOUS"] and poses.dtype == np.float32 and poses.flags["C_CONTIGUOUS"]
        fp = C.byref(filter_params) if filter_params is not None else None
        n = self._lib.fb_update_run(self._h, stream, k0, k1, _ptr(frames), frames.shape[1] * frames.shape[2], _ptr(poses),
                                    poseframe_every, fp, _ptr(out_map))
        if n < 0:
            self._ck(n)
        return n

    def get_stat(self, stream, key):
        v = C.c_double(0)
        self._ck(self._lib.fb_get_stat(self._h, stream, key.encode(), C.byref(v)))
        return v.value

    def update_poseframe_poses(self, stream, img_ids, poses):
        ids, poses = _i32(img_ids), _f32(poses)
        self._ck(self._lib.fb_update_poseframe_poses(self._h, stream, ids.shape[0], _ptr(ids), _ptr(poses)))

    def prune_poseframes(self, stream, img_ids_to_keep):
        ids = _i32(img_ids_to_keep)
        self._ck(self._lib.fb_prune_poseframes(self._h, stream, ids.shape[0], _ptr(ids)))

    def frame_gradient(self, stream, slot):
        out = np.zeros((self.H, self.W), np.float32)
        self._ck(self._lib.fb_frame_gradient(self._h, stream, slot, _ptr(out)))
        return out

    def frame_pyr_down(self, stream, slot):
        out = np.zeros((self.H // 2, self.W // 2), np.uint8)
        self._ck(self._lib.fb_frame_pyr_down(self._h, stream, slot, _ptr(out)))
        return out

    def detect(self, stream, slot, win, border, min_grad_mag, occupied=None):
        cells = (self.W // win) * (self.H // win)
        xy = np.zeros((cells, 2), np.float32)
        ok = np.zeros(cells, np.int32)
        n = C.c_int32(0)
        occ = None if occupied is None else np.ascontiguousarray(occupied, np.uint8).ravel()
        self._ck(self._lib.fb_detect(self._h, stream, slot, win, border, min_grad_mag, _ptr(occ), _ptr(xy),
                                     _ptr(ok), C.byref(n)))
        return n.value, xy, ok

    def get_feature_pool(self, stream):
        N = self.max_features
        out = dict(u_ref=np.zeros((N, 2), np.float32), ref_slot=np.zeros(N, np.int32), mu=np.zeros(N, np.float32),
                   var=np.zeros(N, np.float32), dropouts=np.zeros(N, np.int32), alive=np.zeros(N, np.int32))
        self._ck(self._lib.fb_get_feature_pool(self._h, stream, _ptr(out["u_ref"]), _ptr(out["ref_slot"]),
                                               _ptr(out["mu"]), _ptr(out["var"]), _ptr(out["dropouts"]),
                                               _ptr(out["alive"])))
        return out

    def hotpath_step(self, desc):
        """desc: a StepDesc whose pointer fields the caller keeps alive."""
        self._ck(self._lib.fb_hotpath_step(self._h, C.byref(desc)))

    def pipeline_join(self):
        self._ck(self._lib.fb_pipeline_join(self._h))

    def results_wait(self, lag=0):
        self._ck(self._lib.fb_results_wait(self._h, lag))

    # ------------------------------------------------------------------ profiling
    def profile_enable(self, on=True):
        """True / False: all sections on / off; an int >= 2 = section (on - 2) only."""
        self._ck(self._lib.fb_profile_enable(self._h, int(on) if not isinstance(on, bool) else (1 if on else 0)))

    def profile_reset(self):
        self._ck(self._lib.fb_profile_reset(self._h))

    def profile_get(self, section):
        ms, calls, launches = C.c_float(0), C.c_int64(0), C.c_int64(0)
        self._ck(self._lib.fb_profile_get(self._h, section, C.byref(ms), C.byref(calls), C.byref(launches)))
        return ms.value, calls.value, launches.value

    def launch_count(self):
        return self._lib.fb_launch_count(self._h)
