"""fb_hotpath_step (one frame of every stream in a single C call): the device-pool, blocking-host and
pipelined-host modes must produce identical results, and they must equal the step-by-step API driven
from Python and the CPU oracle run on the same schedule."""
import ctypes as C

import numpy as np
import pytest

from flame_ros_b200 import workload as WL

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _make(capi, datas):
    d0 = datas[0]
    V, E = max(d.V for d in datas), max(d.E for d in datas)
    ctx = capi.Context(len(datas), d0.W, d0.H, WL.N_SLOTS, V, V, E)
    ctx.pool_reserve(len(datas) * WL.POOL_FRAMES)
    for s, d in enumerate(datas):
        ctx.set_intrinsics(s, d.K)
        for k in range(WL.POOL_FRAMES):
            ctx.pool_upload(s * WL.POOL_FRAMES + k, d.frames[k])
        ctx.graph_set(s, d.u_ref, d.edges, d.alpha, d.beta)
        z0 = np.full(d.V, WL.MU0, np.float32)
        ctx.graph_data_set(s, z0)
        ctx.graph_state_set(s)
        ctx.graph_bind_features(s, np.arange(d.V, dtype=np.int32))
        ctx.features_set(s, d.u_ref, np.zeros(d.V, np.int32), z0, np.full(d.V, WL.VAR0, np.float32))
    return ctx


def _run(capi, datas, mode, steps, frame_major=False):
    """frame_major: the S frames of a step are contiguous in host memory (a multi-camera capture buffer):
    the pipelined mode then uploads them as one transfer into its landing buffer."""
    S = len(datas)
    ctx = _make(capi, datas)
    V = ctx.max_vertices
    if frame_major:
        store = capi.PinnedBuffer((WL.POOL_FRAMES, S, datas[0].H, datas[0].W), np.uint8)
        frames_array = store.array.transpose(1, 0, 2, 3)   # indexed [stream, frame] like the other layout
    else:
        store = capi.PinnedBuffer((S, WL.POOL_FRAMES, datas[0].H, datas[0].W), np.uint8)
        frames_array = store.array
    frames = type("Frames", (), {"array": frames_array, "free": store.free})()
    for s, d in enumerate(datas):
        np.copyto(frames.array[s], d.frames)
    xbuf = capi.PinnedBuffer((2, S, V), np.float32)
    params = capi.default_nltgv2_params()
    keep, outs = [], []
    for k in range(steps):
        newpf, ref_slot, ref_idx, cmp_idx = WL.schedule(k)
        d = capi.StepDesc()
        d.new_poseframe, d.ref_slot, d.cmp_slot = int(newpf), ref_slot, WL.CMP_SLOT
        ref_poses = np.ascontiguousarray(np.stack([x.poses[ref_idx] for x in datas]), np.float32)
        cmp_poses = np.ascontiguousarray(np.stack([x.poses[cmp_idx] for x in datas]), np.float32)
        ref_pool = np.array([s * WL.POOL_FRAMES + ref_idx for s in range(S)], np.int32)
        cmp_pool = np.array([s * WL.POOL_FRAMES + cmp_idx for s in range(S)], np.int32)
        ref_ptr = (C.c_void_p * S)(*[frames.array[s, ref_idx].ctypes.data for s in range(S)])
        cmp_ptr = (C.c_void_p * S)(*[frames.array[s, cmp_idx].ctypes.data for s in range(S)])
        keep += [ref_poses, cmp_poses, ref_pool, cmp_pool, ref_ptr, cmp_ptr, d]
        if mode != "resident":
            d.ref_images = C.cast(ref_ptr, C.POINTER(C.c_void_p))
            d.cmp_images = C.cast(cmp_ptr, C.POINTER(C.c_void_p))
            d.x_out = xbuf.array[k % 2].ctypes.data_as(C.POINTER(C.c_float))
        if mode == "pipe":
            d.pipelined = 1
            d.cmp_slot = WL.CMP_SLOT + (k % 2)
        d.ref_pool_idx = ref_pool.ctypes.data_as(C.POINTER(C.c_int32))
        d.cmp_pool_idx = cmp_pool.ctypes.data_as(C.POINTER(C.c_int32))
        d.ref_poses = ref_poses.ctypes.data_as(C.POINTER(C.c_float))
        d.cmp_poses = cmp_poses.ctypes.data_as(C.POINTER(C.c_float))
        d.mu0, d.var0, d.adaptive_weights = WL.MU0, WL.VAR0, 0
        d.iters, d.variant, d.rparams = datas[0].iters, 0, params
        ctx.hotpath_step(d)
        if mode == "pipe":
            if k > 0:
                ctx.results_wait(1)
                outs.append(xbuf.array[(k - 1) % 2].copy())
        elif mode == "sync":
            outs.append(xbuf.array[k % 2].copy())
        else:
            outs.append(ctx.graph_x_get_all().copy())
    if mode == "pipe":
        ctx.results_wait(0)
        outs.append(xbuf.array[(steps - 1) % 2].copy())
    feats = [ctx.features_get(s) for s in range(S)]
    ctx.close()
    frames.free()
    xbuf.free()
    return outs, feats


def test_modes_agree_and_match_the_oracle(capi, oracle):
    datas = [WL.StreamData("tiny", seed=s) for s in range(2)]
    steps = 13
    res, fr = _run(capi, datas, "resident", steps)
    syn, fs = _run(capi, datas, "sync", steps)
    pip, fp = _run(capi, datas, "pipe", steps)
    for k in range(steps):
        assert np.array_equal(res[k], syn[k]), "blocking host mode differs at step %d" % k
        assert np.array_equal(res[k], pip[k]), "pipelined mode differs at step %d" % k
    for s in range(2):
        for key in ("mu", "var", "alive", "status"):
            assert np.array_equal(fr[s][key], fs[s][key]) and np.array_equal(fr[s][key], fp[s][key])
    # the same schedule on the CPU oracle (what bench.py's reference arm runs)
    for s, d in enumerate(datas):
        V = d.V
        imgs = np.zeros((WL.N_SLOTS, d.H, d.W), np.uint8)
        poses = np.zeros((WL.N_SLOTS, 7), np.float32)
        poses[:, 3] = 1
        mu, var = np.full(V, WL.MU0, np.float32), np.full(V, WL.VAR0, np.float32)
        drop, alive, ref = np.zeros(V, np.int32), np.ones(V, np.int32), np.zeros(V, np.int32)
        z, wt = mu.copy(), np.ones(V, np.float32)
        st = oracle.new_state(z, d.E)
        for k in range(steps):
            newpf, ref_slot, ref_idx, cmp_idx = WL.schedule(k)
            if newpf:
                imgs[ref_slot], poses[ref_slot] = d.frames[ref_idx], d.poses[ref_idx]
                mu[:], var[:], drop[:], alive[:], ref[:] = WL.MU0, WL.VAR0, 0, 1, ref_slot
            imgs[WL.CMP_SLOT], poses[WL.CMP_SLOT] = d.frames[cmp_idx], d.poses[cmp_idx]
            oracle.idepth_update(imgs, poses, d.K, WL.CMP_SLOT, ref, d.u_ref, mu, var, drop, alive, oracle.EpiParams.default())
            a = alive == 1
            z[a] = mu[a]
            wt[:] = a.astype(np.float32)
            oracle.nltgv2_solve(d.u_ref, d.edges, d.alpha, d.beta, z, wt, st, oracle.NLTGV2Params.default(), d.iters)
            assert np.max(np.abs(res[k][s, :V] - st["x"])) < TOL, "stream %d step %d" % (s, k)
        assert np.array_equal(fr[s]["alive"], alive) and np.max(np.abs(fr[s]["mu"] - mu)) < TOL


@pytest.mark.parametrize("env", [{}, {"FB_GEO_PINNED": "1"}, {"FB_PIPE_SINGLE_DATA": "1"}, {"FB_PIPE_SINGLE_STAGE": "1"}])
def test_pipelined_variants_are_bit_identical(capi, monkeypatch, env):
    """The pipelined step's optional mechanisms -- one-transfer upload into the landing buffer (frames of
    a step contiguous on the host), poses / slots as kernel parameters vs read from pinned memory,
    double-buffered data term, second stage on its own stream -- never change a bit of the result."""
    datas = [WL.StreamData("tiny", seed=10 + s) for s in range(3)]
    steps = 12
    res, fr = _run(capi, datas, "resident", steps)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    pip, fp = _run(capi, datas, "pipe", steps, frame_major=True)
    for k in range(steps):
        assert np.array_equal(res[k], pip[k]), "pipelined mode %s differs at step %d" % (env, k)
    for s in range(3):
        for key in ("mu", "var", "alive", "status"):
            assert np.array_equal(fr[s][key], fp[s][key])
