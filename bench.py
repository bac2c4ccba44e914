#!/usr/bin/env python
"""bench.py -- FLaME hot path on B200: frames/s (and solver iterations/s) on synthetic 640x480 streams.

A "step" is one frame of every stream of the batch through the hot path:
    epipolar inverse-depth update of ~5k features  ->  data-term assembly  ->  50 NLTGV2-L1
    primal-dual iterations on the 5k-vertex Delaunay graph (warm-started), every 5th step a new
    poseframe re-initialises the feature filters (flame_ros_b200/workload.py).

  python bench.py --gpus N --steps K --warmup W            GPU arm (one process per GPU under torchrun)
  python bench.py --impl reference ...                     CPU arm: the oracle restatement of the
                                                           reference's CPU algorithm on all host cores

Prints ONE JSON line (rank 0).  `value` = whole-job frames/s with inputs resident in HBM (CUDA-event
timed, L2 flushed between steps); `e2e` = the same through the C-ABI with HOST buffers (pinned H2D of
each frame, D2H of the vertex inverse depths, host wall clock per step).  See DESIGN.md "Measurement".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from flame_ros_b200 import workload as WL  # noqa: E402

METRIC = "frames_per_second"
UNIT = "frames/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=8, help="independent streams per GPU (batched launch)")
    ap.add_argument("--config", default="C2", choices=list(WL.CONFIGS))
    ap.add_argument("--variant", type=int, default=0, help="solver variant: 0 auto, 1 streaming, 2 cluster, 3 grid-resident")
    ap.add_argument("--no-single", action="store_true", help="skip the extra single-stream measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-update", action="store_true", help="skip the full-pipeline (fb_update) leg")
    ap.add_argument("--no-c4", action="store_true", help="skip the C4 (1280x720, 20k vertices, 100 iterations) block")
    ap.add_argument("--total-streams", type=int, default=8,
                    help="strong-scaling leg (BASELINE configs[4] as written): this many streams in total, stream s -> rank s mod N")
    ap.add_argument("--update-streams", type=int, default=8, help="independent flame::Flame instances per GPU in the e2e_update leg")
    ap.add_argument("--update-frames", type=int, default=36, help="timed frames per stream in the e2e_update leg")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def roofline_traffic(key):
    """DRAM bytes per launch of the dominant kernel for workload `key` ("C2x8", "C2x1", "C4x1"), from the
    `ncu --set full` capture of that workload (profiles/roofline_traffic.json: {key: {bytes, profile}});
    None when that workload has not been captured -- a number is never reused across workloads."""
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        ent = json.load(open(tp)).get(key)
        return int(ent["dram_bytes_per_launch"]) if ent else None
    except Exception:
        return None


def dram_gbs(traffic, launch_us):
    """Physical DRAM rate of the launch (measured bytes / measured duration), next to the algorithmic one."""
    return (traffic / (launch_us * 1e-6) / 1e9) if (traffic and launch_us > 0) else None


# --------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples the SM clock and the throttle reasons every few ms while the timed legs run (NVML via
    nvidia_ml_py; falls back to `nvidia-smi -lms` when NVML cannot be loaded)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_s=0.003):
        super().__init__(daemon=True)
        self.index, self.period, self.stop_flag, self.proc = index, period_s, False, None
        self.sm, self.sm_max, self.reasons = [], None, set()

    def _run_nvml(self):
        import pynvml as N
        N.nvmlInit()
        h = N.nvmlDeviceGetHandleByIndex(self.index)
        self.sm_max = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
        names = {getattr(N, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                 getattr(N, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                 getattr(N, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                 getattr(N, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap"}
        get_reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            self.sm.append(float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)))
            r = int(get_reasons(h))
            for bit, name in names.items():
                if r & bit:
                    self.reasons.add(name)
            time.sleep(self.period)
        N.nvmlShutdown()

    def _run_smi(self):
        self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits", "-lms", "100"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            if self.stop_flag:
                break
            c = [v.strip() for v in line.split(",")]
            if len(c) >= 6 and c[0].replace(".", "").isdigit():
                self.sm.append(float(c[0]))
                self.sm_max = float(c[1])
                self.reasons |= {names[k] for k in range(4) if c[2 + k] == "Active"}

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            try:
                self._run_smi()
            except Exception:
                pass

    def stop(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2.0)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


# --------------------------------------------------------------------------------------- GPU arm
BASELINE_CONFIG = {"tiny": "smoke-sized, not a BASELINE config", "C2": "configs[1]", "C3": "configs[2]", "C4": "configs[3]"}
RESULT_LAG = 3            # e2e leg: steps between a step's enqueue and the host's read of its result
RESULT_BUFFERS = RESULT_LAG + 1


class GpuRun:
    """One context with S streams of the workload; runs steps in `resident` or `e2e` mode."""

    def __init__(self, capi, datas, device, cuda_stream, variant):
        d0 = datas[0]
        self.capi, self.datas, self.S, self.variant = capi, datas, len(datas), variant
        V = max(d.V for d in datas)
        E = max(d.E for d in datas)
        self.ctx = capi.Context(self.S, d0.W, d0.H, WL.N_SLOTS, V, V, E, device=device, cuda_stream=cuda_stream)
        self.params = capi.default_nltgv2_params()
        self.iters = int(os.environ.get("FB_BENCH_PD_ITERS", d0.iters))   # diagnosis only: host-bound floor of the legs
        ctx = self.ctx
        # device frame pool, replicated so that the pool the resident leg cycles through is larger
        # than the 126 MB L2 (a frame is re-read only after > L2 bytes of other frames went by)
        fbytes = d0.W * d0.H
        self.pool_copies = max(1, -(-(160 << 20) // (self.S * WL.POOL_FRAMES * fbytes)))
        self.pool_copies = 2 * -(-self.pool_copies // 2)   # schedule period (10 x copies) divisible by the 4 result buffers
        ctx.pool_reserve(self.pool_copies * self.S * WL.POOL_FRAMES)
        for s, d in enumerate(datas):
            ctx.set_intrinsics(s, d.K)
            for r in range(self.pool_copies):
                for k in range(WL.POOL_FRAMES):
                    ctx.pool_upload((r * self.S + s) * WL.POOL_FRAMES + k, d.frames[k])
            ctx.graph_set(s, d.u_ref, d.edges, d.alpha, d.beta)
            z0 = np.full(d.V, WL.MU0, np.float32)
            ctx.graph_data_set(s, z0)
            ctx.graph_state_set(s)
            ctx.graph_bind_features(s, np.arange(d.V, dtype=np.int32))
            ctx.features_set(s, d.u_ref, np.zeros(d.V, np.int32), z0, np.full(d.V, WL.VAR0, np.float32))
        # e2e leg: the camera frames sit in pinned host memory (as a capture driver would deliver
        # them) and the vertex idepths come back into a pinned buffer
        # frame-major: the S frames of one step are contiguous (one multi-camera capture buffer), so a
        # step's upload is a single transfer
        self.h_frames = capi.PinnedBuffer((WL.POOL_FRAMES, self.S, d0.H, d0.W), np.uint8)
        for s, d in enumerate(datas):
            np.copyto(self.h_frames.array[:, s], d.frames)
        self.h_x = capi.PinnedBuffer((RESULT_BUFFERS, self.S, V), np.float32)   # ring of result buffers
        self.maxV = V
        self.cmp = np.full(self.S, WL.CMP_SLOT, np.int32)
        ctx.sync()

    def close(self):
        self.ctx.sync()
        for b in (self.h_frames, self.h_x):
            b.free()
        self.ctx.close()

    def _descs(self):
        """One prebuilt fb_step_desc per (schedule phase, mode): the timed loop is a single C call per step."""
        import ctypes as C
        capi, S = self.capi, self.S
        self._keep = []
        table = {}
        period = 2 * WL.EPOCH * self.pool_copies
        for k in range(period):
            newpf, ref_slot, ref_idx, cmp_idx = WL.schedule(k)
            r = (k // (2 * WL.EPOCH)) % self.pool_copies      # which replica of the pool this step reads
            ref_poses = np.ascontiguousarray(np.stack([d.poses[ref_idx] for d in self.datas]), np.float32)
            cmp_poses = np.ascontiguousarray(np.stack([d.poses[cmp_idx] for d in self.datas]), np.float32)
            ref_pool = np.array([(r * S + s) * WL.POOL_FRAMES + ref_idx for s in range(S)], np.int32)
            cmp_pool = np.array([(r * S + s) * WL.POOL_FRAMES + cmp_idx for s in range(S)], np.int32)
            ref_ptr = (C.c_void_p * S)(*[self.h_frames.array[ref_idx, s].ctypes.data for s in range(S)])
            cmp_ptr = (C.c_void_p * S)(*[self.h_frames.array[cmp_idx, s].ctypes.data for s in range(S)])
            self._keep += [ref_poses, cmp_poses, ref_pool, cmp_pool, ref_ptr, cmp_ptr]
            for mode in ("resident", "e2e_sync", "e2e_pipe"):
                d = capi.StepDesc()
                d.new_poseframe, d.ref_slot, d.cmp_slot = int(newpf), ref_slot, WL.CMP_SLOT
                if mode != "resident":
                    d.ref_images = C.cast(ref_ptr, C.POINTER(C.c_void_p))
                    d.cmp_images = C.cast(cmp_ptr, C.POINTER(C.c_void_p))
                    if not os.environ.get("FB_BENCH_NO_XOUT"):   # diagnosis only
                        d.x_out = self.h_x.array[k % RESULT_BUFFERS].ctypes.data_as(C.POINTER(C.c_float))
                if mode in ("e2e_pipe", "resident"):
                    d.pipelined = 1
                    d.cmp_slot = WL.CMP_SLOT + (k % 2)   # frames alternate between two slots
                d.ref_pool_idx = ref_pool.ctypes.data_as(C.POINTER(C.c_int32))
                d.cmp_pool_idx = cmp_pool.ctypes.data_as(C.POINTER(C.c_int32))
                d.ref_poses = ref_poses.ctypes.data_as(C.POINTER(C.c_float))
                d.cmp_poses = cmp_poses.ctypes.data_as(C.POINTER(C.c_float))
                d.mu0, d.var0, d.adaptive_weights = WL.MU0, WL.VAR0, 0
                # the poseframe of this epoch is the frame of the previous step, already on the device
                # (table entry 0 also serves the very first call, which has no previous frame: it uploads)
                if newpf and k > 0:
                    d.ref_from_slot = 1 + (WL.CMP_SLOT + ((k - 1) % 2) if d.pipelined else WL.CMP_SLOT)
                d.iters, d.variant, d.rparams = self.iters, self.variant, self.params
                table[(k, mode)] = d
        return table, period

    def step(self, k, mode):
        """One frame of every stream.  resident: frames come from the device pool, no synchronisation.
        e2e_sync: pinned host frames uploaded in-stream, vertex idepths read back, blocking.
        e2e_pipe: the same through the pipelined API (uploads on a copy stream overlap the previous
        frame's kernels; results land asynchronously in a pinned double buffer)."""
        if not hasattr(self, "_table"):
            self._table, self._period = self._descs()
        self.ctx.hotpath_step(self._table[(k % self._period, mode)])

    def consume(self, k):
        """Touch the result of step k (the application's read of the vertex inverse depths)."""
        return float(self.h_x.array[k % RESULT_BUFFERS, 0, 0])

    def bytes_per_step(self):
        d = self.datas[0]
        h2d = self.S * (d.W * d.H + 28)   # one frame + pose per stream; poseframes are frames already uploaded
        d2h = self.S * self.maxV * 4
        return int(h2d), int(d2h)


def run_gpu_leg(torch, run, steps, warmup, flush_buf, mode, barrier, min_region_s=0.05):
    """Returns (seconds over `steps` timed steps, launches in ONE timed block, solver ms, solver calls, blocks).
    A block of `steps` steps is short (2 ms at the driver's --steps 20): when it is below `min_region_s`
    the block is repeated and the MEDIAN block time is reported (steps stays what was asked for)."""
    ctx = run.ctx
    for k in range(warmup):
        run.step(k, mode)
    ctx.sync()
    ctx.profile_enable(2 + run.capi.PROF_SOLVE)   # event pairs around every solver launch only (markers: nothing blocks)
    ctx.profile_reset()
    barrier()
    torch.cuda.synchronize()
    l0 = ctx.launch_count()
    launches_block = None
    times = []
    k_next = warmup

    def one_block(k0):
        if mode == "resident":
            # K frames back to back from the device pool (> L2, so no frame is L2-resident when re-read);
            # one CUDA event before the first and one after the last step on the launching stream, which
            # is joined on the device with the library's auxiliary streams first
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                run.step(k0 + i, mode)
            ctx.pipeline_join()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) * 1e-3
        if mode == "e2e_pipe":
            # K steps back to back; every step uploads its frames from pinned host memory and its vertex
            # idepths are read on the host RESULT_LAG steps later (a ring of RESULT_LAG + 1 pinned buffers),
            # like a streaming consumer.  From the enqueue of a step to its result the chain is upload (52 us)
            # -> epipolar update -> assembly -> solve -> read-back, ~220 us: with a lag of 2 the host could
            # not start the next upload early enough and the upload sat on the critical path (90 us per step).
            t0 = time.perf_counter()
            for i in range(steps):
                k = k0 + i
                run.step(k, mode)
                if i >= RESULT_LAG:
                    ctx.results_wait(RESULT_LAG)
                    run.consume(k - RESULT_LAG)
            for lag in range(min(RESULT_LAG, steps) - 1, -1, -1):
                ctx.results_wait(lag)
                run.consume(k0 + steps - 1 - lag)
            torch.cuda.synchronize()
            return time.perf_counter() - t0
        total = 0.0
        for i in range(steps):
            k = k0 + i
            flush_buf.zero_()                      # evict L2 between timed steps (outside the timed bracket)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            run.step(k, mode)                      # ends with a synchronising D2H
            run.consume(k)
            total += time.perf_counter() - t0
        torch.cuda.synchronize()
        return total

    spent = 0.0
    while True:
        t = one_block(k_next)
        k_next += steps
        if launches_block is None:
            launches_block = ctx.launch_count() - l0
        times.append(t)
        spent += t
        if mode == "e2e_sync" or spent >= min_region_s or len(times) >= 25:
            break
    barrier()
    ms, calls, _ = ctx.profile_get(run.capi.PROF_SOLVE)
    ctx.profile_enable(False)
    return float(np.median(times)), launches_block, ms, calls, len(times)


PINNING = {"mode": "none"}
HOST_CORES_TOTAL = [None]   # cores of the process before any pinning (what the JSON line reports as host.cores)


def pin_rank_to_its_cores(local_rank, world):
    """Several ranks on one host: every rank gets its own share of the cores NVML names as local to its GPU
    (same socket / NUMA node), before any buffer is allocated or any thread started -- pinned host buffers
    then sit in local memory and the ranks' host threads do not migrate over each other.  Best effort:
    silently skipped when NVML or sched_setaffinity is unavailable."""
    HOST_CORES_TOTAL[0] = host_cores()
    if world <= 1 or os.environ.get("FB_BENCH_NO_PIN"):
        return
    try:
        import pynvml
        pynvml.nvmlInit()
        avail = sorted(os.sched_getaffinity(0))
        nwords = (max(avail) + 64) // 64

        def cpus_of(dev):
            words = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(dev), nwords)
            return tuple(i for i in avail if (words[i // 64] >> (i % 64)) & 1)

        mine_all = cpus_of(local_rank)
        peers = [r for r in range(world) if cpus_of(r) == mine_all]
        share = len(mine_all) // len(peers)
        if share < 1:
            return
        k = peers.index(local_rank)
        mine = mine_all[k * share:(k + 1) * share]
        os.sched_setaffinity(0, mine)
        PINNING.update(mode="nvml", cores_per_rank=len(mine), gpu_local_cores=len(mine_all))
    except Exception as exc:   # no NVML, no permission, odd topology: run unpinned
        PINNING.update(mode="none", reason=str(exc)[:80])


def gpu_main(args):
    rank, local_rank, world = dist_env()
    pin_rank_to_its_cores(local_rank, world)
    import torch
    import torch.distributed as dist
    from flame_ros_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner there at NCCL_DEBUG=VERSION
        # (this image's default) and at WARN; INFO/TRACE asked for by the user is left alone
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            del os.environ["NCCL_DEBUG"]
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from flame_ros_b200 import sharding
    red = sharding.Reducer(dist if world > 1 else None, device="cuda")
    barrier, max_over_ranks, sum_over_ranks = red.barrier, red.max, red.sum

    capi.load_library()
    S = args.streams
    datas = [WL.StreamData(args.config, seed=sid) for sid in sharding.stream_ids(rank, world, S)]   # weak: S streams per GPU
    # a dedicated (non-default) torch stream: the library enqueues on it and torch's events see it
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream_ptr = tstream.cuda_stream
    assert stream_ptr != 0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    peak, peak_src = peaks()

    # host->device link rate of this box (pinned memory, 64 MiB, best of 5): the e2e leg uploads every
    # frame, so frames/s x bytes/frame cannot exceed it -- reported next to e2e as its own roofline
    hbuf = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
    dbuf = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    link_gbs = 0.0
    for _ in range(5):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        dbuf.copy_(hbuf, non_blocking=True)
        ev1.record()
        torch.cuda.synchronize()
        link_gbs = max(link_gbs, (64 << 20) / (ev0.elapsed_time(ev1) * 1e-3) / 1e9)
    del hbuf, dbuf

    run = GpuRun(capi, datas, local_rank, stream_ptr, args.variant)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()   # samples while the three timed legs below run (set-up is excluded)
    t_res, launches, solve_ms, solve_calls, blocks_res = run_gpu_leg(torch, run, args.steps, args.warmup, flush, "resident", barrier)
    variant_used = run.ctx.last_solver_variant()
    cluster_size = run.ctx.last_cluster_size()
    transport = run.ctx.last_solver_transport()
    t_sync, _, _, _, _ = run_gpu_leg(torch, run, args.steps, args.warmup, flush, "e2e_sync", barrier)
    t_e2e, _, _, _, blocks_e2e = run_gpu_leg(torch, run, args.steps, args.warmup, flush, "e2e_pipe", barrier)
    h2d, d2h = run.bytes_per_step()
    run_pool_mib = run.pool_copies * S * WL.POOL_FRAMES * datas[0].W * datas[0].H >> 20
    alg_bytes = sum(d.algorithmic_bytes_per_iter() for d in datas) * datas[0].iters
    iters = datas[0].iters
    run.close()
    clocks = sampler.stop() if sampler else None

    single = None
    if not args.no_single and S > 1:
        run1 = GpuRun(capi, datas[:1], local_rank, stream_ptr, args.variant)
        t1, _, ms1, c1, _ = run_gpu_leg(torch, run1, args.steps, args.warmup, flush, "resident", barrier)
        t1s, _, _, _, _ = run_gpu_leg(torch, run1, args.steps, args.warmup, flush, "e2e_sync", barrier)
        t1e, _, _, _, _ = run_gpu_leg(torch, run1, args.steps, args.warmup, flush, "e2e_pipe", barrier)
        run1.close()
        t1, t1e, t1s = max_over_ranks(t1), max_over_ranks(t1e), max_over_ranks(t1s)
        single = {"streams_per_gpu": 1, "value": world * args.steps / t1, "e2e": world * args.steps / t1e,
                  "e2e_sync": world * args.steps / t1s,
                  "ms_per_step": 1e3 * t1 / args.steps, "unit": UNIT,
                  "solver_us_per_frame": 1e3 * ms1 / max(c1, 1),
                  "roofline_frac": (datas[0].algorithmic_bytes_per_iter() * iters / (1e6 * ms1 / max(c1, 1))) / peak}

    upd = None
    if not args.no_update:
        u = update_leg_gpu(capi, args, rank, local_rank, barrier)
        tu_all, tu_one = max_over_ranks(u["t_all"]), max_over_ranks(u["t_one"])
        launches += u["launches"]
        upd = {"value": world * u["S"] * u["K"] / tu_all, "unit": UNIT, "streams_per_gpu": u["S"],
               "frames_per_stream": u["K"], "warmup_frames": WL.UPD_WARMUP,
               "ms_per_frame_per_stream": 1e3 * tu_all / u["K"],
               "single_stream": {"value": world * u["K"] / tu_one, "ms_per_frame": 1e3 * tu_one / u["K"], "unit": UNIT},
               "h2d_bytes_per_frame": u["h2d"], "d2h_bytes_per_frame": u["d2h"],
               "vertices": u["stats"].get("vertices"), "triangles": u["stats"].get("tris"),
               "coverage": u["stats"].get("coverage"),
               "solver_variant": {1: "streaming", 4: "plan-free resident (cluster, L2 exchange)",
                                  5: "tile-resident, planned on the device (cluster of 16, DSMEM exchange)"}.get(u["stats"].get("variant"), u["stats"].get("variant")),
               "call": "per frame and stream (loop in C, fb_update_run): fb_update(host gray image, pose, is_poseframe every 6th) = flame::Flame::update "
                       "(/root/reference/src/flame_nodelet.cc:634), then fb_get_idepthmap(filter) = getFilteredInverseDepthMap "
                       "(:682-683) into pinned host memory; one context + one host thread per stream; detection window 8 px "
                       "(~5.5k Delaunay vertices), 50 PD iterations per frame, topology rebuilt on the device every frame"}
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            upd["cpu_baseline"] = cpu_update_baseline(u["datas"], frames=u["K"])   # the same frames as the GPU leg
            upd["vs_cpu_baseline"] = upd["value"] / upd["cpu_baseline"]["value"]
        del u

    # ---- BASELINE configs[4] as written: a fixed batch of --total-streams streams, stream s -> rank s mod N
    strong = None
    T = args.total_streams
    my_ids = sharding.strong_stream_ids(rank, world, T) if T >= world else []
    if world > 1 and my_ids:
        sdatas = [WL.StreamData(args.config, seed=sid) for sid in my_ids]
        runs = GpuRun(capi, sdatas, local_rank, stream_ptr, args.variant)
        ts, _, mss, cs, _ = run_gpu_leg(torch, runs, args.steps, args.warmup, flush, "resident", barrier)
        tse, _, _, _, _ = run_gpu_leg(torch, runs, args.steps, args.warmup, flush, "e2e_pipe", barrier)
        runs.close()
        ts, tse = max_over_ranks(ts), max_over_ranks(tse)
        strong = {"scaling": "strong", "streams_total": T, "streams_per_gpu": len(my_ids), "value": T * args.steps / ts,
                  "e2e": T * args.steps / tse, "unit": UNIT, "ms_per_step": 1e3 * ts / args.steps,
                  "solver_us_per_launch": 1e3 * mss / max(cs, 1),
                  "note": "BASELINE configs[4] / SURVEY 8(e) as written: %d streams in total, stream s -> rank s mod %d; with one "
                          "stream per GPU a step is latency-bound (one cluster of CTAs busy), so the curve flattens" % (T, world)}
    elif world == 1:
        strong = {"scaling": "strong", "streams_total": S, "streams_per_gpu": S, "value": None, "unit": UNIT,
                  "note": "at 1 GPU the strong-scaling point is the headline itself (%d streams on one GPU)" % S}

    # ---- C4 (BASELINE configs[3]: 1280x720, 20k vertices, 100 PD iterations, one stream): its own block
    c4 = None
    if not args.no_c4 and args.config != "C4":
        d4 = [WL.StreamData("C4", seed=5000 + rank)]
        run4 = GpuRun(capi, d4, local_rank, stream_ptr, args.variant)
        t4, _, ms4, c4calls, _ = run_gpu_leg(torch, run4, args.steps, args.warmup, flush, "resident", barrier)
        v4, tr4, cl4 = run4.ctx.last_solver_variant(), run4.ctx.last_solver_transport(), run4.ctx.last_cluster_size()
        t4e, _, _, _, _ = run_gpu_leg(torch, run4, args.steps, args.warmup, flush, "e2e_pipe", barrier)
        run4.close()
        t4, t4e = max_over_ranks(t4), max_over_ranks(t4e)
        us4 = 1e3 * ms4 / max(c4calls, 1)
        alg4 = d4[0].algorithmic_bytes_per_iter() * d4[0].iters
        ach4 = alg4 / (us4 * 1e-6) / 1e9 if us4 > 0 else 0.0
        c4 = {"workload": "C4 (BASELINE configs[3]): 1280x720, %d vertices, %d edges, %d PD iterations, 1 stream per GPU" % (d4[0].V, d4[0].E, d4[0].iters),
              "value": world * args.steps / t4, "e2e": world * args.steps / t4e, "unit": UNIT, "ms_per_step": 1e3 * t4 / args.steps,
              "solver_variant": v4, "solver_transport": {1: "cluster (DSMEM)", 2: "L2 mailboxes"}.get(tr4), "ctas_per_stream": cl4,
              "roofline": {"bound": "hbm", "achieved": ach4, "peak": peak, "unit": "GB/s", "frac": ach4 / peak,
                           "algorithmic_bytes_per_launch": alg4, "launch_us": us4,
                           "traffic": roofline_traffic("C4x1"), "dram_gbs": dram_gbs(roofline_traffic("C4x1"), us4)}}
        del d4

    t_res, t_e2e, t_sync = max_over_ranks(t_res), max_over_ranks(t_e2e), max_over_ranks(t_sync)
    launches_all = int(sum_over_ranks(launches))
    frames = world * S * args.steps
    solve_ms_per_launch = solve_ms / max(solve_calls, 1)
    achieved = alg_bytes / (solve_ms_per_launch * 1e-3) / 1e9 if solve_ms_per_launch > 0 else 0.0

    out = None
    if rank == 0:
        traffic = roofline_traffic("%sx%d" % (args.config, S))
        out = {
            "metric": METRIC, "value": frames / t_res, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_res / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s (BASELINE %s): synthetic %dx%d stream, %d features = %d Delaunay vertices, "
                                   "%d PD iters/frame; x %d independent streams per GPU, one batched launch"
                                   % (args.config, BASELINE_CONFIG.get(args.config, "-"), datas[0].W, datas[0].H, datas[0].V, datas[0].V, iters, S),
                       "streams_per_gpu": S, "vertices": datas[0].V, "edges": datas[0].E, "pd_iters": iters,
                       "solver_variant": {1: "streaming (2 kernels/iter, CUDA graph)", 2: "persistent cluster (DSMEM)",
                                          3: "resident, one exchange per iteration (%s)" % {1: "cluster of CTAs, DSMEM st.async hand-over",
                                                                                               2: "cooperative grid, tagged 128-bit L2 mailboxes"}.get(transport, "?")}.get(variant_used),
                       "ctas_per_stream": cluster_size if variant_used in (2, 3) else None,
                       "l2": "value: device frame pool of %d MiB cycled through (> 126 MiB L2), no frame is cache-resident "
                             "when re-read; e2e: inputs streamed from pinned host memory; e2e_sync: 256 MiB memset "
                             "between steps" % ((run_pool_mib)),
                       "timing": "value: K steps back to back, one CUDA event before and one after on the launching stream "
                                 "(joined with the library's copy/solve streams on the device); "
                                 "e2e: host clock over the K steps run back to back through the pipelined C-ABI call "
                                 "(pinned-host frames uploaded every step on a copy stream, vertex idepths read back "
                                 "every step; inputs are streamed, never reused, so L2 is not flushed); "
                                 "e2e_sync: blocking call per step with the L2 flush between steps"},
            "solver_iters_per_second": frames * iters / t_res,
            "e2e": {"value": frames / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * t_e2e / args.steps, "mode": "pipelined",
                    "h2d_link_gbs": link_gbs,
                    "h2d_link_frac": (h2d * args.steps / t_e2e / 1e9) / link_gbs if link_gbs > 0 else None},
            "e2e_sync": {"value": frames / t_sync, "unit": UNIT, "ms_per_step": 1e3 * t_sync / args.steps,
                         "mode": "blocking call per step, L2 flushed between steps"},
            "gpu_launches": launches_all,
            "timed_blocks": {"value": blocks_res, "e2e": blocks_e2e,
                             "note": "a block = --steps steps; blocks are repeated until >= 50 ms were timed and the median block is reported; gpu_launches counts one block"},
            "host": {"cores": HOST_CORES_TOTAL[0] or host_cores(), "ranks": world, "pinning": PINNING},
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": {2: "k_nltgv2_cluster", 3: "k_nltgv2_grid"}.get(variant_used, "k_dual_edges+k_primal_vertices"),
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "dram_gbs": dram_gbs(traffic, 1e3 * solve_ms_per_launch), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "launch_us": 1e3 * solve_ms_per_launch,
                         "note": "algorithmic = (40E+64V) B/iter x iters x streams (SURVEY 8d); the persistent solvers keep the "
                                 "graph in shared memory/registers across iterations, so achieved may exceed the HBM peak "
                                 "while DRAM traffic (`traffic`) is ~1/iters of it"},
        }
        if single:
            out["single_stream"] = single
        if upd:
            out["e2e_update"] = upd
        if strong:
            out["strong_scaling"] = strong
        if c4:
            out["configs"] = {"C4": c4}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:   # the CPU baseline is an N=1 figure
        out["cpu_baseline"] = cpu_baseline(args, datas, sample_steps=3)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))



# --------------------------------------------------------------------------------------- full pipeline (fb_update)
class UpdateRun:
    """The reference-facing call: one context per camera stream (= one flame::Flame), one host thread
    each, every frame goes  host image -> fb_update (frame upload, epipolar update, projection, device
    triangulation, graph sync, 50 NLTGV2 iterations, interpolation, detection on poseframes) ->
    fb_get_idepthmap with the display filters (getFilteredInverseDepthMap) into host memory."""

    def __init__(self, capi, datas, device, triangulator=0):
        self.capi, self.datas, self.S = capi, datas, len(datas)
        d0 = datas[0]
        self.ctxs, self.h_frames, self.h_map = [], [], []
        self.filter = capi.default_tri_filter_params()
        up = capi.default_update_params()
        up.detection_win_size, up.iters, up.triangulator = d0.win, d0.iters, triangulator
        for d in datas:
            ctx = capi.Context(1, d.W, d.H, WL.UPD_N_SLOTS, WL.UPD_MAX_FEATURES, WL.UPD_MAX_VERTICES,
                               3 * WL.UPD_MAX_VERTICES, device=device)
            ctx.set_intrinsics(0, d.K)
            ctx.set_update_params(up)
            hf = capi.PinnedBuffer((d.n_frames, d.H, d.W), np.uint8)   # frames as a capture driver delivers them
            np.copyto(hf.array, d.frames)
            self.ctxs.append(ctx)
            self.h_frames.append(hf)
            self.h_map.append(capi.PinnedBuffer((d.H, d.W), np.float32))
        self.stats = [dict() for _ in datas]
        self.poses = [np.ascontiguousarray(d.poses, np.float32) for d in datas]

    def _frames(self, s, k0, k1):
        ctx, d, hf, hm = self.ctxs[s], self.datas[s], self.h_frames[s].array, self.h_map[s].array
        # the per-camera loop (fb_update, then fb_get_idepthmap with the display filters, per frame) runs in C
        # (fb_update_run), as a C++ frontend's camera thread would drive the library: with the loop in Python
        # eight camera threads spend a tenth of their time queueing for the interpreter lock
        if os.environ.get("FB_BENCH_PY_LOOP"):
            chk = 0.0
            for k in range(k0, k1):
                ok = ctx.update(0, k / 30.0, k, d.poses[k], hf[k], k % WL.UPD_POSEFRAME_EVERY == 0)
                if ok:
                    ctx.get_idepthmap(0, self.filter, out=hm)
                    chk += float(hm[d.H // 2, d.W // 2] == hm[d.H // 2, d.W // 2])   # the consumer's read
        else:
            chk = float(ctx.update_run(0, k0, k1, hf, self.poses[s], WL.UPD_POSEFRAME_EVERY, self.filter, hm))
            chk += float(hm[d.H // 2, d.W // 2] == hm[d.H // 2, d.W // 2])
        self.stats[s] = dict(vertices=ctx.get_stat(0, "num_vtx"), tris=ctx.get_stat(0, "num_tris"),
                             coverage=ctx.get_stat(0, "coverage"), variant=ctx.last_solver_variant())
        return chk

    def run(self, k0, k1, streams=None):
        """Frames [k0, k1) of the chosen streams, one host thread per stream; returns wall seconds."""
        from concurrent.futures import ThreadPoolExecutor
        streams = list(range(self.S)) if streams is None else streams
        if len(streams) == 1:
            t0 = time.perf_counter()
            self._frames(streams[0], k0, k1)
            return time.perf_counter() - t0
        with ThreadPoolExecutor(len(streams)) as ex:
            t0 = time.perf_counter()
            list(ex.map(lambda s: self._frames(s, k0, k1), streams))
            return time.perf_counter() - t0

    def launches(self):
        return sum(c.launch_count() for c in self.ctxs)

    def close(self):
        for c in self.ctxs:
            c.sync()
            c.close()
        for b in self.h_frames + self.h_map:
            b.free()


def update_leg_gpu(capi, args, rank, local_rank, barrier):
    """e2e_update: frames/s through fb_update + getFilteredInverseDepthMap, host image in, dense map out."""
    S, K = args.update_streams, args.update_frames
    # one host thread per camera: no more cameras per GPU than this rank has cores (they spin in the frame's sync)
    S = max(1, min(S, host_cores()))
    n_frames = WL.UPD_WARMUP + K
    datas = WL.update_streams(args.config, [1000 + rank * args.update_streams + s for s in range(S)], n_frames)
    run = UpdateRun(capi, datas, local_rank)
    run.run(0, WL.UPD_WARMUP)
    barrier()
    l0 = run.launches()
    t_all = run.run(WL.UPD_WARMUP, n_frames)
    launches = run.launches() - l0
    stats = run.stats[0]
    run.close()
    # one camera alone (the reference's own use: a single flame::Flame)
    run1 = UpdateRun(capi, datas[:1], local_rank)
    run1.run(0, WL.UPD_WARMUP)
    barrier()
    t_one = run1.run(WL.UPD_WARMUP, n_frames)
    run1.close()
    d0 = datas[0]
    return dict(S=S, K=K, t_all=t_all, t_one=t_one, launches=launches, stats=stats, datas=datas,
                h2d=d0.W * d0.H + 28, d2h=4 * d0.W * d0.H)


class CpuUpdateRun:
    """The same frames through the oracle's C restatement of the whole pipeline (oracle/flame_pipeline.c):
    one pipeline per stream, streams in parallel host threads x OpenMP threads inside the stages."""

    def __init__(self, datas, cores):
        from oracle import oracle as O
        self.O, self.datas, self.S = O, datas, len(datas)
        self.native = True
        try:
            O.load(native=True)
        except Exception:
            self.native = False
            O.load()
        self.par = min(self.S, cores)
        self.nthreads = max(1, cores // self.par)
        up = O.UpdateParams.default()
        up.detection_win_size, up.iters = datas[0].win, datas[0].iters
        self.filter = O.TriFilterParams.default()
        self.pipes = [O.Pipeline(d.W, d.H, d.K, WL.UPD_N_SLOTS, WL.UPD_MAX_FEATURES, WL.UPD_MAX_VERTICES, up,
                                 nthreads=self.nthreads, native=self.native) for d in datas]
        self.stage = None

    def _frames(self, s, k0, k1):
        p, d = self.pipes[s], self.datas[s]
        acc = {}
        for k in range(k0, k1):
            if p.update(k, d.poses[k], d.frames[k], k % WL.UPD_POSEFRAME_EVERY == 0):
                p.idepthmap(self.filter)
            for key, v in p.stage_ms().items():
                acc[key] = acc.get(key, 0.0) + v
        if s == 0:
            self.stage = {k: v / max(1, k1 - k0) for k, v in acc.items()}

    def run(self, k0, k1):
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(self.par) as ex:
            t0 = time.perf_counter()
            list(ex.map(lambda s: self._frames(s, k0, k1), range(self.S)))
            return time.perf_counter() - t0

    def close(self):
        for p in self.pipes:
            p.close()


def cpu_update_baseline(datas, frames):
    """Bounded sample of the e2e_update workload on the host cores: the same warm-up, then `frames`
    timed frames per stream."""
    cores = host_cores()
    run = CpuUpdateRun(datas, cores)
    run.run(0, WL.UPD_WARMUP)
    dt = run.run(WL.UPD_WARMUP, WL.UPD_WARMUP + frames)
    m = run.pipes[0].mesh()
    out = {"value": len(datas) * frames / dt, "unit": UNIT, "cores": cores, "kind": "port",
           "ms_per_frame_per_stream": 1e3 * dt / frames, "vertices": int(len(m["vtx"])),
           "stage_ms_stream0": {k: round(v, 3) for k, v in (run.stage or {}).items()},
           "sample": "%d timed frames x %d streams after the same %d warm-up frames; oracle/flame_pipeline.c "
                     "(whole pipeline incl. its own Delaunay triangulator and getFilteredInverseDepthMap; %s), %d streams in "
                     "parallel x %d OpenMP threads; this repo's restatement of FLaME, not robustrobotics/flame itself"
                     % (frames, len(datas), WL.UPD_WARMUP, "-O3 -march=native" if run.native else "portable -O2", run.par, run.nthreads)}
    run.close()
    return out

# --------------------------------------------------------------------------------------- CPU arm
class CpuRun:
    """The same schedule on the host through the oracle (oracle/flame_oracle.c): streams in parallel
    Python threads (ctypes releases the GIL), OpenMP inside each call."""

    def __init__(self, datas, cores):
        from concurrent.futures import ThreadPoolExecutor
        from oracle import oracle as O
        self.O, self.datas, self.S = O, datas, len(datas)
        self.lib_native = True
        try:
            O.load(native=True)
        except Exception:
            self.lib_native = False
            O.load()
        self.cores = cores
        self.par = min(self.S, cores)
        self.nthreads = max(1, cores // self.par)
        self.pool = ThreadPoolExecutor(self.par)
        self.ep, self.rp = O.EpiParams.default(), O.NLTGV2Params.default()
        self.st = []
        for d in datas:
            V = d.V
            z0 = np.full(V, WL.MU0, np.float32)
            self.st.append(dict(imgs=np.zeros((WL.N_SLOTS, d.H, d.W), np.uint8), poses=np.zeros((WL.N_SLOTS, 7), np.float32),
                                mu=z0.copy(), var=np.full(V, WL.VAR0, np.float32), drop=np.zeros(V, np.int32),
                                alive=np.ones(V, np.int32), ref=np.zeros(V, np.int32),
                                wt=np.ones(V, np.float32), z=z0.copy(), state=O.new_state(z0, d.E)))
            self.st[-1]["poses"][:, 3] = 1.0

    def _one(self, s, k):
        O, d, st = self.O, self.datas[s], self.st[s]
        newpf, ref_slot, ref_idx, cmp_idx = WL.schedule(k)
        if newpf:
            st["imgs"][ref_slot] = d.frames[ref_idx]
            st["poses"][ref_slot] = d.poses[ref_idx]
            st["mu"][:] = WL.MU0
            st["var"][:] = WL.VAR0
            st["drop"][:] = 0
            st["alive"][:] = 1
            st["ref"][:] = ref_slot
        st["imgs"][WL.CMP_SLOT] = d.frames[cmp_idx]
        st["poses"][WL.CMP_SLOT] = d.poses[cmp_idx]
        O.idepth_update(st["imgs"], st["poses"], d.K, WL.CMP_SLOT, st["ref"], d.u_ref, st["mu"], st["var"],
                        st["drop"], st["alive"], self.ep, nthreads=self.nthreads, native=self.lib_native)
        alive = st["alive"] == 1
        st["z"][alive] = st["mu"][alive]
        st["wt"][:] = alive.astype(np.float32)
        O.nltgv2_solve(d.u_ref, d.edges, d.alpha, d.beta, st["z"], st["wt"], st["state"], self.rp, d.iters,
                       nthreads=self.nthreads, native=self.lib_native)

    def step(self, k):
        list(self.pool.map(lambda s: self._one(s, k), range(self.S)))


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_baseline(args, datas, sample_steps):
    cores = host_cores()
    run = CpuRun(datas, cores)
    run.step(0)
    t0 = time.perf_counter()
    for k in range(1, 1 + sample_steps):
        run.step(k)
    dt = time.perf_counter() - t0
    return {"value": len(datas) * sample_steps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d steps x %d streams of the same workload after 1 warm-up step; oracle/flame_oracle.c "
                      "(-O3 -march=native%s), %d streams in parallel x %d OpenMP threads; CPU baseline is this "
                      "repo's restatement of FLaME, not robustrobotics/flame itself (source absent: parity unpinned)"
                      % (sample_steps, len(datas), "" if run.lib_native else " unavailable: portable -O2 build",
                         run.par, run.nthreads)}


def reference_main(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    S = args.streams
    datas = [WL.StreamData(args.config, seed=s) for s in range(S)]
    cores = host_cores()
    run = CpuRun(datas, cores)
    # bounded sample: cap steps so the run stays within a few minutes on any host
    steps, warmup = min(args.steps, 30), min(args.warmup, 5)
    for k in range(warmup):
        run.step(k)
    t0 = time.perf_counter()
    for k in range(warmup, warmup + steps):
        run.step(k)
    dt = time.perf_counter() - t0
    frames = S * steps
    val = frames / dt
    sample = ("%d timed steps (of %d requested) x %d streams after %d warm-up; oracle restatement, %d streams in "
              "parallel x %d OpenMP threads, %s build" % (steps, args.steps, S, warmup, run.par, run.nthreads,
                                                           "-O3 -march=native" if run.lib_native else "portable -O2"))
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
           "warmup": warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "%s (BASELINE %s) x %d independent streams, CPU oracle restatement of the "
                                  "reference algorithm (robustrobotics/flame source is not in /root/reference)"
                                  % (args.config, BASELINE_CONFIG.get(args.config, "-"), S), "streams_per_gpu": S, "vertices": datas[0].V,
                      "edges": datas[0].E, "pd_iters": datas[0].iters},
           "solver_iters_per_second": frames * datas[0].iters / dt,
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    if not args.no_update:
        # the same full-pipeline workload as the GPU arm's e2e_update leg, on the host cores
        Su = args.update_streams
        frames = args.update_frames
        udatas = WL.update_streams(args.config, [1000 + s for s in range(Su)], WL.UPD_WARMUP + frames)
        out["e2e_update"] = cpu_update_baseline(udatas, frames)
        out["e2e_update"]["streams_per_gpu"] = Su
    print(json.dumps(out))


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        reference_main(a)
    else:
        gpu_main(a)
