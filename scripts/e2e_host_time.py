"""Where the e2e leg's step goes: host time of the enqueue call and of the result wait, per step, in
blocks of 20 and in one long run (fill/drain amortised).  Diagnosis only."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench as B
from flame_ros_b200 import capi, workload as WL

capi.load_library()
torch.cuda.set_device(0)
ts = torch.cuda.Stream()
torch.cuda.set_stream(ts)
datas = [WL.StreamData("C2", seed=s) for s in range(8)]
run = B.GpuRun(capi, datas, 0, ts.cuda_stream, 0)
ctx = run.ctx
import ctypes as C
variants = [("resident", "resident", None), ("resident + result read-back", "resident", True), ("e2e without result read-back", "e2e_pipe", False), ("e2e_pipe", "e2e_pipe", True)]
for label, mode, xout in variants:
    run.step(0, mode)   # builds the descriptor table
    for (kk, mm), d in run._table.items():
        if mm != mode or xout is None:
            continue
        d.x_out = run.h_x.array[kk % B.RESULT_BUFFERS].ctypes.data_as(C.POINTER(C.c_float)) if xout else None
    for k in range(30):
        run.step(k, mode)
    ctx.sync(); torch.cuda.synchronize()
    for steps in (20, 400):
        enq, wait = [], []
        t0 = time.perf_counter()
        for i in range(steps):
            a = time.perf_counter()
            run.step(30 + i, mode)
            b = time.perf_counter()
            if (mode == "e2e_pipe" and xout is not False or xout) and i >= B.RESULT_LAG:
                ctx.results_wait(B.RESULT_LAG)
            c = time.perf_counter()
            enq.append(b - a); wait.append(c - b)
        if (mode == "e2e_pipe" and xout is not False) or xout:
            ctx.results_wait(0)
        ctx.pipeline_join()
        torch.cuda.synchronize()
        t = time.perf_counter() - t0
        print("%s, %d steps: %.1f us/step wall; enqueue median %.1f us (p90 %.1f), wait median %.1f us" %
              (label, steps, 1e6 * t / steps, 1e6 * np.median(enq), 1e6 * np.percentile(enq, 90), 1e6 * np.median(wait)), flush=True)
run.close()
