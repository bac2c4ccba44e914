"""FB_PIPE_TRACE timeline of the e2e leg (see flame_b200.cu: PipeTrace).  Diagnosis only."""
import os
import sys

os.environ["FB_PIPE_TRACE"] = "40"
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
mode = sys.argv[1] if len(sys.argv) > 1 else "e2e_pipe"
for i in range(60):
    run.step(i, mode)
    if mode == "e2e_pipe" and i >= B.RESULT_LAG:
        run.ctx.results_wait(B.RESULT_LAG)
run.ctx.sync()
run.close()
