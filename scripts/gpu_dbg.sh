cat > /tmp/pipe_dbg.py <<'PY'
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from flame_ros_b200 import capi
from flame_ros_b200 import workload as WL
import test_gpu_hotpath_step as T
S = int(sys.argv[1]); cfg = sys.argv[2]
datas = [WL.StreamData(cfg, seed=s) for s in range(S)]
outs, feats = T._run(capi, datas, "pipe", 14)
print("pipe ok", float(np.nansum(outs[-1])))
PY
timeout 300 python /tmp/pipe_dbg.py 8 C2 2>&1 | tail -3
timeout 600 compute-sanitizer --tool memcheck python /tmp/pipe_dbg.py 4 C2 2>&1 | grep -vE "^$" | head -40
