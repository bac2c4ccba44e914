"""flame_ros_b200 -- B200-native FLaME hot path (NLTGV2-L1 primal-dual solver on the Delaunay graph +
per-feature epipolar inverse-depth update) behind the reference's `flame::Flame` boundary.

csrc/      hand-written sm_100a CUDA kernels + the C-ABI (include/flame_b200.h)
capi.py    ctypes mirror of the C-ABI used by tests and bench.py
synth.py   seeded synthetic graphs / image+pose streams (no dataset is available offline)
"""
from . import build  # noqa: F401

__all__ = ["build", "capi", "synth"]
