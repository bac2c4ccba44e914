"""Host-side logic of bench.py that needs no GPU: the per-rank core pinning of multi-rank runs."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _fake_nvml(masks):
    m = types.ModuleType("pynvml")
    m.nvmlInit = lambda: None
    m.nvmlDeviceGetHandleByIndex = lambda i: i
    m.nvmlDeviceGetCpuAffinity = lambda h, n: [(masks[h] >> (64 * k)) & ((1 << 64) - 1) for k in range(n)]
    return m


def test_ranks_split_the_cores_local_to_their_gpu(monkeypatch):
    avail = set(range(32))
    # GPUs 0-3 sit on cores 0-15, GPUs 4-7 on cores 16-31
    masks = [0xFFFF] * 4 + [0xFFFF0000] * 4
    got = {}
    monkeypatch.setitem(sys.modules, "pynvml", _fake_nvml(masks))
    monkeypatch.setattr(os, "sched_getaffinity", lambda pid: set(avail))
    monkeypatch.setattr(os, "sched_setaffinity", lambda pid, cpus: got.__setitem__("cpus", tuple(cpus)))
    monkeypatch.delenv("FB_BENCH_NO_PIN", raising=False)
    seen = []
    for r in range(8):
        bench.PINNING.clear()
        bench.PINNING["mode"] = "none"
        bench.pin_rank_to_its_cores(r, 8)
        assert bench.PINNING["mode"] == "nvml" and bench.PINNING["cores_per_rank"] == 4
        seen.append(got["cpus"])
    assert sorted(c for cpus in seen for c in cpus) == list(range(32))      # disjoint, complete
    assert all(max(seen[r]) < 16 for r in range(4)) and all(min(seen[r]) >= 16 for r in range(4, 8))
    assert bench.HOST_CORES_TOTAL[0] == 32


def test_pinning_is_skipped_quietly(monkeypatch):
    called = []
    monkeypatch.setattr(os, "sched_setaffinity", lambda pid, cpus: called.append(cpus))
    bench.PINNING.clear()
    bench.PINNING["mode"] = "none"
    bench.pin_rank_to_its_cores(0, 1)                      # a single rank is never pinned
    assert not called and bench.PINNING["mode"] == "none"
    bad = types.ModuleType("pynvml")
    bad.nvmlInit = lambda: (_ for _ in ()).throw(RuntimeError("no NVML here"))
    monkeypatch.setitem(sys.modules, "pynvml", bad)
    bench.pin_rank_to_its_cores(1, 2)                      # NVML missing: run unpinned, say why
    assert not called and bench.PINNING["mode"] == "none" and "no NVML" in bench.PINNING.get("reason", "")
