"""world_size-2 gloo test of the stream sharding + reduction logic bench.py uses for N > 1 GPUs, with
the per-rank work done by the CPU oracle on a tiny workload (the N>1 path has no data collective)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from flame_ros_b200 import sharding
    from flame_ros_b200 import workload as WL
    from oracle import oracle as O
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    red = sharding.Reducer(dist)
    ids = sharding.stream_ids(rank, world, 2)
    checksum = 0.0
    for sid in ids:
        d = WL.StreamData("tiny", seed=sid)
        st = O.new_state(np.full(d.V, 0.5, np.float32), d.E)
        z = (0.5 + 0.001 * np.arange(d.V) + 0.01 * sid).astype(np.float32)
        O.nltgv2_solve(d.u_ref, d.edges, d.alpha, d.beta, z, np.ones(d.V, np.float32), st,
                       O.NLTGV2Params.default(), 5)
        checksum += float(st["x"].astype(np.float64).sum())
    red.barrier()
    fps = sharding.whole_job_throughput(red, frames_this_rank=len(ids) * 10, seconds_this_rank=1.0 + rank)
    total = red.sum(checksum)
    q.put((rank, ids, fps, checksum, total))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, ids0, fps0, c0, t0), (r1, ids1, fps1, c1, t1) = res
    assert ids0 == [0, 1] and ids1 == [2, 3]          # disjoint cover of the 4 streams
    assert fps0 == fps1 == pytest.approx(40 / 2.0)    # total frames / MAX time over ranks
    assert c0 != c1 and t0 == t1 == pytest.approx(c0 + c1)


def test_stream_ids_cover_without_overlap():
    from flame_ros_b200 import sharding
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            seen += sharding.stream_ids(r, world, 3)
        assert seen == list(range(3 * world))
        assert all(sharding.owner_of(s, 3) == s // 3 for s in seen)
    with pytest.raises(ValueError):
        sharding.stream_ids(2, 2, 1)
    # BASELINE configs[4] as written: 8 streams in total, stream s -> rank s mod world
    for world in (1, 2, 4, 8):
        parts = [sharding.strong_stream_ids(r, world, 8) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(8))
        assert all(len(p) == 8 // world for p in parts)
        assert all(s % world == r for r, p in enumerate(parts) for s in p)
