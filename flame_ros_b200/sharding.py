"""Multi-GPU sharding of the hot path: independent camera streams are the unit.

Frames of one stream are sequentially dependent and one graph (<= 3 MB) is far too small to cut
across GPUs, so the path shards ACROSS streams only: rank r owns streams [r*S, (r+1)*S); there is no
data-path collective.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used for the start
barrier and for reducing the per-rank timings: whole-job throughput = total frames / max-over-ranks
time.
"""


def stream_ids(rank, world, streams_per_rank):
    """Global ids (= texture / feature seeds) of the streams rank `rank` owns."""
    if not (0 <= rank < world) or streams_per_rank < 1:
        raise ValueError("bad rank/world/streams_per_rank")
    return list(range(rank * streams_per_rank, (rank + 1) * streams_per_rank))


def owner_of(stream_id, streams_per_rank):
    return stream_id // streams_per_rank


def strong_stream_ids(rank, world, total_streams):
    """BASELINE configs[4] as written: a FIXED batch of `total_streams`
    streams, stream s -> rank s mod world (8 / 4 / 2 / 1 streams per GPU at 1 / 2 / 4 / 8 GPUs)."""
    if not (0 <= rank < world) or total_streams < 1:
        raise ValueError("bad rank/world/total_streams")
    return [s for s in range(total_streams) if s % world == rank]


class Reducer:
    """max / sum over ranks of python floats; identity when not distributed."""

    def __init__(self, dist=None, device="cpu"):
        self.dist = dist if (dist is not None and dist.is_available() and dist.is_initialized()) else None
        self.device = device

    def _reduce(self, v, op_name):
        if self.dist is None:
            return float(v)
        import torch
        t = torch.tensor([float(v)], dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op_name))
        return float(t.item())

    def max(self, v):
        return self._reduce(v, "MAX")

    def sum(self, v):
        return self._reduce(v, "SUM")

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()


def whole_job_throughput(reducer, frames_this_rank, seconds_this_rank):
    """frames/s of the whole job: sum of frames over ranks / max of seconds over ranks."""
    return reducer.sum(frames_this_rank) / reducer.max(seconds_this_rank)
