"""Multi-GPU plumbing for the batch path.

Chunks are independent streams (ref: lib/decompress_template.h:550 -- no preset dictionary,
lib/deflate_compress.c:2616 -- match finder re-initialised per call), so the batch shards by
contiguous chunk index ranges with NO data-path collective: rank r of G owns chunks
[r*n_per_rank, (r+1)*n_per_rank).  torch.distributed is used only for the barrier around the
timed region and for the max / sum reductions of scalars (device time, byte counts).
"""
import os


def dist_env():
    """(rank, world_size, local_rank) from the torchrun environment."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_range(rank, world, n_per_rank):
    """Weak-scaling shard: global chunk indices owned by `rank`."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    return rank * n_per_rank, (rank + 1) * n_per_rank


def split_range(rank, world, n_total):
    """Strong-scaling shard of a fixed batch: contiguous, sizes differ by at most one."""
    lo = (n_total * rank) // world
    hi = (n_total * (rank + 1)) // world
    return lo, hi


class Reducer:
    """Scalar all-reduce helpers on top of an initialised torch.distributed group (or a
    no-op for a single process)."""

    def __init__(self, dist=None, device="cpu"):
        self.dist = dist
        self.device = device

    def _reduce(self, x, op_name):
        if self.dist is None:
            return float(x)
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op_name))
        return float(t.item())

    def max(self, x):
        return self._reduce(x, "MAX")

    def sum(self, x):
        return self._reduce(x, "SUM")

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()


def whole_job_throughput(units_per_rank, steps, ms_per_rank, reducer):
    """value = units all ranks processed / max-over-ranks time (the bench contract)."""
    total = reducer.sum(units_per_rank) * steps
    ms = max(reducer.max(ms_per_rank), 1e-9)
    return total / (ms / 1e3), ms
