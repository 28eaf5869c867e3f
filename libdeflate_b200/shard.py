"""Multi-GPU plumbing for the batch path.

Chunks are independent streams (ref: lib/decompress_template.h:550 -- no preset dictionary,
lib/deflate_compress.c:2616 -- match finder re-initialised per call), so a batch shards by
contiguous chunk index ranges and no kernel ever needs another rank's data.  Two ways to run it:

* pre-sharded (bench.py's main number): every rank already holds its chunks; torch.distributed is
  used only for the barrier around the timed region and the max / sum reductions of scalars.
* single origin (`OriginRoundTrip`, BASELINE configs[4]): ONE rank owns the whole batch.  The data
  plane is NCCL point-to-point over NVLink: grouped send/recv scatters
  sub-batches of input chunks root -> ranks while the previous sub-batch is being compressed;
  compressed chunks are packed on the device (libdeflate_b200_pack_batch), the per-rank byte
  totals are all-gathered (NCCL has no gatherv), then one send/recv per rank brings the packed
  bytes and the per-chunk size tables to the root; decompressed sub-batches travel back the same
  way while the next one is being decoded.  No reduction crosses chunks, so there is nothing to
  fuse into the kernels beyond writing results where NCCL sends them from.
  The kernels see plain device pointers (tensor.data_ptr()); torch tensors are only the memory
  NCCL can address.  With backend "gloo" and host tensors the same code runs against the emulated
  library (tests/test_multirank_gloo.py).
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


class OriginRoundTrip:
    """compress + decompress of a batch that lives on ONE rank (root 0), sharded over all ranks.

    n_total chunks of `chunk` bytes; rank r works on chunks split_range(r, world, n_total), cut into
    `stages` sub-batches so that transfers overlap kernels.  After step():
      root:  comp_all / comp_offsets / comp_sizes = the packed compressed batch (chunk i at
             comp_all[comp_offsets[i] : + comp_sizes[i]], rank shards back to back),
             out_all = the decompressed batch (must equal the input);
      every rank: results of its shard in `res` (libdeflate_result per chunk).
    """

    def __init__(self, ctx, dist, device, n_total, chunk, fmt, level, stages=4):
        import torch
        self.torch = torch
        self.ctx, self.l, self.dist, self.device = ctx, ctx.l, dist, device
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1
        self.n_total, self.chunk, self.fmt, self.level = n_total, chunk, fmt, level
        self.lo, self.hi = split_range(self.rank, self.world, n_total)
        self.n = self.hi - self.lo
        bound = getattr(self.l, "libdeflate_%s_compress_bound" % ("deflate", "zlib", "gzip")[fmt])(None, chunk)
        self.cstride = (bound + 15) & ~15
        self.stages = max(1, min(stages, self.n)) if self.n else 1
        u8 = torch.uint8
        n = max(self.n, 1)
        self.in_loc = torch.empty(n * chunk, dtype=u8, device=device)
        self.slots = torch.empty(n * self.cstride, dtype=u8, device=device)
        self.packed = torch.empty(n * self.cstride, dtype=u8, device=device)
        self.out_loc = torch.empty(n * chunk, dtype=u8, device=device)
        i64 = torch.int64
        idx = torch.arange(n, dtype=i64, device=device)
        self.p_in = self.in_loc.data_ptr() + idx * chunk
        self.p_slot = self.slots.data_ptr() + idx * self.cstride
        self.p_out = self.out_loc.data_ptr() + idx * chunk
        self.s_chunk = torch.full((n,), chunk, dtype=i64, device=device)
        self.s_slot = torch.full((n,), self.cstride, dtype=i64, device=device)
        self.csz = torch.zeros(n, dtype=i64, device=device)          # compressed sizes
        self.aout = torch.zeros(n, dtype=i64, device=device)
        self.res = torch.zeros(n, dtype=torch.int32, device=device)
        self.offs = torch.zeros(n + 1, dtype=i64, device=device)      # packed offsets of this rank's shard
        if self.rank == 0:
            self.comp_all = torch.empty(n_total * self.cstride, dtype=u8, device=device)
            self.comp_sizes = torch.zeros(n_total, dtype=i64, device=device)
            self.comp_offsets = torch.zeros(n_total + 1, dtype=i64, device=device)
            self.out_all = torch.empty(n_total * chunk, dtype=u8, device=device)
        # torch ops (NCCL included) are ordered against the library's kernels by making the library's
        # stream torch's current stream
        self.stream = None
        if str(device).startswith("cuda"):
            self.stream = torch.cuda.ExternalStream(ctx.stream, device=device)
        self.nvlink_bytes = {"root_out": 0, "root_in": 0}

    # ---- helpers -----------------------------------------------------------------
    def _stage(self, rank_n, j):
        """local chunk range of sub-batch j for a rank that owns rank_n chunks"""
        return (rank_n * j) // self.stages, (rank_n * (j + 1)) // self.stages

    def _p2p(self, ops):
        if not ops:
            return []
        return self.dist.batch_isend_irecv(ops)

    @staticmethod
    def _wait(reqs):
        for r in reqs:
            r.wait()

    def _ptr(self, t, first):
        return t.data_ptr() + first * t.element_size()

    # ---- the step -------------------------------------------------------------------
    def step(self, root_in=None):
        """root_in: uint8 tensor of n_total * chunk bytes on the root (ignored elsewhere)."""
        torch = self.torch
        if self.stream is not None:
            with torch.cuda.stream(self.stream):
                return self._step(root_in)
        return self._step(root_in)

    def _step(self, root_in):
        torch, dist, P2POp = self.torch, self.dist, None
        if dist is not None:
            P2POp = dist.P2POp
        chunk, S = self.chunk, self.stages
        ctx, l = self.ctx, self.l
        self.nvlink_bytes = {"root_out": 0, "root_in": 0}

        # ---- scatter (root -> ranks) pipelined with compression -------------------------------
        def issue_scatter(j):
            ops = []
            if self.rank == 0:
                for r in range(self.world):
                    rlo, rhi = split_range(r, self.world, self.n_total)
                    a, b = self._stage(rhi - rlo, j)
                    if b <= a:
                        continue
                    src = root_in[(rlo + a) * chunk:(rlo + b) * chunk]
                    if r == 0:
                        self.in_loc[a * chunk:b * chunk].copy_(src)
                    else:
                        ops.append(P2POp(dist.isend, src, r))
                        self.nvlink_bytes["root_out"] += (b - a) * chunk
            else:
                a, b = self._stage(self.n, j)
                if b > a:
                    ops.append(P2POp(dist.irecv, self.in_loc[a * chunk:b * chunk], 0))
            return self._p2p(ops)

        pending = issue_scatter(0)
        for j in range(S):
            nxt = issue_scatter(j + 1) if j + 1 < S else []
            self._wait(pending)
            a, b = self._stage(self.n, j)
            if b > a:
                ctx._check(l.libdeflate_b200_compress_batch(ctx.h, self.fmt, self.level, self._ptr(self.p_in, a), self._ptr(self.s_chunk, a),
                                                            self._ptr(self.p_slot, a), self._ptr(self.s_slot, a), self._ptr(self.csz, a), b - a), "compress_batch")
            pending = nxt

        # ---- pack, exchange byte totals, gather compressed chunks + size tables at the root ---------
        if self.n:
            ctx._check(l.libdeflate_b200_pack_batch(ctx.h, self.p_slot.data_ptr(), self.csz.data_ptr(), self.n, self.packed.data_ptr(),
                                                    self.packed.numel(), self.offs.data_ptr()), "pack_batch")
        ctx.sync()
        total = int(self.offs[self.n].item()) if self.n else 0
        if dist is not None and self.world > 1:
            t = torch.tensor([total], dtype=torch.int64, device=self.device)
            allt = [torch.zeros(1, dtype=torch.int64, device=self.device) for _ in range(self.world)]
            dist.all_gather(allt, t)
            totals = [int(x.item()) for x in allt]
        else:
            totals = [total]
        ops = []
        if self.rank == 0:
            base = 0
            self.comp_offsets[0] = 0
            for r in range(self.world):
                rlo, rhi = split_range(r, self.world, self.n_total)
                if r == 0:
                    self.comp_all[:total].copy_(self.packed[:total])
                    self.comp_sizes[rlo:rhi].copy_(self.csz[:self.n])
                    self.comp_offsets[rlo:rhi + 1].copy_(self.offs[:self.n + 1] + base)
                elif rhi > rlo:
                    ops.append(P2POp(dist.irecv, self.comp_all[base:base + totals[r]], r))
                    ops.append(P2POp(dist.irecv, self.comp_sizes[rlo:rhi], r))
                    self.nvlink_bytes["root_in"] += totals[r] + 8 * (rhi - rlo)
                base += totals[r]
            self._wait(self._p2p(ops))
            # offsets of the other ranks' chunks: their sizes rounded up to 16, as packed
            for r in range(1, self.world):
                rlo, rhi = split_range(r, self.world, self.n_total)
                if rhi > rlo:
                    sz = (self.comp_sizes[rlo:rhi] + 15) & ~15
                    start = int(self.comp_offsets[rlo].item())
                    self.comp_offsets[rlo + 1:rhi + 1] = start + torch.cumsum(sz, 0)
        elif self.n:
            ops.append(P2POp(dist.isend, self.packed[:total], 0))
            ops.append(P2POp(dist.isend, self.csz[:self.n], 0))
            self._wait(self._p2p(ops))

        # ---- decompress the local shard, sub-batch by sub-batch, results travel back as they finish ----
        def issue_gather(j):
            ops = []
            if self.rank == 0:
                for r in range(1, self.world):
                    rlo, rhi = split_range(r, self.world, self.n_total)
                    a, b = self._stage(rhi - rlo, j)
                    if b > a:
                        ops.append(P2POp(dist.irecv, self.out_all[(rlo + a) * chunk:(rlo + b) * chunk], r))
                        self.nvlink_bytes["root_in"] += (b - a) * chunk
            else:
                a, b = self._stage(self.n, j)
                if b > a:
                    ops.append(P2POp(dist.isend, self.out_loc[a * chunk:b * chunk], 0))
            return self._p2p(ops)

        reqs = []
        for j in range(S):
            a, b = self._stage(self.n, j)
            if b > a:
                ctx._check(l.libdeflate_b200_decompress_batch(ctx.h, self.fmt, 0, self._ptr(self.p_slot, a), self._ptr(self.csz, a),
                                                              self._ptr(self.p_out, a), self._ptr(self.s_chunk, a), None, self._ptr(self.aout, a),
                                                              self._ptr(self.res, a), b - a), "decompress_batch")
            if self.rank == 0 and b > a:
                self.out_all[(self.lo + a) * chunk:(self.lo + b) * chunk].copy_(self.out_loc[a * chunk:b * chunk])
            if dist is not None and self.world > 1:
                reqs += issue_gather(j)
        self._wait(reqs)
        ctx.sync()
        return {"compressed_bytes_total": sum(totals), "nvlink_bytes": dict(self.nvlink_bytes)}
