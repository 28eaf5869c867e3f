"""world_size-2 `gloo` test of the N>1 path on CPU: shard ownership, the reductions bench.py
uses, and that the union of the shards is exactly the single-process batch (checked through
the oracle's CRC-32 of every chunk of a synthetic batch)."""
import ctypes
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _worker(rank, world, port, n_per_rank, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from libdeflate_b200 import shard
    import bench
    from conftest import Oracle
    orc = Oracle()
    synth = bench.load_synth()
    lo, hi = shard.shard_range(rank, world, n_per_rank)
    chunk = 4096
    buf = (ctypes.c_uint8 * (n_per_rank * chunk))()
    synth.synth_fill(buf, chunk, lo, n_per_rank, 6, 2)
    raw = bytes(buf)
    crcs = torch.tensor([orc.crc32(raw[i * chunk:(i + 1) * chunk]) for i in range(n_per_rank)], dtype=torch.int64)
    gathered = [torch.zeros_like(crcs) for _ in range(world)]
    dist.all_gather(gathered, crcs)
    red = shard.Reducer(dist)
    value, ms = shard.whole_job_throughput(n_per_rank * chunk, 3, 10.0 * (rank + 1), red)
    red.barrier()
    if rank == 0:
        out.put((torch.cat(gathered).tolist(), value, ms, lo, hi))
    dist.destroy_process_group()


def test_two_rank_sharding_and_reductions():
    import bench
    from conftest import Oracle
    world, n_per_rank, chunk = 2, 24, 4096
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_per_rank, q)) for r in range(world)]
    for p in procs:
        p.start()
    crcs, value, ms, lo, hi = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference of the same global batch
    orc = Oracle()
    synth = bench.load_synth()
    n = world * n_per_rank
    buf = (ctypes.c_uint8 * (n * chunk))()
    synth.synth_fill(buf, chunk, 0, n, 6, 2)
    raw = bytes(buf)
    assert crcs == [orc.crc32(raw[i * chunk:(i + 1) * chunk]) for i in range(n)]
    # max over ranks of the step time (20 ms), sum over ranks of the bytes
    assert abs(ms - 20.0) < 1e-9
    assert abs(value - (world * n_per_rank * chunk * 3) / 0.020) < 1e-3
    assert (lo, hi) == (0, n_per_rank)


def test_shard_helpers():
    from libdeflate_b200 import shard
    assert shard.shard_range(3, 8, 65536) == (196608, 262144)
    covered = []
    for r in range(8):
        lo, hi = shard.split_range(r, 8, 524288 + 5)
        covered += [lo, hi]
    assert covered[0] == 0 and covered[-1] == 524288 + 5
    assert all(covered[2 * i + 1] == covered[2 * i + 2] for i in range(7))


def _origin_worker(rank, world, port, n_total, chunk, out):
    """Single-origin round trip (libdeflate_b200.shard.OriginRoundTrip) over gloo with host tensors: the
    kernels are the emulator build of the SAME sources (tests only), the data plane code is the product's."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import zlib
    import libdeflate_b200 as ldb
    from libdeflate_b200 import build, shard
    import bench
    lib = ldb.load_library(build.build_emu())
    ctx = ldb.Context(0, lib)
    rt = shard.OriginRoundTrip(ctx, dist, "cpu", n_total, chunk, ldb.GZIP, 6, stages=3)
    root_in = None
    if rank == 0:
        synth = bench.load_synth()
        buf = (ctypes.c_uint8 * (n_total * chunk))()
        synth.synth_fill(buf, chunk, 0, n_total, 6, 2)
        root_in = torch.frombuffer(bytearray(bytes(buf)), dtype=torch.uint8)
    info = rt.step(root_in)
    ok_local = bool((rt.res[:rt.n] == 0).all()) and bool((rt.aout[:rt.n] == chunk).all())
    flags = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(flags, torch.tensor([1 if ok_local else 0], dtype=torch.int64))
    if rank == 0:
        raw = bytes(root_in.numpy())
        same = bytes(rt.out_all.numpy()) == raw
        # every gathered compressed chunk is a gzip member of its input chunk, found through the offset table
        offs, sizes, comp = rt.comp_offsets.tolist(), rt.comp_sizes.tolist(), bytes(rt.comp_all.numpy())
        members_ok = all(zlib.decompress(comp[offs[i]:offs[i] + sizes[i]], 31) == raw[i * chunk:(i + 1) * chunk] for i in range(n_total))
        out.put((same, members_ok, [int(f.item()) for f in flags], info, offs[-1]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_single_origin_round_trip():
    world, n_total, chunk = 2, 37, 8192       # odd count: the shards differ in size
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_origin_worker, args=(r, world, port, n_total, chunk, q)) for r in range(world)]
    for p in procs:
        p.start()
    same, members_ok, flags, info, packed_total = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert same and members_ok and flags == [1, 1]
    assert info["compressed_bytes_total"] == packed_total
    # what crossed the "link": the other rank's input out, its compressed bytes + size table and its output in
    n1 = n_total - (n_total // 2)
    assert info["nvlink_bytes"]["root_out"] == n1 * chunk
    assert info["nvlink_bytes"]["root_in"] > n1 * chunk
