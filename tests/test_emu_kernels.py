"""Kernel LOGIC on the CPU: the .cu sources compiled against tests/emu (SIMT emulator).

This is test infrastructure (the dev container has no GPU); the product path is only
ever the nvcc build, exercised by test_gpu_parity.py under `-m gpu`.
"""
import parity_checks as pc


def test_checksum_kernels_emulated(emu_api, oracle):
    pc.check_checksums(emu_api, oracle, sizes=[0, 1, 15, 16, 17, 33, 511, 512, 513, 5553, 65536, 262144 + 17])


def test_checksum_batch_emulated(emu_ctx, oracle):
    pc.check_checksum_batch(emu_ctx, oracle, n_chunks=40, max_len=20000)


def test_inflate_valid_streams_emulated(emu_ctx, oracle, reflib):
    streams = pc.make_valid_streams(sizes=(0, 1, 100, 5000, 40000), levels=(1, 6), ref=reflib)
    pc.check_decompress_valid(emu_ctx, oracle, streams)


def test_inflate_fuzz_emulated(emu_ctx, oracle):
    v = pc.check_decompress_fuzz(emu_ctx, oracle, pc.fuzz_cases(1500, seed=11))
    assert set(v) >= {0, 1, 3}, v
