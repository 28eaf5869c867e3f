"""Host<->device copy bandwidth through the library's own pinned allocations (tuning aid for the e2e path)."""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import libdeflate_b200 as ldb  # noqa: E402

l = ldb.load_library()
ctx = ldb.Context(0, l)
n = 1 << 30
h = l.libdeflate_b200_pinned_malloc(n)
d = l.libdeflate_b200_device_malloc(ctx.h, n)
ctypes.memset(h, 1, n)
for name, fn, a, b in (("h2d", l.libdeflate_b200_memcpy_h2d, d, h), ("d2h", l.libdeflate_b200_memcpy_d2h, h, d)):
    for rep in range(3):
        l.libdeflate_b200_ctx_sync(ctx.h)
        t0 = time.perf_counter()
        fn(ctx.h, a, b, n)
        l.libdeflate_b200_ctx_sync(ctx.h)
        dt = time.perf_counter() - t0
        print("%s 1 GiB: %.1f GB/s" % (name, n / dt / 1e9))
