"""`-m "not gpu"`: the C-ABI library loads and exports every symbol include/*.h declares
(no compute calls: there is no GPU here and no CPU fallback to call into)."""
import ctypes
import os
import re
import subprocess

import libdeflate_b200 as ldb
from libdeflate_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in ("libdeflate.h", "libdeflate_b200.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(libdeflate_[a-z0-9_]+)\s*\(", src))
    return names


def test_library_builds_and_exports_every_declared_symbol():
    so = build.build()
    out = subprocess.check_output(["nm", "-D", "--defined-only", so], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    decl = declared_symbols()
    assert len([d for d in decl if not d.startswith("libdeflate_b200_")]) == 21
    missing = decl - exported
    assert not missing, missing
    # nothing but the declared API leaks out of the library
    assert {e for e in exported if not e.startswith("_")} <= decl, exported - decl
    assert set(ldb.CLASSIC_SYMBOLS) | set(ldb.BATCH_SYMBOLS) == decl


def test_host_only_entry_points_without_gpu():
    l = ldb.load_library(build.build())
    # *_bound is host arithmetic and accepts a NULL compressor (ref: libdeflate.h:99-101)
    for n, want in ((0, 5), (1, 6), (4999, 5004), (5000, 5005), (5001, 5011), (65536, 65606), (1 << 20, 1049626)):
        assert l.libdeflate_deflate_compress_bound(None, n) == want
        assert l.libdeflate_zlib_compress_bound(None, n) == want + 6
        assert l.libdeflate_gzip_compress_bound(None, n) == want + 18
    # alloc/free contract (ref: lib/deflate_compress.c:3885-3896)
    for lvl in range(-1, 13):
        c = l.libdeflate_alloc_compressor(lvl)
        assert c
        l.libdeflate_free_compressor(c)
    assert not l.libdeflate_alloc_compressor(13)
    assert not l.libdeflate_alloc_compressor(-2)
    bad = ldb.Options(sizeof_options=8)
    assert not l.libdeflate_alloc_compressor_ex(6, ctypes.byref(bad))
    assert not l.libdeflate_alloc_decompressor_ex(ctypes.byref(bad))
    l.libdeflate_free_compressor(None)
    l.libdeflate_free_decompressor(None)
    # NULL-buffer checksums need no device (ref: lib/crc32.c:259, lib/adler32.c:159)
    assert l.libdeflate_crc32(0, None, 0) == 0
    assert l.libdeflate_adler32(1, None, 0) == 1


def test_custom_allocator_one_malloc_one_free():
    """ref: programs/test_custom_malloc.c:39-155 -- exactly one malloc and one free per object."""
    l = ldb.load_library(build.build())
    libc = ctypes.CDLL(None)
    libc.malloc.restype = ctypes.c_void_p
    libc.malloc.argtypes = [ctypes.c_size_t]
    libc.free.argtypes = [ctypes.c_void_p]
    calls = {"m": 0, "f": 0}
    MALLOC = ctypes.CFUNCTYPE(ctypes.c_void_p, ctypes.c_size_t)
    FREE = ctypes.CFUNCTYPE(None, ctypes.c_void_p)

    def m(n):
        calls["m"] += 1
        return libc.malloc(n)

    def f(p):
        calls["f"] += 1
        libc.free(p)

    mcb, fcb = MALLOC(m), FREE(f)
    o = ldb.Options(sizeof_options=ctypes.sizeof(ldb.Options), malloc_func=ctypes.cast(mcb, ctypes.c_void_p),
                    free_func=ctypes.cast(fcb, ctypes.c_void_p))
    for lvl in range(13):
        c = l.libdeflate_alloc_compressor_ex(lvl, ctypes.byref(o))
        assert c and calls["m"] == lvl + 1
        l.libdeflate_free_compressor(c)
        assert calls["f"] == lvl + 1
    d = l.libdeflate_alloc_decompressor_ex(ctypes.byref(o))
    assert d and calls["m"] == 14
    l.libdeflate_free_decompressor(d)
    assert calls["f"] == 14
    # failing allocator => NULL (fault injection, test_custom_malloc.c:120-143)
    fail = MALLOC(lambda n: None)
    o2 = ldb.Options(sizeof_options=ctypes.sizeof(ldb.Options), malloc_func=ctypes.cast(fail, ctypes.c_void_p),
                     free_func=ctypes.cast(fcb, ctypes.c_void_p))
    assert not l.libdeflate_alloc_compressor_ex(6, ctypes.byref(o2))
    assert not l.libdeflate_alloc_decompressor_ex(ctypes.byref(o2))
