"""Compression ratio of a deflate-kernel variant on the bench corpus, computed on the CPU emulator build
(dev tooling: ratio questions do not need the GPU, only timing does).
    python scripts/emu_ratio.py NAME [-DX=..] ... [--chunks N] [--level L]"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import libdeflate_b200 as ldb  # noqa: E402
from libdeflate_b200 import build as b  # noqa: E402
import bench  # noqa: E402


def main():
    args = sys.argv[1:]
    name = args[0]
    defs = [x for x in args[1:] if x.startswith("-D")]
    n = int(args[args.index("--chunks") + 1]) if "--chunks" in args else 48
    level = int(args[args.index("--level") + 1]) if "--level" in args else 6
    chunk = 65536
    out = os.path.join(ROOT, "tests", "emu", "_build", "var")
    os.makedirs(out, exist_ok=True)
    b.build_emu()
    common = ["g++", "-O1", "-std=c++17", "-fPIC", "-DLDB_EMU", "-I", b.EMU_DIR, "-include", "cuda_emu.h", "-Wno-unused-function", "-fno-strict-aliasing"]
    objs = []
    for src in b.SOURCES:
        if src == "deflate_kernel.cu":
            obj = os.path.join(out, "%s_deflate.o" % name)
            subprocess.check_call(common + defs + ["-x", "c++", "-c", os.path.join(b.CSRC, src), "-o", obj])
        else:
            obj = os.path.join(os.path.dirname(b.EMU_LIB), src.replace(".cu", ".emu.o"))
        objs.append(obj)
    objs.append(os.path.join(os.path.dirname(b.EMU_LIB), "cuda_emu.o"))
    so = os.path.join(out, "lib_%s.so" % name)
    subprocess.check_call(["g++", "-shared", "-o", so] + objs + ["-lpthread"])
    lib = ldb.load_library(so)
    ctx = ldb.Context(0, lib)
    synth = bench.load_synth()
    buf = (ctypes.c_uint8 * (n * chunk))()
    synth.synth_fill(buf, chunk, 0, n, 0, 2)
    raw = bytes(buf)
    outs = ctx.compress_batch_host([raw[i * chunk:(i + 1) * chunk] for i in range(n)], level=level, fmt=ldb.GZIP)
    tot = sum(len(o) for o in outs)
    print("%s %s: %d chunks, ratio %.5f" % (name, " ".join(defs), n, tot / (n * chunk)))


main()
