"""Tuning sweeps: builds variant libraries (different -D geometry) under build/variants/.
Dev tooling only; the product library is always libdeflate_b200/libdeflate_b200.so."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libdeflate_b200 import build as b  # noqa: E402

GEO_A = ["-DINF_LB=8", "-DINF_LSUB_SM=64", "-DINF_OB=6", "-DINF_OSUB_SM=64"]      # 896 B/lane, 7 warps/SM (default)
GEO_G = ["-DINF_LB=7", "-DINF_LSUB_SM=96", "-DINF_OB=6", "-DINF_OSUB_SM=64"]      # 704 B, 9 warps
GEO_C = ["-DINF_LB=7", "-DINF_LSUB_SM=64", "-DINF_OB=5", "-DINF_OSUB_SM=32"]      # 512 B, 13 warps
GEO_D = ["-DINF_LB=7", "-DINF_LSUB_SM=32", "-DINF_OB=5", "-DINF_OSUB_SM=32"]      # 448 B, 15 warps
VARIANTS = {
    # resolve kernel (32 single-warp CTAs per SM is the hardware's CTA limit)
    # deflate kernel: half the threads per CTA (2 search runs per thread, 128 registers each)
    "t512": ["-DLZ_THREADS=512"],
    "run8": ["-DLZ_RUN_SHORT=8"],
    "run8d32": ["-DLZ_RUN_SHORT=8", "-DLZ_L6_DEPTH=32"],
    "d16": ["-DLZ_L6_DEPTH=16"],
    # decode table geometry at the same 448 B per lane (15 warps per SM): 6-bit offset main table, 16 + 16 shared subtable entries
    "g6": ["-DINF_LSUB_SM=16", "-DINF_OB=6", "-DINF_OSUB_SM=16"],
    "g6b": ["-DINF_LSUB_SM=0", "-DINF_OB=6", "-DINF_OSUB_SM=32"],
}
# decode kernel, offset table geometry: 6-bit main + 16 shared subtable entries at 14 warps per SM 17.5 ms (= default);
# 13-warp geometries 27-28 ms (65536 chunks no longer fit one wave of lanes)
# decode steps between service phases, -DINF_QUANTUM=128/384/1024: 17.4 / 17.5 / 17.6 ms
# resolve kernel, warps (= chunks) per SM, -DRES_PER_SM=8/16/24/32: 25.9 / 14.2 / 10.6 / 9.0 ms per 65536 x 64 KiB
# measured and dropped: resumable chain walk of the deflate search (-DLZ_QUANTUM=6/8/12): 86.9 / 86.4 / 84.3 ms vs 79.2 ms
# per 16384 chunks (profiles/r02_deflate_quantum_ab.md)


def main():
    outdir = os.path.join(ROOT, "build", "variants")
    os.makedirs(outdir, exist_ok=True)
    nvcc = "/usr/local/cuda/bin/nvcc"
    for name, defs in VARIANTS.items():
        objs = []
        for src in b.SOURCES:
            obj = os.path.join(outdir, "%s_%s.o" % (name, src.replace(".cu", "")))
            subprocess.check_call([nvcc] + [f for f in b.NVCC_FLAGS if f not in ("-Xptxas", "-v")] + defs + ["-c", os.path.join(b.CSRC, src), "-o", obj])
            objs.append(obj)
        so = os.path.join(outdir, "libdeflate_b200_%s.so" % name)
        subprocess.check_call([nvcc, "-shared", "-o", so] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static"])
        print(so)


if __name__ == "__main__":
    main()
