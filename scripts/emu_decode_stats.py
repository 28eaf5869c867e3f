"""Step mix of the decode kernel on reference level-6 streams of the bench corpus (CPU emulator build with -DINF_STATS;
dev tooling).  python scripts/emu_decode_stats.py [chunks] [-DINF_...]"""
import ctypes, os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
defs = [a for a in sys.argv[1:] if a.startswith("-D")]
os.environ["LDB_EMU_DEFS"] = " ".join(["-DINF_STATS"] + defs)
import libdeflate_b200 as ldb
from libdeflate_b200 import build
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 and not sys.argv[1].startswith("-D") else 64
lib = ldb.load_library(build.build_emu(force=True))
ctx = ldb.Context(0, lib)
synth = bench.load_synth()
chunk = 65536
buf = (ctypes.c_uint8 * (n * chunk))()
synth.synth_fill(buf, chunk, 0, n, 0, 2)
raw = bytes(buf)
streams = []
for i in range(n):
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    streams.append(c.compress(raw[i * chunk:(i + 1) * chunk]) + c.flush())
st = (ctypes.c_ulonglong * 16)()
lib.ldb_inf_stats(st, 1)
outs = ctx.decompress_batch_host(streams, [chunk] * n, fmt=ldb.RAW) if hasattr(ctx, "decompress_batch_host") else None
lib.ldb_inf_stats(st, 0)
v = list(st)
names = ["litlen sub smem", "litlen sub global", "offset sub smem", "offset sub global", "steps in ST_LIT", "steps in ST_OFF", "", "",
         "literal first", "length+offset", "length only", "offset only", "end of block", "idle lane-steps", "follow-on literals", ""]
tot = sum(v[8:13])
for k, name in enumerate(names):
    if name:
        print("%-20s %12d  %6.2f %% of active lane-steps" % (name, v[k], 100.0 * v[k] / max(tot, 1)))
print("symbols per active lane-step: %.3f" % ((v[8] + v[14] + 2 * v[9] + v[10] + v[11] + v[12]) / max(tot, 1)))
os.environ.pop("LDB_EMU_DEFS")
