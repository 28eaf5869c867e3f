"""python -m libdeflate_b200.gz [-d] [-1..-12] [-c] [-k] FILE...

A minimal gzip-style front end over the blocked-gzip (BGZF) calls, standing in for the part of the
reference's programs/gzip.c that drives the library (compress: programs/gzip.c:170-174, decompress loop:
:249-273).  Compressed files are ordinary multi-member .gz files (readable by any gunzip); decompression
accepts blocked gzip files (ours, bgzip's).  There is no CPU fallback: without a CUDA device this fails.
"""
import argparse
import os
import struct
import sys

import libdeflate_b200 as ldb


def uncompressed_size(data):
    """Sum of the members' ISIZE fields, found through the BC subfields (no decoding)."""
    pos, total = 0, 0
    while pos + 26 <= len(data):
        if data[pos:pos + 3] != b"\x1f\x8b\x08" or not data[pos + 3] & 4:
            raise ValueError("not a blocked gzip (BGZF) file")
        xlen = struct.unpack_from("<H", data, pos + 10)[0]
        if pos + 12 + xlen + 8 > len(data):      # the extra field and a trailer must fit (as the C walker checks)
            raise ValueError("corrupt BGZF member at byte %d" % pos)
        x, bsize = 0, 0
        while x + 4 <= xlen:
            si1, si2, slen = struct.unpack_from("<BBH", data, pos + 12 + x)
            if (si1, si2, slen) == (66, 67, 2):
                bsize = struct.unpack_from("<H", data, pos + 12 + x + 4)[0] + 1
            x += 4 + slen
        if bsize < 12 + xlen + 8 or pos + bsize > len(data):
            raise ValueError("corrupt BGZF member at byte %d" % pos)
        total += struct.unpack_from("<I", data, pos + bsize - 4)[0]
        pos += bsize
    if pos != len(data):
        raise ValueError("trailing bytes after the last BGZF member")
    if total > 1032 * len(data) + 64:            # DEFLATE cannot expand more than ~1032:1: a lying ISIZE
        raise ValueError("BGZF size fields exceed what the file can decode to")
    return total


def compress_bytes(ctx, data, level=6):
    out = ctx.bgzf_compress(data, level)
    assert out is not None
    return out


def decompress_bytes(ctx, data):
    n = uncompressed_size(data)
    res, out = ctx.bgzf_decompress(data, n)
    if res != 0:
        raise ValueError("decompression failed: libdeflate_result %d" % res)
    return out


def decompress_members(api, data):
    """Any multi-member gzip file, one member at a time through libdeflate_gzip_decompress_ex -- the loop of
    programs/gzip.c:249-273 (output buffer doubled on INSUFFICIENT_SPACE, next member at actual_in).  Every
    member is a single stream, i.e. one lane of the GPU: correct, not fast; blocked files go through
    decompress_bytes()."""
    out, pos = [], 0
    while pos < len(data):
        avail = max(4 * (len(data) - pos), 1 << 16)
        while True:
            res, piece, ain, _aout = api.decompress(data[pos:], avail, ldb.GZIP)
            if res != 3:        # LIBDEFLATE_INSUFFICIENT_SPACE
                break
            if avail > 1032 * (len(data) - pos) + (1 << 16):
                break           # more room cannot help: DEFLATE expands at most ~1032:1 (and streams >= 4 GiB are unsupported)
            avail *= 2
        if res != 0:
            raise ValueError("decompression failed: libdeflate_result %d at byte %d" % (res, pos))
        out.append(piece)
        pos += ain
    return b"".join(out)


def main(argv=None, ctx=None, api=None):
    ap = argparse.ArgumentParser(prog="python -m libdeflate_b200.gz", description=__doc__.split("\n\n")[1])
    ap.add_argument("-d", "--decompress", action="store_true")
    ap.add_argument("-c", "--stdout", action="store_true")
    ap.add_argument("-k", "--keep", action="store_true")
    for lvl in range(0, 13):
        ap.add_argument("-%d" % lvl, dest="level", action="store_const", const=lvl)
    ap.add_argument("files", nargs="+")
    args = ap.parse_args(argv)
    level = 6 if args.level is None else args.level
    ctx = ctx or ldb.Context(0)
    for path in args.files:
        data = open(path, "rb").read()
        if args.decompress:
            try:
                out = decompress_bytes(ctx, data)
            except ValueError:
                out = decompress_members(api or ldb.Api(), data)     # not blocked: member by member
            dst = path[:-3] if path.endswith(".gz") else path + ".out"
        else:
            out = compress_bytes(ctx, data, level)
            dst = path + ".gz"
        if args.stdout:
            sys.stdout.buffer.write(out)
        else:
            with open(dst, "wb") as f:
                f.write(out)
            if not args.keep:
                os.remove(path)
    return 0


if __name__ == "__main__":
    sys.exit(main())
