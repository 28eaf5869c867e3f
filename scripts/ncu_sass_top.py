"""Top SASS instructions of a kernel by stall samples, with the dominant stall reasons.
    python scripts/ncu_sass_top.py rep.ncu-rep [top N]"""
import csv, io, subprocess, sys

def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    hdr, rows = None, []
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if "Source" in r and "# Samples" in r:
            hdr = r; continue
        if hdr is None or len(r) < len(hdr):
            continue
        rows.append(r)
    si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
    src = hdr.index("Source")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "not_issued" not in h]
    def num(x):
        try: return int(x)
        except ValueError: return 0
    tot = sum(num(r[si]) for r in rows) or 1
    toti = sum(num(r[ii]) for r in rows) or 1
    print("total samples %d, warp instructions %d" % (tot, toti))
    agg = {}
    for r in rows:
        for i in stall_cols:
            agg[hdr[i]] = agg.get(hdr[i], 0) + num(r[i])
    s = sum(agg.values()) or 1
    print("stall mix:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / s) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for idx, r in sorted(enumerate(rows), key=lambda t: -num(t[1][si]))[:top]:
        reasons = sorted(((num(r[i]), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
        print("%5.1f%%  #%-5d %-60s %s" % (100.0 * num(r[si]) / tot, idx, r[src].strip()[:60], " ".join("%s=%d" % (n, v) for v, n in reasons if v)))

main()
