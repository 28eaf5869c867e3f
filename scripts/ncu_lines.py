"""Per-kernel source-line table of an .ncu-rep: warp instructions, thread instructions, stall samples per
source line, plus sums over line ranges given on the command line (name:lo-hi ...).
    python scripts/ncu_lines.py rep.ncu-rep <function substring> [top N] [name:lo-hi ...]"""
import csv, io, subprocess, sys

def main():
    rep, want = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    groups = []
    for g in sys.argv[4:]:
        name, r = g.split(":")
        lo, hi = r.split("-")
        groups.append((name, int(lo), int(hi)))
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    fn, hdr, per = None, None, {}
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "Function Name":
            fn = r[1]; continue
        if r[0] == "Line No":
            hdr = r; continue
        if hdr is None or fn is None or want not in fn or len(r) < len(hdr):
            continue
        try:
            ln = int(r[0])
            n = int(r[hdr.index("Instructions Executed")]); s = int(r[hdr.index("# Samples")]); t = int(r[hdr.index("Thread Instructions Executed")])
        except ValueError:
            continue
        d = per.setdefault(ln, [0, 0, 0, r[1].strip()[:110]])
        d[0] += n; d[1] += s; d[2] += t
    data = [(ln, v[0], v[1], v[2], v[3]) for ln, v in per.items()]
    tot = sum(d[1] for d in data) or 1
    tots = sum(d[2] for d in data) or 1
    print("function ~ %s: total warp instructions %d, samples %d" % (want, tot, tots))
    for name, lo, hi in groups:
        sel = [d for d in data if lo <= d[0] <= hi]
        n = sum(d[1] for d in sel); s = sum(d[2] for d in sel); t = sum(d[3] for d in sel)
        print("%-28s %5.1f%% inst %5.1f%% samples  %4.1f thr/inst" % (name, 100.0 * n / tot, 100.0 * s / tots, t / max(1, n)))
    print("| % inst | % samples | thr/inst | line | source |")
    for ln, n, s, t, code in sorted(data, key=lambda d: d[1], reverse=True)[:top]:
        print("| %.1f | %.1f | %.1f | %d | `%s` |" % (100.0 * n / tot, 100.0 * s / tots, t / max(1, n), ln, code))

main()
