"""Top stalled SASS instructions of a kernel from an .ncu-rep (source page, CSV).  Usage:
    python tools/ncu_hot.py REP [N] [--ctx C]
Prints the N instructions with the most warp-stall samples plus C neighbouring instructions."""
import csv, subprocess, sys, io
rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 25
ctx = int(sys.argv[sys.argv.index("--ctx") + 1]) if "--ctx" in sys.argv else 0
extra = ["--print-source", "sass"]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = out.split('"Kernel Name",')
for blk in blocks[1:2]:
    lines = blk.split("\n")
    print("kernel:", lines[0][:120])
    rd = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
    hdr = rd[0]
    rows = [r for r in rd[1:] if len(r) == len(hdr)]
    iS = hdr.index("# Samples"); iSrc = hdr.index("Source")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[iS] or 0) for r in rows)
    print("total samples", tot, "instructions", len(rows))
    agg = {}
    for r in rows:
        for i in stall_cols:
            agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
    print("stall mix:", {k: round(100 * v / max(tot, 1), 1) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]})
    order = sorted(range(len(rows)), key=lambda i: -int(rows[i][iS] or 0))[:n]
    for i in sorted(order):
        for j in range(max(0, i - ctx), min(len(rows), i + ctx + 1)):
            r = rows[j]
            top = sorted(((int(r[c] or 0), hdr[c]) for c in stall_cols), reverse=True)[:2]
            print("%s%5d %5.1f%%  %-60s %s" % ("*" if j == i else " ", j, 100 * int(r[iS] or 0) / max(tot, 1), r[iSrc].strip()[:60],
                                             " ".join("%s=%d" % (h[6:], v) for v, h in top if v)))
        if ctx: print()
