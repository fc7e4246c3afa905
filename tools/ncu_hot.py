"""Hot spots of one .ncu-rep (source page): python tools/ncu_hot.py file.ncu-rep [min_pct]"""
import csv, subprocess, sys, collections
f = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 1.5
out = subprocess.run(["ncu", "-i", f, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[1]
isrc, iex, ist = h.index('Source'), h.index('Instructions Executed'), h.index('Warp Stall Sampling (All Samples)')
iw, ie = h.index('L1 Wavefronts Shared'), h.index('L1 Wavefronts Shared Excessive')
data = []
for r in rows[2:]:
    try: data.append((r[isrc], int(r[iex] or 0), int(r[ist] or 0), int(r[iw] or 0), int(r[ie] or 0)))
    except Exception: pass
ts = sum(d[2] for d in data); te = sum(d[1] for d in data); tw = sum(d[3] for d in data)
print("instr %d  samples %d  smem wavefronts %d (excess %d)" % (te, ts, tw, sum(d[4] for d in data)))
ops = collections.Counter()
for s, e, st, w, x in data:
    t = s.split()
    if not t: continue
    op = t[1] if t[0].startswith('@') and len(t) > 1 else t[0]
    ops[op.split('.')[0]] += e
print("opcode mix:", ", ".join("%s %.1f%%" % (k, 100 * v / te) for k, v in ops.most_common(14)))
for i, d in enumerate(data):
    if d[2] > ts * thr / 100:
        print("%5d %6.2f%% st  exec %9d  wf %9d   %s" % (i, 100 * d[2] / ts, d[1], d[3], d[0][:100]))
