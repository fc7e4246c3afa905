"""Per-kernel histogram of the SASS opcodes that prove a Blackwell-native path (B200_PROFILING.md):
    python tools/sass_hist.py [lib.so] > profiles/<round>_sass_histogram.txt
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = tensor-map TMA, UBLKCP = 1-D bulk TMA,
SYNCS = mbarrier ops, HMMA = legacy mma.sync (none expected)."""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "desire_b200", "csrc", "libdesire_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "USETMAXREG", "LDGSTS"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
cur, hist = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m:
        op = m.group(1)
        hist[cur]["_total"] += 1
        for o in OPS:
            if op.startswith(o):
                hist[cur][o] += 1
print("# SASS opcode histogram of %s (cuobjdump -sass), kernels with tensor-core / TMA / mbarrier instructions" % os.path.basename(lib))
print("%-78s %6s " % ("kernel", "instr") + " ".join("%8s" % o for o in OPS))
tot = collections.Counter()
for k, h in hist.items():
    tot.update(h)
    if not any(h[o] for o in OPS if o not in ("LDGSTS",)):
        continue
    name = re.sub(r"\(anonymous namespace\)::", "", demangle(k))
    name = re.sub(r"\(.*", "", name)
    print("%-78s %6d " % (name[:78], h["_total"]) + " ".join("%8d" % h[o] for o in OPS))
print("%-78s %6d " % ("ALL %d kernels" % len(hist), tot["_total"]) + " ".join("%8d" % tot[o] for o in OPS))
