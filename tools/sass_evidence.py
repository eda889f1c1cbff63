"""SASS evidence for profiles/: per kernel of libmiso_b200.so the mnemonics that prove what the
DESIGN says -- TMA bulk copy + mbarrier (UBLKCP, SYNCS.*), no tensor-core instructions
(UTCMMA / LDTM / HMMA: deliberately unused), the fp64 pipe, IMAD.WIDE (Philox), shuffles.
    python tools/sass_evidence.py > profiles/r2_sass_tma_excerpt.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "miso_b200", "libmiso_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ["UBLKCP", "SYNCS.ARRIVE", "SYNCS.PHASECHK", "UTMALDG", "UTCMMA", "LDTM", "HMMA", "IMAD.WIDE", "DFMA", "DMUL",
         "DADD", "MUFU", "SHFL", "LDS", "STS", "LDL", "STL", "REDUX", "ATOMG", "NANOSLEEP"]
arch, fn = None, None
counts = collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch = m.group(1)
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        fn = re.sub(r"\(misob200::ChainParams\)|misob200::|\(anonymous namespace\)::", "", fn)
        counts[fn] = collections.Counter(total=0, arch=arch)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and fn:
        op = m.group(1)
        counts[fn]["total"] += 1
        for w in WATCH:
            if op.startswith(w):
                counts[fn][w] += 1
print("cuobjdump -sass %s  (%s)" % (os.path.relpath(lib, ROOT), "; ".join(sorted({str(c["arch"]) for c in counts.values()}))))
print("static instruction counts per kernel; 0 everywhere for UTCMMA / LDTM / HMMA / UTMALDG: no tensor-core and no")
print("tensor-map instructions -- tiles travel by 1-D bulk copies (UBLKCP) completing on an mbarrier (SYNCS)\n")
print("%-58s %6s " % ("kernel", "instr") + " ".join("%7s" % w[:7] for w in WATCH))
for fn, c in counts.items():
    print("%-58s %6d " % (fn[:58], c["total"]) + " ".join("%7d" % c[w] for w in WATCH))
tot = collections.Counter()
for c in counts.values():
    for w in WATCH:
        tot[w] += c[w]
print("\nkernels: %d; with UBLKCP: %d; tensor-core instructions in the whole library: %d" % (
    len(counts), sum(1 for c in counts.values() if c["UBLKCP"]), tot["UTCMMA"] + tot["LDTM"] + tot["HMMA"]))
