#!/usr/bin/env python3
"""Per-kernel SASS opcode histogram of libfhe_b200.so (cuobjdump -sass): the mnemonics that show what each kernel is built
from -- UBLKCP (cp.async.bulk / TMA bulk copy), LDTM / STTM / UTCCP (tensor memory), IMMA / UTC*MMA (tensor cores),
DFMA / DADD / DMUL (FP64 pipe), LDS / STS / SHFL (shared-memory pipe), BAR / SYNCS (barriers, mbarriers).
usage: tools/sass_hist.py [lib.so] > profiles/rNN_sass_opcode_histogram.txt"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fhe_regex_b200", "libfhe_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for ln in sass.splitlines():
    m = re.match(r"\s+Function : (\S+)", ln)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
    if m and kern:
        t = m.group(1).split()
        op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
        hist[kern][op.split(".")[0] + ("." + ".".join(op.split(".")[1:3]) if op.startswith(("IMMA", "UTC", "LDTM", "STTM", "UBLKCP", "LDSM", "LDGSTS")) else "")] += 1
KEYS = ["UBLKCP", "UTMALDG", "LDTM", "STTM", "UTCCP", "UTCBAR", "IMMA", "UTCIMMA", "UTCHMMA", "HMMA", "LDSM", "LDGSTS", "DFMA", "DADD", "DMUL", "LDS", "STS", "SHFL",
        "LDG", "STG", "ATOMS", "BAR", "SYNCS", "I2F", "F2I"]
for k, c in hist.items():
    tot = sum(c.values())
    print("%s: %d instructions" % (k, tot))
    line = []
    for key in KEYS:
        n = sum(v for op, v in c.items() if op.split(".")[0] == key)
        if n:
            variants = sorted({op for op in c if op.split(".")[0] == key and "." in op})
            line.append("%s=%d%s" % (key, n, (" (" + ", ".join(variants) + ")") if variants else ""))
    print("   " + "  ".join(line))
