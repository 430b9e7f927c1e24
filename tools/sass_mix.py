#!/usr/bin/env python3
"""Instruction mix of one kernel's SASS (cuobjdump -sass -fun <mangled> lib.so > x.sass), in windows along the
program order: shows how well FP64 work and shared-memory / integer work interleave.  usage: sass_mix.py x.sass [window]"""
import re, sys, collections
win = int(sys.argv[2]) if len(sys.argv) > 2 else 200
ins = []
for ln in open(sys.argv[1]):
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if not m:
        continue
    t = m.group(2).split()
    op = t[1] if t[0].startswith("@") else t[0]
    ins.append(op)
def cls(op):
    b = op.split(".")[0]
    if b in ("DFMA", "DADD", "DMUL"): return "fp64"
    if b in ("LDS", "STS", "SHFL", "LDTM", "STTM", "ATOMS", "LDG", "STG", "LDSM"): return "lsu"
    if b in ("BAR", "SYNCS", "WARPSYNC", "BRA", "BSSY", "BSYNC", "VOTE"): return "ctl"
    return "int"
tot = collections.Counter(cls(o) for o in ins)
print("instructions", len(ins), dict(tot))
ops = collections.Counter(o.split(".")[0] for o in ins)
print(" ".join("%s=%d" % kv for kv in ops.most_common(24)))
print("window  fp64  lsu  int  ctl")
for i in range(0, len(ins), win):
    c = collections.Counter(cls(o) for o in ins[i:i + win])
    marks = [o for o in ins[i:i + win] if o.startswith(("BAR", "SYNCS", "BRA"))]
    print("%5d  %4d %4d %4d %4d  %s" % (i, c["fp64"], c["lsu"], c["int"], c["ctl"], " ".join(m.split(".")[0] for m in marks[:8])))
