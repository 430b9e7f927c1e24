#!/usr/bin/env python3
"""Top stall locations of an ncu source-page CSV (SASS level), grouped by opcode and listed by address.
usage: ncu -i x.ncu-rep --page source --csv > x.csv ; tools/ncu_hot.py x.csv [N]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    try: ns = int(r[ix["# Samples"]])
    except ValueError: continue
    data.append((ns, r))
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
byop = collections.Counter(); bystall = collections.Counter(); execop = collections.Counter()
for ns, r in data:
    op = r[ix["Source"]].split()[0] if r[ix["Source"]].split() else "?"
    if op.startswith("@"): op = r[ix["Source"]].split()[1]
    byop[op.split(".")[0]] += ns
    execop[op.split(".")[0]] += int(r[ix["Instructions Executed"]] or 0)
    for c in stall_cols:
        bystall[c] += int(r[ix[c]] or 0)
print("-- samples by opcode (share of samples | executed warp-instr)")
for op, n in byop.most_common(25): print("  %-10s %8d %5.1f%%   %12d" % (op, n, 100.0 * n / tot, execop[op]))
print("-- samples by stall reason")
for c, n in bystall.most_common(12): print("  %-22s %8d %5.1f%%" % (c, n, 100.0 * n / tot))
print("-- top instructions")
for ns, r in sorted(data, key=lambda x: -x[0])[:N]:
    top = sorted(((int(r[ix[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    print("  %6d %5.2f%%  %-60s %s" % (ns, 100.0 * ns / tot, r[ix["Source"]].strip()[:60], " ".join("%s=%d" % (c[6:], n) for n, c in top)))
