#!/usr/bin/env python3
"""Split an ncu source-page CSV (SASS order) into segments at BAR/SYNCS/LDTM markers and print per-segment
executed instructions, stall samples and the dominant stall reasons.  usage: ncu_phases.py x.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
segs = []; cur = {"name": "start", "n": 0, "exec": 0, "samp": 0, "st": collections.Counter(), "ops": collections.Counter()}
def flush(name):
    global cur
    segs.append(cur); cur = {"name": name, "n": 0, "exec": 0, "samp": 0, "st": collections.Counter(), "ops": collections.Counter()}
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip(); toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    if op.startswith("BAR") or op.startswith("SYNCS") or op.startswith("WARPSYNC") or op.startswith("ATOMS"):
        flush(op)
    try: ns = int(r[ix["# Samples"]]); ex = int(r[ix["Instructions Executed"]] or 0)
    except ValueError: continue
    cur["n"] += 1; cur["exec"] += ex; cur["samp"] += ns
    base = op.split(".")[0]
    cur["ops"][base if base in ("DADD", "DMUL", "DFMA") else ("LDS/STS" if base in ("LDS", "STS") else "other")] += ex
    for c in stall_cols: cur["st"][c[6:]] += int(r[ix[c]] or 0)
segs.append(cur)
tot = sum(s["samp"] for s in segs); tex = sum(s["exec"] for s in segs)
print("total samples %d  executed %d" % (tot, tex))
print("%4s %-22s %6s %12s %6s %7s %6s  %s" % ("#", "starts at", "instr", "executed", "ex%", "samp%", "fp64%", "top stalls"))
for i, s in enumerate(segs):
    if s["samp"] < tot * 0.004: continue
    fp = sum(s["ops"][k] for k in ("DADD", "DMUL", "DFMA"))
    top = " ".join("%s=%.0f%%" % (k, 100.0 * v / max(1, s["samp"])) for k, v in s["st"].most_common(4))
    print("%4d %-22s %6d %12d %5.1f%% %6.1f%% %5.0f%%  %s" % (i, s["name"][:22], s["n"], s["exec"], 100.0 * s["exec"] / tex, 100.0 * s["samp"] / tot, 100.0 * fp / max(1, s["exec"]), top))
