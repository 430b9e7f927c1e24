#!/usr/bin/env python3
"""Latency blind rotation with one PBS per CTA (br_wide.cu) against two PBS per CTA (br_wide2.cu): launch time at batch sizes
between one and two waves of SMs.  usage: pair_probe.py [out.json] [sizes...]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fhe_regex_b200 as fb
out_path = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].isdigit() else None
sizes = [int(a) for a in sys.argv[1:] if a.isdigit()] or [2, 252, 296]
ck = fb.ClientKey.load(os.path.join(ROOT, "tests", "golden", "client_key"))
ksk, bsk = fb.keygen_server_raw(ck, seed=0)
sk = fb.ServerKey(ksk, bsk)
lut = fb.make_lut(lambda x: (x + 1) % 16)
base = ck.encrypt_blocks(np.arange(64) % 16, seed=3)
sk.timing(True)
sk.set_latency_threshold(1 << 30)
rows = []
for B in sizes:
    cts = np.ascontiguousarray(np.tile(base, ((B + 63) // 64, 1))[:B])
    idx = np.zeros(B, dtype=np.uint32)
    row = {"batch": B}
    outs = {}
    for name, pair, npre, skew, off in (("one_per_cta", 0, 1, 200, 0), ("pair_npre1", 2, 1, 200, 0), ("pair_npre1_skew400", 2, 1, 400, 0),
                                        ("pair_off1500", 2, 1, 400, 1500), ("pair_off3000", 2, 1, 400, 3000), ("pair_off4500", 2, 1, 400, 4500),
                                        ("pair_off3000_skew0", 2, 1, 0, 3000)):
        sk.set_option("wide_pair", pair)
        sk.set_option("wide_pair_offset", off)
        sk.set_option("wide_pair_prefetch", npre)
        sk.set_option("wide_skew", skew)
        outs[name] = sk.pbs(cts, lut[None], idx)
        sk.kernel_stats(reset=True)
        for _ in range(5):
            sk.pbs(cts, lut[None], idx)
        st = sk.kernel_stats(reset=True)
        row[name + "_br_ms"] = st["br_ms"] / 5
    row["bit_identical"] = bool(all((outs[k] == outs["one_per_cta"]).all() for k in outs))
    row["decrypt_ok"] = bool(all(ck.decrypt_block(outs["pair_npre1"][i]) == (i % 64 % 16 + 1) % 16 for i in range(0, B, max(1, B // 16))))
    rows.append(row)
    print(json.dumps(row), flush=True)
sk.close()
if out_path:
    json.dump(rows, open(out_path, "w"), indent=1)
