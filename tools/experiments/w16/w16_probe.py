#!/usr/bin/env python3
"""Latency blind rotations: 8 points per thread (br_wide.cu / br_wide2.cu) against 16 points per thread with 1..3 samples per CTA
(br_w16.cu): launch time and decryptions at a few batch sizes.  usage: w16_probe.py [out.json] [sizes...]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fhe_regex_b200 as fb
out_path = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].isdigit() else None
sizes = [int(a) for a in sys.argv[1:] if a.isdigit()] or [1, 3, 148, 252, 296, 297, 444]
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
    for name, kern, S, skew in (("wide8", 0, 1, 200), ("w16_s1", 1, 1, 200), ("w16_s2", 1, 2, 200), ("w16_s2_skew0", 1, 2, 0), ("w16_s3", 1, 3, 200), ("w16_s3_skew0", 1, 3, 0)):
        sk.set_option("latency_kernel", kern)
        sk.set_option("w16_samples", S)
        sk.set_option("wide_skew", skew)
        out = sk.pbs(cts, lut[None], idx)
        sk.kernel_stats(reset=True)
        for _ in range(5):
            sk.pbs(cts, lut[None], idx)
        st = sk.kernel_stats(reset=True)
        row[name + "_br_ms"] = round(st["br_ms"] / 5, 4)
        row[name + "_ok"] = bool(all(ck.decrypt_block(out[i]) == (i % 64 % 16 + 1) % 16 for i in range(0, B, max(1, B // 24))))
    rows.append(row)
    print(json.dumps(row), flush=True)
sk.close()
if out_path:
    json.dump(rows, open(out_path, "w"), indent=1)
