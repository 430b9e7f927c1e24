#!/usr/bin/env python3
"""Latency blind rotation (br_wide.cu): launch time over its two tuning options.  usage: wide_probe.py [batch]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fhe_regex_b200 as fb
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
ck = fb.ClientKey.load(os.path.join(ROOT, "tests", "golden", "client_key"))
ksk, bsk = fb.keygen_server_raw(ck, seed=0)
sk = fb.ServerKey(ksk, bsk)
lut = fb.make_lut(lambda x: (x + 1) % 16)
base = ck.encrypt_blocks(np.arange(64) % 16, seed=3)
sk.timing(True)
sk.set_latency_threshold(1 << 30)
sk.set_option("wide_pair", 0)
cts = np.ascontiguousarray(np.tile(base, ((B + 63) // 64, 1))[:B])
idx = np.zeros(B, dtype=np.uint32)
for npre in (0, 1, 2, 3, 4):
    for skew in (0, 100, 200, 300, 400):
        sk.set_option("wide_prefetch", npre)
        sk.set_option("wide_skew", skew)
        out = sk.pbs(cts, lut[None], idx)
        ok = all(ck.decrypt_block(out[i]) == (i % 64 % 16 + 1) % 16 for i in range(0, B, max(1, B // 16)))
        sk.kernel_stats(reset=True)
        for _ in range(5):
            sk.pbs(cts, lut[None], idx)
        st = sk.kernel_stats(reset=True)
        print(json.dumps({"batch": B, "wide_prefetch": npre, "wide_skew": skew, "br_ms": round(st["br_ms"] / 5, 4), "ok": ok}), flush=True)
sk.close()
