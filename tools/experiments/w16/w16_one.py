#!/usr/bin/env python3
"""One configuration of the radix-16 latency kernel, for profiling: w16_one.py <samples_per_cta> <batch>"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fhe_regex_b200 as fb
S, B = int(sys.argv[1]), int(sys.argv[2])
ck = fb.ClientKey.load(os.path.join(ROOT, "tests", "golden", "client_key"))
ksk, bsk = fb.keygen_server_raw(ck, seed=0)
sk = fb.ServerKey(ksk, bsk)
lut = fb.make_lut(lambda x: (x + 1) % 16)
base = ck.encrypt_blocks(np.arange(64) % 16, seed=3)
sk.set_latency_threshold(1 << 30)
sk.set_option("wide_pair", 0)
sk.set_option("latency_kernel", 1)
sk.set_option("w16_samples", S)
cts = np.ascontiguousarray(np.tile(base, ((B + 63) // 64, 1))[:B])
idx = np.zeros(B, dtype=np.uint32)
for _ in range(3):
    out = sk.pbs(cts, lut[None], idx)
print("ok", all(ck.decrypt_block(out[i]) == (i % 64 % 16 + 1) % 16 for i in range(0, B, max(1, B // 16))))
sk.close()
