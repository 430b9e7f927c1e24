#!/usr/bin/env python3
"""Blind-rotation launch time versus batch size (one wave at S = 1..4 samples per SM, and wave boundaries)."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fhe_regex_b200 as fb
ck = fb.ClientKey.load(os.path.join(ROOT, "tests", "golden", "client_key"))
ksk, bsk = fb.keygen_server_raw(ck, seed=0)
sk = fb.ServerKey(ksk, bsk)
lut = fb.make_lut(lambda x: x)
base = ck.encrypt_blocks(np.arange(64) % 16, seed=3)
sk.timing(True)
out = {}
# both blind rotations: "throughput" = up to 4 PBS per CTA (kernels.cu), "latency" = one PBS per CTA (br_wide.cu)
for variant, thr, sizes in (("throughput", 0, [1, 37, 74, 148, 222, 296, 370, 444, 518, 592, 620, 740, 888, 1184]),
                            ("latency", 1 << 30, [1, 37, 74, 148, 149, 222, 296, 297, 444, 592])):
    sk.set_latency_threshold(thr)
    out[variant] = {}
    for B in sizes:
        cts = np.ascontiguousarray(np.tile(base, ((B + 63) // 64, 1))[:B])
        idx = np.zeros(B, dtype=np.uint32)
        sk.pbs(cts, lut[None], idx)
        sk.kernel_stats(reset=True)
        for _ in range(3):
            sk.pbs(cts, lut[None], idx)
        st = sk.kernel_stats(reset=True)
        out[variant][B] = {"br_ms": st["br_ms"] / 3, "ks_ms": st["ks_ms"] / 3}
        print(variant, B, out[variant][B], flush=True)
sk.close()
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
