#!/usr/bin/env python3
"""Latency blind rotation: launch time at a few batch sizes (env knobs are read at first launch).  usage: wide_probe.py [sizes...]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fhe_regex_b200 as fb
ck = fb.ClientKey.load(os.path.join(ROOT, "tests", "golden", "client_key"))
ksk, bsk = fb.keygen_server_raw(ck, seed=0)
sk = fb.ServerKey(ksk, bsk)
lut = fb.make_lut(lambda x: (x + 1) % 16)
base = ck.encrypt_blocks(np.arange(64) % 16, seed=3)
sk.timing(True)
variant = os.environ.get("PROBE_VARIANT", "latency")
sk.set_cluster_threshold((1 << 30) if variant == "cluster" else 0)
sk.set_latency_threshold((1 << 30) if variant == "latency" else 0)
for B in [int(a) for a in sys.argv[1:]] or [148]:
    cts = np.ascontiguousarray(np.tile(base, ((B + 63) // 64, 1))[:B])
    idx = np.zeros(B, dtype=np.uint32)
    out = sk.pbs(cts, lut[None], idx)
    ok = all(ck.decrypt_block(out[i]) == (i % 64 % 16 + 1) % 16 for i in range(0, B, max(1, B // 16)))
    sk.kernel_stats(reset=True)
    for _ in range(5):
        sk.pbs(cts, lut[None], idx)
    st = sk.kernel_stats(reset=True)
    print(variant, "B=%d br_ms=%.4f ks_ms=%.4f ok=%s env=%s" % (B, st["br_ms"] / 5, st["ks_ms"] / 5, ok,
          {k: v for k, v in os.environ.items() if k.startswith("FB_")}), flush=True)
sk.close()
