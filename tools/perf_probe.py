#!/usr/bin/env python3
"""Quick GPU perf probe: PBS/s for several batch sizes, keyswitch share, one has_match timing."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fhe_regex_b200 as fb  # noqa: E402


def main():
    ck = fb.ClientKey.load(os.path.join(ROOT, "tests", "golden", "client_key"))
    t = time.time()
    ksk, bsk = fb.keygen_server_raw(ck, seed=0)
    print("keygen %.2fs" % (time.time() - t), flush=True)
    sk = fb.ServerKey(ksk, bsk)
    sk.timing(True)
    lut = fb.make_lut(lambda x: x)
    out = {}
    for B in [int(a) for a in (sys.argv[1:] or ["444", "1776", "4440"])]:
        msgs = np.arange(B) % 16
        cts = ck.encrypt_blocks(msgs[: min(B, 256)], seed=3)
        cts = np.ascontiguousarray(np.tile(cts, ((B + cts.shape[0] - 1) // cts.shape[0], 1))[:B])
        idx = np.zeros(B, dtype=np.uint32)
        sk.pbs(cts[:16], lut[None], idx[:16])
        sk.kernel_stats(reset=True)
        t = time.time()
        res = sk.pbs(cts, lut[None], idx)
        wall = time.time() - t
        st = sk.kernel_stats(reset=True)
        dec = np.array([ck.decrypt_block(c) for c in res[:64]])
        ok = bool((dec == (np.arange(64) % 16)[: len(dec)] % 16).all()) if B >= 64 else None
        out[B] = {"wall_s": wall, "br_ms": st["br_ms"], "ks_ms": st["ks_ms"], "pbs_per_s_kernel": B / ((st["br_ms"] + st["ks_ms"]) / 1e3),
                  "br_tflops": B * 194510848 / (st["br_ms"] / 1e3) / 1e12, "ok": ok}
        print(B, json.dumps(out[B]), flush=True)
    for n, pat in ((64, "/a+b?c/"), (64, "/ab{2,4}c/")):
        content = "".join(np.random.default_rng(1).choice(list("abcx"), size=n))
        ct = fb.encrypt_str(ck, content)
        t = time.time()
        res, st = fb.has_match(sk, ct, pat, return_stats=True)
        wall = time.time() - t
        print(pat, n, "res", ck.decrypt(res), "wall %.3f" % wall, json.dumps(st), flush=True)
    sk.close()


if __name__ == "__main__":
    main()
