#!/usr/bin/env python3
"""One warm regex match per config through fb_has_match (for launch lists / profiles of the match path).
usage: [FB_OPTIONS=name=value,...] match_once.py [n_chars pattern]...   (FB_OPTIONS is read by THIS tool and passed to fb_set_option)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fhe_regex_b200 as fb
from oracle import regex_plain as rp   # checker only
ck = fb.ClientKey.load(os.path.join(ROOT, "tests", "golden", "client_key"))
ksk, bsk = fb.keygen_server_raw(ck, seed=0)
sk = fb.ServerKey(ksk, bsk)
for ov in filter(None, os.environ.get("FB_OPTIONS", "").split(",")):
    sk.set_option(ov.split("=")[0], int(ov.split("=")[1]))
args = sys.argv[1:] or ["64", "/a+b?c/"]
for n, pattern in zip(args[0::2], args[1::2]):
    content = "".join(np.random.default_rng(5).choice(list("abcx"), size=int(n)))
    ct = fb.encrypt_str(ck, content, seed=9)
    fb.has_match(sk, ct, pattern)                       # cold: builds and caches the plan
    ms = 1e9
    for _ in range(3):
        t = time.perf_counter()
        res, st = fb.has_match(sk, ct, pattern, return_stats=True)
        ms = min(ms, (time.perf_counter() - t) * 1e3)
    got, exp = ck.decrypt(res), rp.has_match(content, pattern)
    print("%s on %s chars: %.2f ms wall, %.2f ms gpu, %d PBS in levels %s, result %d (expected %d)" %
          (pattern, n, ms, st["gpu_ms"], st["pbs"], fb.plan_level_widths(pattern, int(n)), got, exp), flush=True)
    assert got == exp
sk.close()
