import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import fhe_regex_b200 as fb
ck = fb.ClientKey.load("tests/golden/client_key")
ksk, bsk = fb.keygen_server_raw(ck, seed=0)
sk = fb.ServerKey(ksk, bsk)
c64 = "".join(np.random.default_rng(5).choice(list("abcx"), size=64))
ct = fb.encrypt_str(ck, c64, seed=9)
# big batch first, like the bench
lut = fb.make_lut(lambda x: x)
base = ck.encrypt_blocks(np.arange(64) % 16, seed=3)
cts = np.ascontiguousarray(np.tile(base, (300, 1)))
sk.pbs(cts, lut[None], np.zeros(len(cts), dtype=np.uint32))
for pat in ["/a+b?c/", "/ab{2,4}c/", r"/[a-d][^x-z]\./", r"/[a-d][^x-z]\./", "/x[ab]+y/"]:
    t = time.perf_counter(); r, st = fb.has_match(sk, ct, pat, return_stats=True); dt = (time.perf_counter() - t) * 1e3
    print(pat, "wall %.1f ms gpu %.1f ms" % (dt, st["gpu_ms"]), fb.plan_level_widths(pat, 64), flush=True)
