#!/usr/bin/env python3
"""One small, ragged invocation of every kernel of the library, for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool racecheck python tools/sanitize_case.py

The blind rotations skip CMUX steps whose (mod-switched) mask element is zero, so the inputs here are small-key LWE
ciphertexts with a dozen non-zero mask elements: every code path of a step runs (stage hand-over, barriers, TMEM, MAC,
transposes), a rotation is 12 steps instead of 742, and the run stays short under the sanitizer.  Outputs are checked
against the CPU oracle (decryption of the bootstrapped value)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fhe_regex_b200 as fb
from oracle import tfhe   # checker only


def main():
    ock = tfhe.ClientKey.load(os.path.join(ROOT, "tests", "golden", "client_key"))
    osk = tfhe.keygen_server(ock, seed=0)
    sk = fb.ServerKey(osk.ksk, osk.bsk)
    rng = np.random.default_rng(1)
    lut = tfhe.make_lut(lambda x: (3 * x + 2) % 16)

    def sparse_small(count, nz=12):
        s = np.zeros((count, 743), dtype=np.uint64)
        for b in range(count):
            pos = rng.choice(742, size=nz, replace=False)
            s[b, pos] = rng.integers(0, 2 ** 64, size=nz, dtype=np.uint64)
            s[b, 742] = rng.integers(0, 2 ** 64, dtype=np.uint64)
        return s

    def check(small, out, pick):
        for b in pick:
            ref = tfhe.bootstrap_small(osk, small[b], lut)
            assert tfhe.decrypt_shortint(ock, out[b]) == tfhe.decrypt_shortint(ock, ref), b

    idx = lambda n: np.zeros(n, dtype=np.uint32)
    # throughput kernels at a ragged count (last CTA holds 3 of 4 samples): phase-by-phase body and fused body
    for variant in (0, 2):
        sk.set_option("br_variant", variant)
        sk.set_latency_threshold(0)
        small = sparse_small(451)
        out = sk.bootstrap_small(small, lut[None], idx(451))
        check(small, out, (0, 3, 448, 450))
    # the fused body's other layouts: 6 samples per CTA (ragged last CTA), one plane for both components, full twiddle tables
    sk.set_option("br_variant", 2)
    for samples, planes, count in ((6, 1, 6 * 148 + 5), (4, 1, 451), (4, 3, 451)):
        sk.set_option("br_samples", samples)
        sk.set_option("br_planes", planes)
        small = sparse_small(count)
        out = sk.bootstrap_small(small, lut[None], idx(count))
        check(small, out, (0, 3, count // 2, count - 1))
    sk.set_option("br_samples", 4)
    sk.set_option("br_planes", 2)
    # 1-3 samples per CTA variants of the phase-by-phase body
    for count in (3, 149, 297):
        small = sparse_small(count)
        out = sk.bootstrap_small(small, lut[None], idx(count))
        check(small, out, (0, count - 1))
    # latency kernel (one PBS per CTA) and cluster kernel (one PBS per pair of CTAs)
    sk.set_latency_threshold(1 << 30)
    small = sparse_small(5)
    check(small, sk.bootstrap_small(small, lut[None], idx(5)), range(5))
    sk.set_cluster_threshold(1 << 30)
    small = sparse_small(3)
    check(small, sk.bootstrap_small(small, lut[None], idx(3)), range(3))
    sk.set_cluster_threshold(0)
    sk.set_latency_threshold(296)
    # keyswitch (decompose + tensor-core GEMM) at a count ragged against its 128-row tile, and the linear glue through a match
    cts = tfhe.encrypt_batch(ock, np.arange(37) % 16, seed=4)
    assert (sk.keyswitch(cts) == tfhe.keyswitch(osk, cts)).all()
    res = fb.has_match(sk, fb.trivial_str("xabbc"), "/ab{2,4}c/")     # trivial ciphertexts: lincomb + KS + (skipped) rotations
    assert fb.ClientKey.load(os.path.join(ROOT, "tests", "golden", "client_key")).decrypt(res) == 1
    sk.close()
    print("sanitize case ok")


if __name__ == "__main__":
    main()
