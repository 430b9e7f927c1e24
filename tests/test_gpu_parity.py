"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

  keyswitch            bit-exact (integer work)
  bootstrap            decrypt-exact + output noise within the stated bound (floating point: the
                       f64 FFT makes ciphertext bits implementation-defined, so parity is phase-level)
  has_match            decrypted 0/1 identical to the reference's semantics (oracle/regex_plain.py) and
                       to the reference's own 25 engine vectors (tests/golden/engine_cases.json)
"""
import ctypes
import json
import os

import numpy as np
import pytest

import fhe_regex_b200 as fb
from oracle import regex_plain as rp
from oracle import tfhe

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
ENGINE_CASES = json.load(open(os.path.join(GOLDEN, "engine_cases.json")))["cases"]

# stated tolerance for one bootstrap output (torus fraction): sigma_out ~ 3e-5;
# |err| < 4e-4 is > 10 sigma, and far below the half-box 1/64 (the box of a 4-bit message + padding bit is 1/32 of the
# torus) that would flip a decryption.  The standard deviation is asserted against the expected bound itself
# (3.7e-5; measured 2.6e-5).
PBS_ERR_MAX = 4e-4
PBS_ERR_STD_MAX = 3.7e-5


@pytest.fixture(scope="module")
def gpu_key(server_key):
    sk = fb.ServerKey(server_key.ksk, server_key.bsk)
    yield sk
    sk.close()


@pytest.fixture(scope="module")
def fck():
    return fb.ClientKey.load(os.path.join(GOLDEN, "client_key"))


def test_keyswitch_bit_exact(client_key, server_key, gpu_key):
    rng = np.random.default_rng(0)
    cts = [tfhe.encrypt_batch(client_key, np.arange(37) % 16, seed=21)]
    cts.append(rng.integers(0, 2 ** 64, size=(20, tfhe.BIG), dtype=np.uint64))       # arbitrary words
    edge = np.zeros((6, tfhe.BIG), dtype=np.uint64)
    edge[0, :] = np.uint64(2 ** 64 - 1)
    edge[1, :] = np.uint64(1 << 48)            # rounding boundary of the decomposer
    edge[2, :] = np.uint64((1 << 48) - 1)
    edge[3, :] = np.uint64((4 << 49) + (4 << 52))  # digit ties
    edge[4, 2048] = np.uint64(7 << 59)         # trivial
    edge[5, ::2] = np.uint64(1 << 63)
    cts.append(edge)
    cts = np.concatenate(cts)
    for count in (1, 7, 8, 9, cts.shape[0]):   # ragged against the 8-sample tile
        got = gpu_key.keyswitch(cts[:count])
        exp = tfhe.keyswitch(server_key, cts[:count])
        assert (got == exp).all(), count


def test_fourier_key_matches_emulation(server_key, gpu_key):
    # the device key conversion against the lane-by-lane CPU emulation of the same kernel code
    import subprocess
    emu_dir = os.path.join(os.path.dirname(__file__), "emu")
    so = os.path.join(emu_dir, "libemu_br.so")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(emu_dir, "emu_br.cpp")):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-march=x86-64-v3", "-fPIC", "-shared", "-o", so, os.path.join(emu_dir, "emu_br.cpp")])
    L = ctypes.CDLL(so)
    exp = np.zeros((742, 2, 2, 1024, 2), dtype=np.float64)
    L.emu_bsk_to_fourier(server_key.bsk.ctypes.data_as(ctypes.c_void_p), exp.ctypes.data_as(ctypes.c_void_p))
    got = gpu_key.fourier_bsk()
    scale = np.abs(exp).max()
    assert np.abs(got - exp).max() < 1e-11 * scale


BR_VARIANTS = {   # name -> (cluster threshold, latency threshold, wide_pair); None = leave the default
    "cluster": (1 << 30, 0, None),       # one PBS per pair of SMs (br_duo.cu)
    "latency": (0, 1 << 30, 0),          # one PBS per CTA (br_wide.cu)
    "pair": (0, 1 << 30, 2),             # two PBS per CTA, twiddles in tensor memory (br_wide2.cu)
    "throughput": (0, 0, None),          # up to 4 PBS per CTA (kernels.cu)
    "default": (None, None, None),       # dispatch by batch size, short tails behind full throughput waves included
}


class _Variant:
    def __init__(self, key, name):
        self.key, self.name = key, name

    def __enter__(self):
        c, l, p = BR_VARIANTS[self.name]
        self.prev_c = self.key.set_cluster_threshold(c) if c is not None else None
        self.prev_l = self.key.set_latency_threshold(l) if l is not None else None
        self.prev_p = self.key.set_option("wide_pair", p) if p is not None else None
        return self.name

    def __exit__(self, *exc):
        if self.prev_c is not None:
            self.key.set_cluster_threshold(self.prev_c)
        if self.prev_l is not None:
            self.key.set_latency_threshold(self.prev_l)
        if self.prev_p is not None:
            self.key.set_option("wide_pair", self.prev_p)


@pytest.fixture(params=list(BR_VARIANTS))
def br_variant(request, gpu_key):
    """the same batches through the three blind rotations and through the default dispatch"""
    with _Variant(gpu_key, request.param) as name:
        yield name


def test_bootstrap_trivial_inputs_bit_exact(server_key, gpu_key, br_variant):
    # trivial ciphertexts (what the reference's tests use, engine.rs:282-286): every CMUX is skipped,
    # the result is pure integer work -> bit-exact against the oracle
    fs = [lambda x: x, lambda x: (3 * x) % 16, lambda x: int(x >= 1)]
    luts = np.stack([tfhe.make_lut(f) for f in fs])
    cts = np.stack([tfhe.trivial_shortint(m) for m in range(16)] * 3)
    idx = np.repeat(np.arange(3), 16)
    got = gpu_key.pbs(cts, luts, idx)
    exp = tfhe.pbs(server_key, cts, luts, idx)
    assert (got == exp).all()


def test_bootstrap_decrypt_and_noise(client_key, server_key, gpu_key):
    n = 444 * 2 + 5                                   # several CTAs, ragged tail
    msgs = np.arange(n) % 16
    cts = tfhe.encrypt_batch(client_key, msgs, seed=31)
    fs = [lambda x: x, lambda x: (x * x) % 16, lambda x: int(x == 5), lambda x: int(x >= 1), lambda x: 15 - x]
    luts = np.stack([tfhe.make_lut(f) for f in fs])
    idx = (np.arange(n) // 16) % len(fs)
    got = gpu_key.pbs(cts, luts, idx)
    exp_msg = np.array([fs[i](int(m)) & 15 for m, i in zip(msgs, idx)], dtype=np.uint64)
    ph = tfhe.phase_batch(client_key.big, got)
    dec = ((ph + np.uint64(1 << 58)) >> np.uint64(59)) & np.uint64(15)
    assert (dec == exp_msg).all()
    err = tfhe.torus_err(ph, exp_msg << np.uint64(59))
    assert np.abs(err).max() < PBS_ERR_MAX and err.std() < PBS_ERR_STD_MAX
    # oracle on a subset of the same inputs: same decryptions, noise of the same order
    sub = slice(0, 48)
    ref = tfhe.pbs(server_key, cts[sub], luts, idx[sub])
    ph_ref = tfhe.phase_batch(client_key.big, ref)
    assert ((((ph_ref + np.uint64(1 << 58)) >> np.uint64(59)) & np.uint64(15)) == dec[sub]).all()
    assert tfhe.torus_err(ph_ref, exp_msg[sub] << np.uint64(59)).std() < PBS_ERR_STD_MAX


def test_bootstrap_stagewise_against_oracle(client_key, server_key, gpu_key, br_variant):
    # same keyswitched inputs into both blind rotations
    msgs = np.arange(32) % 16
    cts = tfhe.encrypt_batch(client_key, msgs, seed=41)
    small = tfhe.keyswitch(server_key, cts)
    lut = tfhe.make_lut(lambda x: (x + 1) % 16)
    got = gpu_key.bootstrap_small(small, lut[None], np.zeros(32, dtype=np.uint32))
    exp_msg = ((msgs + 1) % 16).astype(np.uint64)
    for b in range(32):
        ref = tfhe.bootstrap_small(server_key, small[b], lut)
        assert tfhe.decrypt_shortint(client_key, got[b]) == tfhe.decrypt_shortint(client_key, ref) == int(exp_msg[b])
    err = tfhe.torus_err(tfhe.phase_batch(client_key.big, got), exp_msg << np.uint64(59))
    assert np.abs(err).max() < PBS_ERR_MAX


@pytest.mark.parametrize("count", [1, 2, 74, 75, 148, 149, 296, 297, 444, 445, 592, 593, 620, 1185])
def test_bootstrap_batch_size_boundaries(count, fck, gpu_key, br_variant):
    # the throughput blind rotation picks 1..4 samples per SM from the batch size (ragged last CTAs at every
    # boundary); the latency one runs in waves of one CTA per PBS
    if br_variant in ("latency", "pair", "cluster") and count > 445:
        pytest.skip("the narrow-level variants are never chosen for wide batches")
    msgs = (np.arange(count) * 7 + 3) % 16
    base = fck.encrypt_blocks(msgs[:min(count, 96)], seed=77)
    cts = np.ascontiguousarray(np.tile(base, ((count + 95) // 96, 1))[:count])
    msgs = np.tile(msgs[:min(count, 96)], (count + 95) // 96)[:count]
    fs = [lambda x: (x + 5) % 16, lambda x: int(x >= 8)]
    luts = np.stack([fb.make_lut(f) for f in fs])
    idx = (np.arange(count) % 2).astype(np.uint32)
    out = gpu_key.pbs(cts, luts, idx)
    pick = sorted(set([0, count - 1, count // 2] + list(range(max(0, count - 5), count))))
    for i in pick:
        assert fck.decrypt_block(out[i]) == fs[int(idx[i])](int(msgs[i])) & 15, (count, i)


def test_blind_rotation_variants_agree(client_key, gpu_key):
    """same keyswitched inputs through both blind rotations: same decryptions, outputs within FFT rounding"""
    n = 150
    msgs = np.arange(n) % 16
    cts = tfhe.encrypt_batch(client_key, msgs, seed=53)
    fs = [lambda x: (x * 3 + 2) % 16, lambda x: int(x == 2)]
    luts = np.stack([tfhe.make_lut(f) for f in fs])
    idx = (np.arange(n) % 2).astype(np.uint32)
    outs = {}
    for name in ("cluster", "latency", "pair", "throughput"):
        with _Variant(gpu_key, name):
            outs[name] = gpu_key.pbs(cts, luts, idx)
    # the pair kernel runs the stages of the latency kernel with the same twiddles: not one bit differs
    assert (outs["pair"] == outs["latency"]).all()
    exp_msg = np.array([fs[i](int(m)) & 15 for m, i in zip(msgs, idx)], dtype=np.uint64)
    for name, got in outs.items():
        ph = tfhe.phase_batch(client_key.big, got)
        dec = ((ph + np.uint64(1 << 58)) >> np.uint64(59)) & np.uint64(15)
        assert (dec == exp_msg).all(), name
        err = tfhe.torus_err(ph, exp_msg << np.uint64(59))
        assert np.abs(err).max() < PBS_ERR_MAX and err.std() < PBS_ERR_STD_MAX, name
    # the mask words are not comparable (a last-bit difference of an f64 rounding changes later digits, i.e. the
    # noise realisation), the phases are: both are encryptions of the same value with noise of the same size
    ph_t = tfhe.phase_batch(client_key.big, outs["throughput"])
    for name in ("cluster", "latency", "pair"):
        assert np.abs(tfhe.torus_err(tfhe.phase_batch(client_key.big, outs[name]), ph_t)).max() < 2 * PBS_ERR_MAX, name


@pytest.mark.parametrize("quanta,extra", [(8, 1), (11, 300), (20, 37), (24, 0), (25, 311)])
def test_pbs_batch_pipelined_chunks(quanta, extra, fck, gpu_key):
    """fb_pbs_batch splits batches beyond 8 throughput quanta into two, three or (from 24 quanta on) five chunks whose copies
    overlap the bootstraps: every chunk, both sides of every possible chunk boundary and the ragged tail must come back right"""
    q = gpu_key.pbs_quantum()
    count = quanta * q + extra
    base_msgs = (np.arange(128) * 5 + 1) % 16
    base = fck.encrypt_blocks(base_msgs, seed=91)
    cts = np.ascontiguousarray(np.tile(base, ((count + 127) // 128, 1))[:count])
    msgs = np.tile(base_msgs, (count + 127) // 128)[:count]
    fs = [lambda x: (x + 9) % 16, lambda x: int(x < 4), lambda x: (3 * x) % 16]
    luts = np.stack([fb.make_lut(f) for f in fs])
    idx = (np.arange(count) % 3).astype(np.uint32)
    prev = gpu_key.set_option("pbs_chunks", 5 if quanta >= 24 else 3)
    try:
        out = gpu_key.pbs(cts, luts, idx)
    finally:
        gpu_key.set_option("pbs_chunks", prev)
    pick = {0, 1, count - 1, count - 2, count // 2} | set(range(0, count, 997))
    for b in range(q, count, q):                       # chunk boundaries are multiples of the quantum
        pick |= {b - 1, b, min(b + 1, count - 1)}
    for i in sorted(pick):
        assert fck.decrypt_block(out[i]) == fs[int(idx[i])](int(msgs[i])) & 15, (count, i)


def test_empty_batches(gpu_key):
    assert gpu_key.keyswitch(np.zeros((0, tfhe.BIG), dtype=np.uint64)).shape == (0, tfhe.SMALL)
    lut = tfhe.make_lut(lambda x: x)
    assert gpu_key.pbs(np.zeros((0, tfhe.BIG), dtype=np.uint64), lut[None], np.zeros(0, dtype=np.uint32)).shape == (0, tfhe.BIG)


@pytest.mark.parametrize("case", ENGINE_CASES, ids=["%s~%s" % (c["content"], c["pattern"]) for c in ENGINE_CASES])
def test_has_match_reference_vectors_trivial(case, fck, gpu_key):
    # exactly the reference's test_has_match (engine.rs:256-291): trivial content, decrypt, compare
    ct = fb.trivial_str(case["content"])
    res = fb.has_match(gpu_key, ct, case["pattern"])
    assert fck.decrypt(res) == case["expected"]


RANDOM_PATTERNS = ["/a+b?c/", "/ab{2,4}c/", r"/[a-d][^x-z]\./", "/abc/i", "/^a*b+$/", "/(ab|c)+x/", "/[^ab]+c/", "/a{,2}b/",
                   "/[a-c]{2,}x?$/", "/^.a.$/", "/a|b|c/", "/[abc][^a-b]c{2}/", "/x[ab]+y/i", "/aaaa+/", "/a.+b/", "/(a|b)?c$/"]


def test_has_match_random_patterns_on_the_device(fck, gpu_key):
    """the device executes what the plaintext dry run of the plan promises: random (pattern, content) pairs on trivial
    ciphertexts (the reference's own test style, engine.rs:282-286: no CMUX runs, a match takes milliseconds), singly
    and as has_match_many batches, against the oracle"""
    import random
    rnd = random.Random(23)
    for it in range(60):
        pat = rnd.choice(RANDOM_PATTERNS)
        max_n = 8 if "|" in pat and ("+" in pat or "*" in pat) else 40
        n = rnd.randint(1, max_n)
        contents = ["".join(rnd.choice("abcxyAB.") for _ in range(n)) for _ in range(rnd.randint(1, 5))]
        exp = [rp.has_match(c, pat) for c in contents]
        assert fck.decrypt(fb.has_match(gpu_key, fb.trivial_str(contents[0]), pat)) == exp[0], (pat, contents[0])
        outs = fb.has_match_many(gpu_key, np.stack([fb.trivial_str(c) for c in contents]), pat)
        assert [fck.decrypt(o) for o in outs] == exp, (pat, contents)


REAL_CASES = [
    ("abc", "/^abc$/"), ("abd", "/^abc$/"),
    ("aBc", "/^abc$/i"), ("aBc" + "x" * 13, "/^abc$/i"), ("xxaBcxxxxxxxxxxx", "/abc/i"),
    ("bq.", r"/^[a-d][^x-z]\.$/"), ("aq.", r"/^[a-d][^x-z]\.$/"), ("bx.", r"/^[a-d][^x-z]\.$/"), ("by.", r"/^[a-d][^x-z]\.$/"),
    ("abbc", "/^ab{2,4}c$/"), ("abbbbc", "/^ab{2,4}c$/"), ("abc", "/^ab{2,4}c$/"), ("abbbbbc", "/^ab{2,4}c$/"),
    ("aaa", "/^a{,2}$/"), ("", "/^$/"), ("zzz", "/./"), ("zzz", "/^.*$/"),
]


@pytest.mark.parametrize("content,pattern", REAL_CASES, ids=["%s~%s" % c for c in REAL_CASES])
def test_has_match_real_encryption(content, pattern, fck, gpu_key):
    ct = fb.encrypt_str(fck, content, seed=5)
    res, st = fb.has_match(gpu_key, ct, pattern, return_stats=True)
    exp, ex, nb = rp.has_match(content, pattern, return_exec=True)
    assert fck.decrypt(res) == exp
    assert (st["variants"], st["ct_ops"], st["cache_hits"]) == (nb, ex.ct_ops, ex.cache_hits)
    assert (res[1:, :] == 0).all()  # blocks 1-3 trivial zero


def test_has_match_64_char_configs(fck, gpu_key):
    rng = np.random.default_rng(5)
    c64 = "".join(rng.choice(list("abcx"), size=64))
    planted = c64[:20] + "xaabc" + c64[25:]
    for content, pattern in [(c64, r"/[a-d][^x-z]\./"), (planted, "/ab{2,4}c/"), (c64, "/ab{2,4}c/"), (planted, "/a+b?c/"),
                             ("x" * 64, "/a+b?c/"), (c64, r"/^[a-d][^x-z]\.$/"), (c64, "/^ab{2,4}c$/")]:
        res = fb.has_match(gpu_key, fb.encrypt_str(fck, content, seed=9), pattern)
        assert fck.decrypt(res) == rp.has_match(content, pattern), (content, pattern)
    # BASELINE config 4 read literally: an anchored pattern of at most 6 characters against 64 characters has no
    # variant at all (engine.rs:59-65): 0 ciphertext operations, a trivial encryption of false
    res, st = fb.has_match(gpu_key, fb.encrypt_str(fck, c64, seed=9), "/^ab{2,4}c$/", return_stats=True)
    assert (st["variants"], st["ct_ops"], st["pbs"]) == (0, 0, 0) and fck.decrypt(res) == 0 and (res[:, :2048] == 0).all()


def test_has_match_256_char_config5_both_plans(fck, gpu_key):
    # BASELINE config 5 at full size: 65 025 variants; the reference-shaped plan evaluates all of them (~76k PBS),
    # the default plan absorbs implied variants (~1.6k PBS); both must decrypt to the reference's result
    rng = np.random.default_rng(12)
    miss = "".join(rng.choice(list("abx"), size=256))                    # no 'c': no match
    hit = miss[:200] + "aabc" + miss[204:]
    for content in (miss, hit):
        ct = fb.encrypt_str(fck, content, seed=13)
        exp = rp.has_match(content, "/a+b?c/")
        res, st = fb.has_match(gpu_key, ct, "/a+b?c/", return_stats=True)
        assert fck.decrypt(res) == exp and st["pbs"] < 3000 and st["variants"] == 65025
    prev = gpu_key.set_option("plan_reference_shaped", 1)
    try:
        res, st = fb.has_match(gpu_key, fb.encrypt_str(fck, hit, seed=13), "/a+b?c/", return_stats=True)
    finally:
        gpu_key.set_option("plan_reference_shaped", prev)
    assert fck.decrypt(res) == 1 and st["pbs"] > 60000 and (st["ct_ops"], st["cache_hits"]) == (195583, 11118596)


def test_has_match_many_contents_share_the_launches(fck, gpu_key):
    """m contents against one pattern in shared launches: every result is what has_match gives for that content alone"""
    cases = [("/ab{2,4}c/", ["xabbcx", "xabcxx", "abbbbc", "cbbbba", "abbbbb", "aabbcc", "xxxxxx"]),
             ("/a+b?c/", ["aaabcxxxxxxxxxxxxxxx", "xxxxxxxxxxxxxxxxxxac", "bbbbbbbbbbbbbbbbbbbb", "aaaaaaaaaaaaaaaaaaab", "cabcabxxxxxxxxxxxxxx"]),
             ("/^abc$/", ["abc", "abd", "xbc"])]
    for pattern, contents in cases:
        cts = np.stack([fb.encrypt_str(fck, c, seed=40 + i) for i, c in enumerate(contents)])
        outs, st = fb.has_match_many(gpu_key, cts, pattern, return_stats=True)
        assert outs.shape == (len(contents), 4, tfhe.BIG)
        got = [fck.decrypt(o) for o in outs]
        assert got == [rp.has_match(c, pattern) for c in contents], (pattern, got)
        assert st["pbs"] == fb.plan_stats(pattern, len(contents[0]))["pbs"]
    # degenerate shapes: no content at all, a plan that is a constant
    assert fb.has_match_many(gpu_key, np.zeros((0, 3, 4, tfhe.BIG), dtype=np.uint64), "/abc/").shape == (0, 4, tfhe.BIG)
    two = np.stack([fb.encrypt_str(fck, "ab"), fb.encrypt_str(fck, "cd")])
    assert [fck.decrypt(o) for o in fb.has_match_many(gpu_key, two, "/^abc$/")] == [0, 0]


def test_sharded_match_and_or_fold(fck, gpu_key):
    content = "xxabbcxxxxaacxxx"
    ct = fb.encrypt_str(fck, content, seed=2)
    for pattern in ["/a+b?c/", "/^zz/", "/ab{2,4}c/"]:
        world = 4
        parts = np.stack([fb.has_match(gpu_key, ct, pattern, rank=r, world=world)[0] for r in range(world)])
        assert fck.decrypt(gpu_key.or_fold(parts)) == rp.has_match(content, pattern)


def test_errors_surface_like_the_reference(fck, gpu_key):
    ct = fb.trivial_str("ab")
    with pytest.raises(fb.ParseError):
        fb.has_match(gpu_key, ct, "/a(b/")
    with pytest.raises(fb.ReferencePanic):
        fb.has_match(gpu_key, ct, "/^/")


def test_bootstrap_noise_many_trials():
    """North-star criterion: no decryption failure and output noise within the bound over 10^6 fresh
    encryptions (about half a minute on a B200 box; FB_NOISE_TRIALS overrides).  A recorded run is in
    profiles/r01_noise_1e6_trials.json."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("noise_trials", os.path.join(os.path.dirname(__file__), "..", "tools", "noise_trials.py"))
    nt = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(nt)
    res = nt.run(int(os.environ.get("FB_NOISE_TRIALS", "1000000")))
    assert res["decryption_failures"] == 0, res
    assert res["err_std"] < PBS_ERR_STD_MAX and res["err_abs_max"] < PBS_ERR_MAX, res
    # the latency kernel rounds differently (Stockham stages instead of the 32x32 transform): same criterion on a
    # smaller sample (recorded 10^6-trial runs: profiles/r01_noise_1e6_trials_{latency,cluster}.json)
    for variant in ("latency", "cluster"):
        res = nt.run(int(os.environ.get("FB_NOISE_TRIALS_NARROW", "100000")), variant=variant)
        assert res["decryption_failures"] == 0, res
        assert res["err_std"] < PBS_ERR_STD_MAX and res["err_abs_max"] < PBS_ERR_MAX, res


def test_bootstrap_noise_worst_case_input():
    """The same criterion on the noisiest input of the match path: a sum of 15 bootstrapped booleans, keyswitched and
    bootstrapped through the x == k / x >= 1 LUTs (FB_NOISE_TRIALS_WORST overrides the trial count; a recorded
    10^6-trial run is in profiles/r02_noise_worstcase_1e6.json)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("noise_trials", os.path.join(os.path.dirname(__file__), "..", "tools", "noise_trials.py"))
    nt = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(nt)
    res = nt.run_worst_case(int(os.environ.get("FB_NOISE_TRIALS_WORST", "200000")), pool=10 * 28416)
    assert res["decryption_failures"] == 0, res
    assert res["err_std"] < PBS_ERR_STD_MAX and res["err_abs_max"] < PBS_ERR_MAX, res
    assert min(res["input_sum_histogram"][3:13]) > 0      # the sums really spread over the boxes


def test_handle_api_replays_has_match_level_by_level(fck, gpu_key):
    """The op-level boundary (fb_ct_alloc / fb_ct_upload / fb_lincomb / fb_pbs_rows / fb_ct_download) driven from the host
    with the library's own plan (fb_plan_export): a host that keeps the reference's Execution (execution.rs:64-222) flushes
    its levels exactly like this.  Same kernels on the same rows: the result row is bit-identical to fb_has_match's."""
    luts = fb.regex_lut_table()
    for content, pattern in (("xabbcx", "/ab{2,4}c/"), ("bq.", r"/^[a-d][^x-z]\.$/"), ("xxaBcxxxxxxxxxxx", "/abc/i")):
        ct = fb.encrypt_str(fck, content, seed=5)
        ref = fb.has_match(gpu_key, ct, pattern)
        plan = fb.plan_export(pattern, len(content))
        assert plan["result_kind"] == 2
        arena = gpu_key.ct_alloc(plan["n_rows"])
        lut_h = gpu_key.ct_alloc(luts.shape[0], 2048)
        try:
            gpu_key.ct_upload(lut_h, 0, luts)
            gpu_key.ct_upload(arena, 0, ct.reshape(-1, 2049))
            for lv in plan["levels"]:
                gpu_key.lincomb(arena, lv["lin_out_rows"], lv["lin_term_off"], lv["lin_term_rows"], lv["lin_coef"], lv["lin_const"])
                gpu_key.pbs_rows(arena, lv["in_rows"], lut_h, lv["lut_idx"], lv["out_row_base"])
            got = gpu_key.ct_download(arena, plan["result_row"], 1)[0]
        finally:
            gpu_key.ct_free(arena)
            gpu_key.ct_free(lut_h)
        assert (got == ref[0]).all()
        assert fck.decrypt_block(got) == rp.has_match(content, pattern)
    # argument checking: rows out of range, wrong kind of arena
    arena = gpu_key.ct_alloc(4)
    with pytest.raises(fb.FbError):
        gpu_key.pbs_rows(arena, [0], arena, [0], 0)              # LUT handle must hold 2048-word rows
    with pytest.raises(fb.FbError):
        gpu_key.lincomb(arena, [4], [0, 1], [0], [1], [0])       # output row out of range
    with pytest.raises(fb.FbError):
        gpu_key.ct_download(arena, 3, 2)
    gpu_key.ct_free(arena)
    with pytest.raises(fb.FbError):
        gpu_key.ct_free(arena)


def test_every_kernel_on_a_small_ragged_case():
    """tools/sanitize_case.py: sparse-mask inputs (12 CMUX steps per rotation) through every kernel at ragged batch sizes,
    checked against the oracle -- the case compute-sanitizer was to run (the tool is closed on this pool:
    profiles/r02_sanitizer_closed.txt)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("sanitize_case", os.path.join(os.path.dirname(__file__), "..", "tools", "sanitize_case.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.main()


def test_server_keygen_on_the_gpu(client_key, fck):
    """fb_keygen_server_gpu = ServerKey::new(&client_key) (engine.rs:252) in CUDA kernels.  The generated key is checked
    row by row against the secret keys (plaintexts exact, noise of the parameter set's standard deviations), the keyswitch
    through it is bit-exact against the oracle run on the downloaded key, and bootstraps / a match decrypt correctly."""
    sk = fb.ServerKey(keygen_from=fck, seed=11, keep_generated=True)
    try:
        ksk, bsk = sk.generated
        small = fck.small.astype(np.uint64)
        big = fck.big.astype(np.uint64)
        with np.errstate(over="ignore"):
            # KSK row (i, l): phase = body - <mask, s> = big[i] * 2^(64 - 3(l+1)) + e, e ~ N(0, sigma_lwe)
            phase = ksk[:, :, 742] - (ksk[:, :, :742] * small[None, None, :]).sum(axis=2, dtype=np.uint64)
            want = big[:, None] << np.array([61, 58, 55, 52, 49], dtype=np.uint64)[None, :]
            err = (phase - want).view(np.int64).astype(np.float64) / 2.0 ** 64
        assert np.abs(err).max() < 6 * 7.07e-6 and 0.9 * 7.07e-6 < err.std() < 1.1 * 7.07e-6, (err.std(), np.abs(err).max())
        assert abs(err.mean()) < 4 * 7.07e-6 / np.sqrt(err.size)
        # masks are fresh: no two rows share their first words
        assert len({int(x) for x in ksk[:, :, 0].reshape(-1)}) == 2048 * 5
        # BSK GGSW rows: B - A * S = plaintext + e, checked exactly (integer negacyclic product) on a few rows
        ones = np.nonzero(big)[0]
        for i in (0, 1, 371, 741):
            for r in (0, 1):
                A, B = bsk[i, 0, r, 0], bsk[i, 0, r, 1]
                prod = np.zeros(2048, dtype=np.uint64)
                with np.errstate(over="ignore"):
                    for t in ones:
                        rolled = np.roll(A, t)
                        rolled[:t] = np.uint64(0) - rolled[:t]
                        prod += rolled
                    factor = np.uint64(int(small[i]) << 41)
                    pt = (np.uint64(0) - factor * big) if r == 0 else np.concatenate([[factor], np.zeros(2047, dtype=np.uint64)]).astype(np.uint64)
                    e = (B - prod - pt).view(np.int64).astype(np.float64) / 2.0 ** 64
                assert np.abs(e).max() < 6 * 2.95e-16 and 0.85 * 2.94e-16 < e.std() < 1.15 * 2.94e-16, (i, r, e.std())
        # keyswitch through the installed key == oracle keyswitch with the downloaded key, bit for bit
        osk = tfhe.ServerKey(ksk, bsk)
        msgs = np.arange(16)
        cts = tfhe.encrypt_batch(client_key, msgs, seed=41)
        assert (sk.keyswitch(cts) == tfhe.keyswitch(osk, cts)).all()
        lut = fb.make_lut(lambda x: (5 * x + 7) % 16)
        out = sk.pbs(cts, lut[None], np.zeros(16, dtype=np.uint32))
        assert [fck.decrypt_block(c) for c in out] == [(5 * m + 7) % 16 for m in msgs]
        ph = tfhe.phase_batch(client_key.big, out)
        errb = tfhe.torus_err(ph, np.array([((5 * m + 7) % 16) << 59 for m in msgs], dtype=np.uint64))
        assert np.abs(errb).max() < PBS_ERR_MAX
        assert fck.decrypt(fb.has_match(sk, fb.encrypt_str(fck, "xxabbbcx", seed=2), "/ab{2,4}c/")) == 1
    finally:
        sk.close()
    # a different seed gives a different key
    sk2 = fb.ServerKey(keygen_from=fck, seed=12, keep_generated=True)
    assert (sk2.generated[0][0, 0, :8] != ksk[0, 0, :8]).any()
    sk2.close()


def test_demo_entry_point_prints_res(capsys):
    """`python -m fhe_regex_b200 <content> <pattern>` = the reference's `cargo run -- <content> <pattern>` (src/main.rs:9-23,
    src/regex/mod.rs:9-19): keygen, encrypt_str, has_match, decrypt, `res: 0|1` on stdout."""
    from fhe_regex_b200.__main__ import main
    assert main(["fhe-regex", "abc", "/^abc$/"]) == 0
    assert capsys.readouterr().out.strip().splitlines()[-1] == "res: 1"
    assert main(["fhe-regex", "abd", "/^abc$/"]) == 0
    assert capsys.readouterr().out.strip().splitlines()[-1] == "res: 0"
    assert main(["fhe-regex", "abc"]) == 2
    with pytest.raises(fb.ParseError):
        main(["fhe-regex", "abc", "/a(b/"])


def test_has_match_dist_single_rank_communicator(fck, server_key):
    """fb_comm_init + fb_has_match_dist with a communicator of one rank (what a 1-GPU box can run; the multi-rank path is
    exercised by tools/dist_match_check.py under torchrun, profiles/r02_dist_n*.log): same decryption as fb_has_match."""
    sk = fb.ServerKey(server_key.ksk, server_key.bsk)
    try:
        sk.comm_init(fb.comm_unique_id(), 0, 1)
        for content, pattern in (("xabbcx", "/ab{2,4}c/"), ("aq.", r"/^[a-d][^x-z]\.$/"), ("", "/^$/")):
            ct = fb.encrypt_str(fck, content, seed=5)
            res, st = fb.has_match_dist(sk, ct, pattern, return_stats=True)
            assert fck.decrypt(res) == rp.has_match(content, pattern) == fck.decrypt(fb.has_match(sk, ct, pattern))
    finally:
        sk.close()
    with pytest.raises(fb.FbError):
        fb.has_match_dist(sk2 := fb.ServerKey(server_key.ksk, server_key.bsk), fb.trivial_str("a"), "/a/")   # no communicator
    sk2.close()


def test_server_key_in_the_references_own_form(client_key, server_key, gpu_key, fck):
    """A tfhe-rs 0.2.0 ServerKey holds the bootstrapping key in the Fourier domain only (engine.rs:252,
    ciphertext.rs:44).  Feed the key in that domain / serialized order -- computed by the CPU oracle, independently of
    the GPU's own conversion -- through fb_load_server_key_fourier and through the bincode blob of integer::ServerKey:
    the resident key is bit-for-bit what was handed over, agrees with fb_load_server_key_raw's conversion of the same
    key to f64 rounding, and bootstraps / matches identically."""
    from test_wire_formats import natural_fourier_key
    fk = natural_fourier_key(server_key)
    own = gpu_key.fourier_bsk()                       # fb_load_server_key_raw: converted on the device
    scale = np.abs(fk).max()
    assert np.abs(own - fk).max() < 1e-12 * scale      # same transform, same order, same scaling
    blob = fb.server_key_to_bincode(server_key.ksk, fk)
    msgs = np.arange(16)
    cts = tfhe.encrypt_batch(client_key, msgs, seed=31)
    lut = fb.make_lut(lambda x: (7 * x + 3) % 16)
    ref_ks = gpu_key.keyswitch(cts)
    for kwargs in ({"fourier_bsk": fk}, {"bincode": blob}):
        sk = fb.ServerKey(None if "bincode" in kwargs else server_key.ksk, **kwargs)
        try:
            assert (sk.fourier_bsk() == fk).all()      # a copy: no conversion on this path
            assert (sk.keyswitch(cts) == ref_ks).all()
            out = sk.pbs(cts, lut[None], np.zeros(16, dtype=np.uint32))
            assert [fck.decrypt_block(c) for c in out] == [(7 * m + 3) % 16 for m in msgs]
            ct = fb.string_ciphertext_from_bincode(fb.string_ciphertext_to_bincode(fb.encrypt_str(fck, "xabbbc", seed=3)))
            assert fck.decrypt(fb.has_match(sk, ct, "/ab{2,4}c/")) == 1
        finally:
            sk.close()
    with pytest.raises(fb.FbError):
        fb.ServerKey(bincode=blob[:-1])


def test_options_are_per_context_and_checked(gpu_key, server_key):
    """fb_set_option / fb_get_option: defaults, round trip, unknown names and out-of-range values are errors, and a second
    context keeps its own values"""
    assert gpu_key.get_option("br_variant") == 2 and gpu_key.get_option("br_planes") == 2 and gpu_key.get_option("br_samples") == 4 and gpu_key.get_option("ks_variant") == 1 and gpu_key.get_option("latency_threshold") == 296
    prev = gpu_key.set_option("latency_threshold", 100)
    assert prev == 296 and gpu_key.get_option("latency_threshold") == 100
    other = fb.ServerKey(server_key.ksk, server_key.bsk)
    try:
        assert other.get_option("latency_threshold") == 296
    finally:
        other.close()
    gpu_key.set_option("latency_threshold", prev)
    for name, value in (("no_such_option", 1), ("br_variant", 9), ("ks_variant", -1)):
        with pytest.raises(fb.FbError):
            gpu_key.set_option(name, value)


def test_keyswitch_variants_agree_bit_for_bit(client_key, server_key, gpu_key):
    """the tcgen05 GEMM (ks_umma.cu) and the mma.sync GEMM (ks_kernels.cu) are the same exact integer contraction"""
    rng = np.random.default_rng(3)
    cts = rng.integers(0, 2 ** 64, size=(300, tfhe.BIG), dtype=np.uint64)      # 300: ragged against both tile heights
    outs = []
    for v in (0, 1):
        prev = gpu_key.set_option("ks_variant", v)
        outs.append(gpu_key.keyswitch(cts))
        gpu_key.set_option("ks_variant", prev)
    assert (outs[0] == outs[1]).all()
    assert (outs[1][:40] == tfhe.keyswitch(server_key, cts[:40])).all()


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4])
def test_throughput_blind_rotation_variants(variant, client_key, server_key, gpu_key, fck):
    """every body of the throughput blind rotation (phase by phase, fused, fused + I2F digits, + tensor-memory key) on a
    batch with a ragged last CTA: decrypt-exact, error within the stated bound"""
    n = 4 * 148 + 3
    msgs = np.arange(n) % 16
    base = tfhe.encrypt_batch(client_key, msgs[:64], seed=51)
    cts = np.ascontiguousarray(np.tile(base, ((n + 63) // 64, 1))[:n])
    f = lambda x: (9 * x + 4) % 16
    lut = fb.make_lut(f)
    prev = (gpu_key.set_option("br_variant", variant), gpu_key.set_latency_threshold(0))
    try:
        out = gpu_key.pbs(cts, lut[None], np.zeros(n, dtype=np.uint32))
    finally:
        gpu_key.set_option("br_variant", prev[0])
        gpu_key.set_latency_threshold(prev[1])
    exp = np.array([f(int(m % 64 % 16)) for m in range(n)], dtype=np.uint64)
    ph = tfhe.phase_batch(client_key.big, out)
    err = tfhe.torus_err(ph, exp << np.uint64(59))
    assert [fck.decrypt_block(out[i]) for i in (0, 1, 63, 64, n - 2, n - 1)] == [int(exp[i]) for i in (0, 1, 63, 64, n - 2, n - 1)]
    assert np.abs(err).max() < PBS_ERR_MAX and err.std() < PBS_ERR_STD_MAX, (variant, err.std())


def test_throughput_blind_rotation_layouts_bit_identical(client_key, gpu_key, fck):
    """round 2, second half: the fused throughput kernel with 6 PBS per CTA (transpose planes inside the accumulator copies),
    with a plane per component (one barrier per transpose), with full twiddle tables built once per launch, and with the two
    unneeded barriers back -- all the same arithmetic in the same order: outputs bit-identical to the default layout"""
    n = 6 * 148 + 5
    msgs = np.arange(n) % 16
    base = tfhe.encrypt_batch(client_key, msgs[:64], seed=52)
    cts = np.ascontiguousarray(np.tile(base, ((n + 63) // 64, 1))[:n])
    f = lambda x: (7 * x + 2) % 16
    lut = fb.make_lut(f)
    idx = np.zeros(n, dtype=np.uint32)
    names = ("br_variant", "br_samples", "br_planes", "br_barriers", "br_stagger", "br_stagger_groups", "br_sync", "br_resync")
    saved = [gpu_key.get_option(k) for k in names]
    prev_lat = gpu_key.set_latency_threshold(0)
    try:
        outs = {}
        for cfg in ((1, 4, 1, 0, 0, 0, 1, 0), (1, 4, 1, 1, 0, 0, 1, 8), (1, 4, 2, 0, 0, 0, 1, 8), (1, 4, 3, 0, 0, 0, 1, 8), (1, 6, 1, 0, 0, 0, 1, 8),
                    (1, 6, 1, 0, 3000, 1, 1, 0), (2, 4, 1, 0, 0, 0, 1, 0), (2, 4, 2, 0, 0, 0, 1, 8), (2, 4, 3, 0, 0, 0, 1, 8), (2, 6, 1, 0, 0, 0, 1, 8),
                    (2, 4, 2, 0, 0, 0, 0, 0), (2, 4, 2, 0, 0, 0, 1, 1), (2, 6, 1, 0, 0, 0, 1, 3), (2, 4, 2, 0, 0, 0, 0, 63)):
            for k, v in zip(names, cfg):
                gpu_key.set_option(k, v)
            outs[cfg] = gpu_key.pbs(cts, lut[None], idx)
    finally:
        for k, v in zip(names, saved):
            gpu_key.set_option(k, v)
        gpu_key.set_latency_threshold(prev_lat)
    for cfg, out in outs.items():
        ref = outs[(cfg[0], 4, 1, 0, 0, 0, 1, 0)]
        assert (out == ref).all(), cfg
    out = outs[(1, 6, 1, 0, 0, 0, 1, 8)]
    for i in (0, 1, 63, 64, 591, 592, 887, 888, n - 1):
        assert fck.decrypt_block(out[i]) == f(int(msgs[i % 64]))
