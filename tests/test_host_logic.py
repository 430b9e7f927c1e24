"""Host side of the product (C++ parser / variant generator / executor bookkeeping / PBS lowering)
through the C ABI, checked against the oracle and the reference's golden vectors.  No GPU needed."""
import ctypes
import json
import os
import random
import re
import subprocess

import numpy as np
import pytest

import fhe_regex_b200 as fb
from oracle import regex_plain as rp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
PARSER_CASES = json.load(open(os.path.join(GOLDEN, "parser_cases.json")))["cases"]
ENGINE_CASES = json.load(open(os.path.join(GOLDEN, "engine_cases.json")))["cases"]


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "fhe_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(fb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    lib = fb.lib()
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = fb.lib().fb_ctx_create(ctypes.byref(h), 0)
    assert rc == fb.FB_ERR_NO_DEVICE and not h.value
    with pytest.raises(fb.FbError):
        fb.ServerKey.__new__(fb.ServerKey).__init__(__import__("numpy").zeros(fb.KSK_WORDS, dtype="uint64"),
                                                     __import__("numpy").zeros(fb.BSK_WORDS, dtype="uint64"))


@pytest.mark.parametrize("case", PARSER_CASES, ids=[c["pattern"] for c in PARSER_CASES])
def test_parser_golden(case):
    assert fb.parse(case["pattern"]) == rp.debug_fmt(case["ast"])


@pytest.mark.parametrize("pattern", ["abc", "/abc", "/abc/x", "/a(b/", "/[ab/", "/a**/", "/+/", "/a{1,2/"])
def test_parse_errors(pattern):
    with pytest.raises(fb.ParseError):
        fb.parse(pattern)
    with pytest.raises(rp.ParseError):
        rp.parse(pattern)


def test_reference_panics():
    with pytest.raises(fb.ReferencePanic):
        fb.parse("/a{}/")
    assert fb.parse("/^/") == "<^<>>"
    with pytest.raises(fb.ReferencePanic):
        fb.plan_stats("/^/", 1)


@pytest.mark.parametrize("case", ENGINE_CASES, ids=["%s~%s" % (c["content"], c["pattern"]) for c in ENGINE_CASES])
def test_engine_golden_dry_run(case):
    assert fb.plan_eval_plain(case["pattern"], case["content"]) == case["expected"]


COUNT_ROWS = [
    ("abc", "/^abc$/", 1, 5, 0),
    ("aBc", "/^abc$/i", 1, 11, 0),
    ("aBc" + "x" * 13, "/^abc$/i", 0, 0, 0),
    ("x" * 13 + "aBc", "/abc/i", 14, 167, 0),
    ("bq.", r"/^[a-d][^x-z]\.$/", 1, 10, 0),
    ("q" * 64, r"/^[a-d][^x-z]\.$/", 0, 0, 0),
    ("q" * 64, r"/[a-d][^x-z]\./", 62, 681, 0),
    ("abbbc", "/^ab{2,4}c$/", 1, 9, 0),
    ("q" * 64, "/ab{2,4}c/", 180, 903, 892),
    ("q" * 64, "/a+b?c/", 3969, 12031, 170500),
    ("q" * 256, "/a+b?c/", 65025, 195583, 11118596),
]


@pytest.mark.parametrize("row", COUNT_ROWS, ids=["%d:%s" % (len(r[0]), r[1]) for r in COUNT_ROWS])
def test_counters_match_reference_bookkeeping(row):
    # what engine.rs:36-40 would log; BASELINE.md section 2
    content, pattern, variants, ops, hits = row
    st = fb.plan_stats(pattern, len(content))
    assert (st["variants"], st["ct_ops"], st["cache_hits"]) == (variants, ops, hits)
    if len(content) <= 64:
        _, ex, nb = rp.has_match(content, pattern, return_exec=True)
        assert (nb, ex.ct_ops, ex.cache_hits) == (variants, ops, hits)
        by = ex.ops_by_type
        assert [st["ops_eq"], st["ops_gt"], st["ops_le"], st["ops_and"], st["ops_or"], st["ops_not"]] == \
               [by.get(k, 0) for k in ("eq", "gt", "le", "and", "or", "not")]


def test_config5_plan_shape():
    # reference-shaped plan (every variant evaluated): ~75k PBS in <= 16 dependent levels
    st = fb.plan_stats("/a+b?c/", 256, reference_shaped=True)
    assert 60000 < st["pbs"] < 80000 and st["levels"] <= 16 and st["max_level_width"] > 10000
    ref_counters = (st["variants"], st["ct_ops"], st["cache_hits"])
    # default plan: OR operands implied by another operand are absorbed (x | (x & y) = x); the reference's
    # bookkeeping counters are unaffected, the PBS count collapses from O(n^2) to O(n)
    st = fb.plan_stats("/a+b?c/", 256)
    assert st["pbs"] < 3000 and st["levels"] <= 8
    assert (st["variants"], st["ct_ops"], st["cache_hits"]) == ref_counters == (65025, 195583, 11118596)


def test_absorbed_and_reference_shaped_plans_agree():
    rnd = random.Random(11)
    for it in range(300):
        pat = rnd.choice(PATTERNS)
        max_n = 8 if "|" in pat and ("+" in pat or "*" in pat) else 20
        content = "".join(rnd.choice("abcxyAB.") for _ in range(rnd.randint(0, max_n)))
        a = fb.plan_eval_plain(pat, content, reference_shaped=True)
        assert fb.plan_eval_plain(pat, content) == a == rp.has_match(content, pat), (pat, content)


PATTERNS = ["/a+b?c/", "/ab{2,4}c/", r"/[a-d][^x-z]\./", "/abc/i", "/^a*b+$/", "/(ab|c)+x/", "/[^ab]+c/", "/a{,2}b/",
            "/[a-c]{2,}x?$/", "/^.a.$/", "/a|b|c/", "/(a|b)*c/", "/[abc][^a-b]c{2}/", "/x[ab]+y/i", "/aaaa+/", "/a.+b/",
            "/^abc$/", "/./", "/^.*$/", "/a?/", "/[^a-c]/", "/a{3}/", "/(a|b)?c$/"]


def test_lowering_matches_oracle_random():
    rnd = random.Random(7)
    for it in range(600):
        pat = rnd.choice(PATTERNS)
        # alternation under repetition enumerates 2^n variants (in the reference too): keep those short
        max_n = 8 if "|" in pat and ("+" in pat or "*" in pat) else 24
        content = "".join(rnd.choice("abcxyAB.") for _ in range(rnd.randint(0, max_n)))
        exp = rp.has_match(content, pat)
        assert fb.plan_eval_plain(pat, content) == exp, (pat, content)
        world = rnd.choice([2, 3, 8])
        parts = [fb.plan_eval_plain(pat, content, r, world) for r in range(world)]
        assert int(any(parts)) == exp, (pat, content, parts)


def test_long_content_absorbed_and_sharded_plans_match_oracle():
    """value-level absorption and the contiguous root slices at the sizes the bench uses"""
    rnd = random.Random(19)
    for pat, n in [("/a+b?c/", 64), ("/a+b?c/", 150), ("/ab{2,4}c/", 64), (r"/[a-d][^x-z]\./", 64), ("/x[ab]+y/i", 80), ("/aaaa+/", 90)]:
        for alphabet in ("abcx", "ab", "xyAB."):
            content = "".join(rnd.choice(alphabet) for _ in range(n))
            exp = rp.has_match(content, pat)
            assert fb.plan_eval_plain(pat, content) == exp, (pat, content)
            for world in (2, 5, 8):
                parts = [fb.plan_eval_plain(pat, content, r, world) for r in range(world)]
                assert int(any(parts)) == exp, (pat, content, world, parts)
    # the slices partition the work: per-rank leaf levels shrink with the world size
    w1 = fb.plan_level_widths("/a+b?c/", 256)
    w8 = [fb.plan_level_widths("/a+b?c/", 256, r, 8) for r in range(8)]
    assert all(w[0] <= w1[0] // 8 + 16 for w in w8), (w1, w8)


def test_lowering_config_rows():
    rnd = random.Random(3)
    c64 = "".join(rnd.choice("abcx") for _ in range(64))
    for pat in ["/a+b?c/", "/ab{2,4}c/", r"/[a-d][^x-z]\./", "/^a+b?c$/"]:
        assert fb.plan_eval_plain(pat, c64) == rp.has_match(c64, pat)
    assert fb.plan_eval_plain(r"/^[a-d][^x-z]\.$/", "aq.") == 0   # gt quirk
    assert fb.plan_eval_plain(r"/^[a-d][^x-z]\.$/", "bx.") == 1
    assert fb.plan_eval_plain("/^a{,2}$/", "aaa") == 1            # {,m} quirk
    assert fb.plan_eval_plain("/^$/", "") == 0


def test_client_glue_against_oracle(client_key):
    import numpy as np
    from oracle import tfhe
    ck = fb.ClientKey.load(os.path.join(GOLDEN, "client_key"))
    assert (ck.big == client_key.big).all() and (ck.small == client_key.small).all()
    with pytest.raises(fb.FbError):
        fb.ClientKey.from_bincode(open(os.path.join(GOLDEN, "client_key"), "rb").read()[:-8])
    # product encrypt -> oracle decrypt and vice versa
    cts = ck.encrypt_blocks(range(16), seed=4)
    assert [tfhe.decrypt_shortint(client_key, c) for c in cts] == list(range(16))
    assert [ck.decrypt_block(c) for c in tfhe.encrypt_batch(client_key, range(16))] == list(range(16))
    s = fb.encrypt_str(ck, "aZ~")
    assert [tfhe.decrypt_radix(client_key, s[i]) for i in range(3)] == [97, 90, 126]
    assert [ck.decrypt(tfhe.trivial_radix(v)) for v in (0, 1, 200)] == [0, 1, 200]
    assert (fb.trivial_str("az") == np.stack([tfhe.trivial_radix(97), tfhe.trivial_radix(122)])).all()
    with pytest.raises(ValueError):
        fb.encrypt_str(ck, "é")
    for f in (lambda x: x, lambda x: int(x == 3), lambda x: 15 - x):
        assert (fb.make_lut(f) == tfhe.make_lut(f)).all()


def test_library_reads_nothing_from_the_environment():
    """every knob is a per-context option (fb_set_option): no getenv anywhere in the product sources"""
    csrc = os.path.join(ROOT, "fhe_regex_b200", "csrc")
    for name in sorted(os.listdir(csrc)):
        if name.endswith((".cu", ".cpp", ".cuh", ".h")):
            assert "getenv" not in open(os.path.join(csrc, name)).read(), name


def test_product_does_not_touch_the_oracle():
    """oracle/ is test infrastructure: nothing under fhe_regex_b200/ or include/ imports, links or opens it"""
    for base in ("fhe_regex_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for name in files:
                if name.endswith((".py", ".cu", ".cpp", ".cuh", ".h")):
                    txt = open(os.path.join(dirpath, name)).read()
                    assert "from oracle" not in txt and "import oracle" not in txt and "libtfhe_oracle" not in txt and "oracle/" not in txt, name


def test_pinned_empty_is_an_ordinary_array_without_a_device():
    """fb_host_alloc / pinned_empty: page-locked when a device is present, a plain numpy array otherwise -- the same
    shape, dtype and contiguity either way, and encrypt_str (which allocates through it) still round-trips"""
    a = fb.pinned_empty((5, 4, fb.BIG), np.uint64)
    assert a.shape == (5, 4, fb.BIG) and a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    a[:] = 7
    assert int(a.sum()) == 7 * a.size
    assert fb.pinned_empty((0, 4, fb.BIG)).shape == (0, 4, fb.BIG)
    ck = fb.ClientKey.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "client_key"))
    ct = fb.encrypt_str(ck, "a+b", seed=5)
    assert ct.shape == (3, 4, fb.BIG) and [ck.decrypt(c) for c in ct] == [ord(ch) for ch in "a+b"]
