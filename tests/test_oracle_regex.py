"""Oracle (plaintext host logic) against the reference's own golden vectors."""
import json
import os

import pytest

from oracle import regex_plain as rp

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
PARSER_CASES = json.load(open(os.path.join(GOLDEN, "parser_cases.json")))["cases"]
ENGINE_CASES = json.load(open(os.path.join(GOLDEN, "engine_cases.json")))["cases"]


def test_golden_counts():
    # parser.rs:358-678 holds 49 cases, engine.rs:256-280 holds 25
    assert len(PARSER_CASES) == 49
    assert len(ENGINE_CASES) == 25


@pytest.mark.parametrize("case", PARSER_CASES, ids=[c["pattern"] for c in PARSER_CASES])
def test_parser_golden(case):
    assert rp.parse(case["pattern"]) == case["ast"]


@pytest.mark.parametrize("case", ENGINE_CASES, ids=["%s~%s" % (c["content"], c["pattern"]) for c in ENGINE_CASES])
def test_engine_golden(case):
    assert rp.has_match(case["content"], case["pattern"]) == case["expected"]


@pytest.mark.parametrize("pattern", ["abc", "/abc", "/abc/x", "/a(b/", "/[ab/", "/a**/", "/+/", "/a{1,2/"])
def test_parse_errors(pattern):
    with pytest.raises(rp.ParseError):
        rp.parse(pattern)


def test_reference_panics():
    # /a{}/ -> parse_digits("").unwrap() panics (parser.rs:349-351)
    with pytest.raises(rp.RefPanic):
        rp.parse("/a{}/")
    # Seq[] panics when evaluated (engine.rs:189-190); parse itself succeeds (parser.rs:545-556)
    assert rp.parse("/^/") == {"Seq": ["SOF", {"Seq": []}]}
    with pytest.raises(rp.RefPanic):
        rp.has_match("a", "/^/")


def test_quirks():
    # ct_ge is smart_gt (execution.rs:93): [a-d] rejects 'a', [^x-z] accepts 'x'
    assert rp.has_match("bq.", r"/^[a-d][^x-z]\.$/") == 1
    assert rp.has_match("aq.", r"/^[a-d][^x-z]\.$/") == 0
    assert rp.has_match("bx.", r"/^[a-d][^x-z]\.$/") == 1
    assert rp.has_match("by.", r"/^[a-d][^x-z]\.$/") == 0
    # {,m} allows m+1 repetitions (engine.rs:139-160)
    assert rp.has_match("aaa", "/^a{,2}$/") == 1
    assert rp.has_match("aaaa", "/^a{,2}$/") == 0
    # empty content never matches, even /^$/ (engine.rs:15,22-26)
    assert rp.has_match("", "/^$/") == 0
    # /i rewrites only Char (parser.rs:67)
    assert rp.has_match("B", "/[a-c]/i") == 0


# (content, pattern, variants, ct_ops, cache_hits): BASELINE.md section 2 / SURVEY.md 8d
COUNT_ROWS = [
    ("abc", "/^abc$/", 1, 5, 0),
    ("aBc", "/^abc$/i", 1, 11, 0),
    ("aBc" + "x" * 13, "/^abc$/i", 0, 0, 0),
    ("x" * 13 + "aBc", "/abc/i", 14, 167, 0),
    ("bq.", r"/^[a-d][^x-z]\.$/", 1, 10, 0),
    ("q" * 64, r"/^[a-d][^x-z]\.$/", 0, 0, 0),
    ("q" * 64, r"/[a-d][^x-z]\./", 62, 681, 0),
    ("abbc", "/^ab{2,4}c$/", 1, 7, 0),
    ("abbbc", "/^ab{2,4}c$/", 1, 9, 0),
    ("abbbbc", "/^ab{2,4}c$/", 1, 11, 0),
    ("q" * 64, "/ab{2,4}c/", 180, 903, 892),
    ("q" * 64, "/a+b?c/", 3969, 12031, 170500),
]


@pytest.mark.parametrize("row", COUNT_ROWS, ids=["%d:%s" % (len(r[0]), r[1]) for r in COUNT_ROWS])
def test_op_counts(row):
    content, pattern, variants, ops, hits = row
    _, ex, nb = rp.has_match(content, pattern, return_exec=True)
    assert (nb, ex.ct_ops, ex.cache_hits) == (variants, ops, hits)


def test_debug_fmt():
    assert rp.debug_fmt(rp.parse("/^ab?c$/")) == "<^<ab?c>$>"
    assert rp.debug_fmt(rp.parse("/[a-d]|[^xyz]{2,}/")) == "([a->d]|[^[xyz]]{2,*})"
