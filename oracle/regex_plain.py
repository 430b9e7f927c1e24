"""Plaintext restatement of the reference's host logic: parser, variant generator, memoising executor.

TEST INFRASTRUCTURE ONLY (see oracle/tfhe_oracle.c header).  Pinned against the reference's own
golden vectors: 49 parser cases (parser.rs:358-678) and 25 engine cases (engine.rs:256-280), via
tests/golden/*.json.

Follows, function by function:
  parse            parser.rs:146-184   (grammar parser.rs:208-351, case folding parser.rs:43-81)
  build_branches   engine.rs:45-214
  has_match        engine.rs:8-42
  Execution        execution.rs:37-223 (cache keyed by structural provenance, constant short-circuits)

AST encoding = the JSON encoding of tests/golden/make_golden.py:
  "SOF" | "EOF" | "AnyChar" | {"Char": c} | {"Between": [f, t]} | {"Range": [..]} | {"Not": a}
  | {"Either": [l, r]} | {"Optional": a} | {"Repeated": [a, lo|None, hi|None]} | {"Seq": [..]}

The reference's quirks are reproduced on purpose: ct_ge is a strict '>'
(execution.rs:93), {,m} allows m+1 repetitions (engine.rs:139-160), Seq[] panics (engine.rs:189-190),
empty content never matches (engine.rs:15,22-26).
"""
from __future__ import annotations

NON_ESCAPABLE = b"&;:,`~-_!@#%'\""  # parser.rs:238-240


class ParseError(Exception):
    """anyhow::Error of parse (parser.rs:146-184)."""


class RefPanic(Exception):
    """A Rust panic in the reference (index out of bounds, unwrap on a failed parse, ...)."""


def _is_letter(b: int) -> bool:
    return (65 <= b <= 90) or (97 <= b <= 122)


def _is_digit(b: int) -> bool:
    return 48 <= b <= 57


class _P:
    """Backtracking recursive descent over bytes; each method returns (ast, new_pos) or None."""

    def __init__(self, s: bytes):
        self.s = s

    def byte(self, i, b):
        return i + 1 if i < len(self.s) and self.s[i] == b else None

    # regex := term '|' regex | term                                   parser.rs:208-222
    def regex(self, i):
        t = self.term(i)
        if t is not None:
            l_re, j = t
            k = self.byte(j, ord("|"))
            if k is not None:
                r = self.regex(k)
                if r is not None:
                    return {"Either": [l_re, r[0]]}, r[1]
        return self.term(i)

    # term := factor*                                                   parser.rs:224-236
    def term(self, i):
        xs = []
        while True:
            f = self.factor(i)
            if f is None:
                break
            xs.append(f[0])
            i = f[1]
        if len(xs) == 1:
            return xs[0], i
        return {"Seq": xs}, i

    # factor := atom '?' | repeated | atom                              parser.rs:243-256
    def factor(self, i):
        a = self.atom(i)
        if a is not None:
            k = self.byte(a[1], ord("?"))
            if k is not None:
                return {"Optional": a[0]}, k
        r = self.repeated(i)
        if r is not None:
            return r
        return self.atom(i)

    # atom                                                              parser.rs:262-277
    def atom(self, i):
        s = self.s
        if i >= len(s):
            return None
        b = s[i]
        if b == ord("."):
            return "AnyChar", i + 1
        if b == ord("\\") and i + 1 < len(s):
            return {"Char": s[i + 1]}, i + 2
        if _is_letter(b) or b in NON_ESCAPABLE:
            return {"Char": b}, i + 1
        if b == ord("["):
            r = self.range_(i + 1)
            if r is None:
                return None
            k = self.byte(r[1], ord("]"))
            return None if k is None else (r[0], k)
        if b == ord("("):
            r = self.regex(i + 1)
            if r is None:
                return None
            k = self.byte(r[1], ord(")"))
            return None if k is None else (r[0], k)
        return None

    # range := '^' range | letter '-' letter | letter+                  parser.rs:279-299
    def range_(self, i):
        s = self.s
        if i < len(s) and s[i] == ord("^"):
            r = self.range_(i + 1)
            return None if r is None else ({"Not": r[0]}, r[1])
        if i + 2 < len(s) and _is_letter(s[i]) and s[i + 1] == ord("-") and _is_letter(s[i + 2]):
            return {"Between": [s[i], s[i + 2]]}, i + 3
        j = i
        while j < len(s) and _is_letter(s[j]):
            j += 1
        if j == i:
            return None
        return {"Range": list(s[i:j])}, j

    def digits(self, i):
        j = i
        while j < len(self.s) and _is_digit(self.s[j]):
            j += 1
        return self.s[i:j], j

    # repeated                                                          parser.rs:301-347
    def repeated(self, i):
        a = self.atom(i)
        if a is None:
            return None
        re_, j = a
        s = self.s
        if j < len(s) and s[j] in (ord("*"), ord("+")):
            return {"Repeated": [re_, None if s[j] == ord("*") else 1, None]}, j + 1
        k = self.byte(j, ord("{"))
        if k is None:
            return None
        d1, k1 = self.digits(k)
        e = self.byte(k1, ord("}"))
        if e is not None:
            n = _parse_digits(d1)
            return {"Repeated": [re_, n, n]}, e
        c = self.byte(k1, ord(","))
        if c is None:
            return None
        d2, k2 = self.digits(c)
        e = self.byte(k2, ord("}"))
        if e is None:
            return None
        lo = None if len(d1) == 0 else _parse_digits(d1)
        hi = None if len(d2) == 0 else _parse_digits(d2)
        return {"Repeated": [re_, lo, hi]}, e


def _parse_digits(d: bytes) -> int:
    # parser.rs:349-351: unwrap() of str::parse::<usize> -> panics on "" (e.g. /a{}/)
    if len(d) == 0:
        raise RefPanic("parse_digits on empty string")
    return int(d)


def _case_insensitive(ast):
    # parser.rs:43-81: only Char is rewritten; Between/Range/others are left untouched
    if isinstance(ast, dict):
        (k, v), = ast.items()
        if k == "Char":
            if 97 <= v <= 122:
                return {"Range": [v, v - 32]}
            if 65 <= v <= 90:
                return {"Range": [v, v + 32]}
            return {"Range": [v]}
        if k == "Not":
            return {"Not": _case_insensitive(v)}
        if k == "Either":
            return {"Either": [_case_insensitive(v[0]), _case_insensitive(v[1])]}
        if k == "Optional":
            return {"Optional": _case_insensitive(v)}
        if k == "Repeated":
            return {"Repeated": [_case_insensitive(v[0]), v[1], v[2]]}
        if k == "Seq":
            return {"Seq": [_case_insensitive(x) for x in v]}
    return ast


def parse(pattern) -> object:
    """parser.rs:146-184."""
    s = pattern.encode("latin-1") if isinstance(pattern, str) else bytes(pattern)
    p = _P(s)
    i = p.byte(0, ord("/"))
    if i is None:
        raise ParseError("expected '/'")
    sof = p.byte(i, ord("^"))
    if sof is not None:
        i = sof
    r = p.regex(i)
    if r is None:
        raise ParseError("failed to parse regular expression")
    re_, i = r
    eof = p.byte(i, ord("$"))
    if eof is not None:
        i = eof
    j = p.byte(i, ord("/"))
    if j is None:
        raise ParseError("expected closing '/'")
    i = j
    if sof is not None or eof is not None:
        xs = []
        if sof is not None:
            xs.append("SOF")
        xs.append(re_)
        if eof is not None:
            xs.append("EOF")
        re_ = {"Seq": xs}
    ci = p.byte(i, ord("i"))
    if ci is not None:
        i = ci
        re_ = _case_insensitive(re_)
    if i != len(s):
        raise ParseError("failed to parse regular expression, unexpected token at start of: %r" % s[i:])
    return re_


def debug_fmt(ast) -> str:
    """impl Debug for RegExpr (parser.rs:87-144)."""
    if ast == "SOF":
        return "^"
    if ast == "EOF":
        return "$"
    if ast == "AnyChar":
        return "."
    (k, v), = ast.items()
    if k == "Char":
        return chr(v)
    if k == "Not":
        return "[^" + debug_fmt(v) + "]"
    if k == "Between":
        return "[%s->%s]" % (chr(v[0]), chr(v[1]))
    if k == "Range":
        return "[" + "".join(chr(c) for c in v) + "]"
    if k == "Either":
        return "(" + debug_fmt(v[0]) + "|" + debug_fmt(v[1]) + ")"
    if k == "Repeated":
        f = lambda n: "*" if n is None else str(n)
        return debug_fmt(v[0]) + "{" + f(v[1]) + "," + f(v[2]) + "}"
    if k == "Optional":
        return debug_fmt(v) + "?"
    if k == "Seq":
        return "<" + "".join(debug_fmt(x) for x in v) + ">"
    raise AssertionError(k)


# ---------------------------------------------------------------------------------------------
# Execution (execution.rs): plaintext values, structural cache keys interned to ints
# ---------------------------------------------------------------------------------------------
CT_FALSE, CT_TRUE = 0, 1


class Execution:
    """execution.rs:37-223 on plaintext bytes.  A result is (value, key) like ExecutedResult; keys
    are interned structural tuples so equality/hash cost is O(1) but cache behaviour (ct_ops,
    cache_hits) is identical to the reference's HashMap<Executed, RadixCiphertext>."""

    def __init__(self):
        self.cache = {}
        self.ct_ops = 0
        self.cache_hits = 0
        self.ops_by_type = {}
        self._intern = {}

    def _key(self, *t):
        k = self._intern.get(t)
        if k is None:
            k = len(self._intern)
            self._intern[t] = k
        return k

    def _const_of(self, key):
        return self._consts.get(key) if hasattr(self, "_consts") else None

    def ct_constant(self, c):
        k = self._key("C", c)
        if not hasattr(self, "_consts"):
            self._consts = {}
        self._consts[k] = c
        return (c, k)

    def ct_true(self):
        return self.ct_constant(CT_TRUE)

    def ct_false(self):
        return self.ct_constant(CT_FALSE)

    def ct_pos(self, at, value):
        return (value, self._key("P", at))

    def _with_cache(self, key, name, f):
        if key in self.cache:
            self.cache_hits += 1
            return (self.cache[key], key)
        self.ct_ops += 1
        self.ops_by_type[name] = self.ops_by_type.get(name, 0) + 1
        v = f()
        self.cache[key] = v
        return (v, key)

    def ct_eq(self, a, b):
        return self._with_cache(self._key("eq", a[1], b[1]), "eq", lambda: int(a[0] == b[0]))

    def ct_ge(self, a, b):  # execution.rs:93 calls smart_gt: strict
        return self._with_cache(self._key("ge", a[1], b[1]), "gt", lambda: int(a[0] > b[0]))

    def ct_le(self, a, b):
        return self._with_cache(self._key("le", a[1], b[1]), "le", lambda: int(a[0] <= b[0]))

    def ct_and(self, a, b):
        key = self._key("and", a[1], b[1])
        ca, cb = self._const_of(a[1]), self._const_of(b[1])
        if ca == CT_TRUE:
            return (b[0], key)
        if ca == CT_FALSE:
            return (a[0], key)
        if cb == CT_TRUE:
            return (a[0], key)
        if cb == CT_FALSE:
            return (b[0], key)
        return self._with_cache(key, "and", lambda: a[0] & b[0])

    def ct_or(self, a, b):
        key = self._key("or", a[1], b[1])
        ca, cb = self._const_of(a[1]), self._const_of(b[1])
        if ca == CT_TRUE:
            return (a[0], key)
        if cb == CT_TRUE:
            return (b[0], key)
        if ca == CT_FALSE and cb == CT_FALSE:
            return (a[0], key)
        return self._with_cache(key, "or", lambda: a[0] | b[0])

    def ct_not(self, a):
        return self._with_cache(self._key("not", a[1]), "not", lambda: a[0] ^ 1)


def build_branches(content: bytes, re_, c_pos: int):
    """engine.rs:45-214.  Returns [(thunk(exec) -> result, next_pos)]."""
    n = len(content)
    if re_ == "SOF":
        return [(lambda ex: ex.ct_true(), c_pos)] if c_pos == 0 else []
    if re_ == "EOF":
        return [(lambda ex: ex.ct_true(), c_pos)] if c_pos == n else []
    if c_pos >= n:
        return []
    if re_ == "AnyChar":
        return [(lambda ex: ex.ct_true(), c_pos + 1)]
    (k, v), = re_.items()
    if k == "Char":
        return [(lambda ex: ex.ct_eq(ex.ct_pos(c_pos, content[c_pos]), ex.ct_constant(v)), c_pos + 1)]
    if k == "Not":
        def wrap(branch):
            return lambda ex: ex.ct_not(branch(ex))
        return [(wrap(b), p) for b, p in build_branches(content, v, c_pos)]
    if k == "Either":
        return build_branches(content, v[0], c_pos) + build_branches(content, v[1], c_pos)
    if k == "Between":
        def between(ex):
            ch = ex.ct_pos(c_pos, content[c_pos])
            ct_from = ex.ct_constant(v[0])
            ct_to = ex.ct_constant(v[1])
            ge_from = ex.ct_ge(ch, ct_from)
            le_to = ex.ct_le(ch, ct_to)
            return ex.ct_and(ge_from, le_to)
        return [(between, c_pos + 1)]
    if k == "Range":
        def rng(ex):
            ch = ex.ct_pos(c_pos, content[c_pos])
            res = ex.ct_eq(ch, ex.ct_constant(v[0]))
            for c in v[1:]:
                e = ex.ct_eq(ch, ex.ct_constant(c))
                res = ex.ct_or(res, e)
            return res
        return [(rng, c_pos + 1)]
    if k == "Repeated":
        sub, lo, hi = v
        at_least = 0 if lo is None else lo
        at_most = (n - c_pos) if hi is None else hi
        if at_least > at_most:
            return []
        res = [
            [(lambda ex: ex.ct_true(), c_pos)] if at_least == 0 else [],
            build_branches(content, {"Seq": [sub] * max(1, at_least)}, c_pos),
        ]
        for _ in range(at_least + 1, at_most + 1):
            nxt = []
            for bp, bpos in res[-1]:
                for bx, bxpos in build_branches(content, sub, bpos):
                    nxt.append((_and_then(bp, bx), bxpos))
            res.append(nxt)
        return [x for lst in res for x in lst]
    if k == "Optional":
        return build_branches(content, v, c_pos) + [(lambda ex: ex.ct_true(), c_pos)]
    if k == "Seq":
        if len(v) == 0:
            raise RefPanic("Seq{re_xs: []}: index out of bounds (engine.rs:189-190)")
        conts = build_branches(content, v[0], c_pos)
        for re_x in v[1:]:
            nxt = []
            for bp, bpos in conts:
                for bx, bxpos in build_branches(content, re_x, bpos):
                    nxt.append((_and_then(bp, bx), bxpos))
            conts = nxt
        return conts
    raise RefPanic("unmatched regex variant")


def _and_then(branch_prev, branch_x):
    def f(ex):
        res_prev = branch_prev(ex)
        res_x = branch_x(ex)
        return ex.ct_and(res_prev, res_x)
    return f


def has_match(content, pattern, return_exec: bool = False):
    """engine.rs:8-42 on plaintext: returns the value the reference's decrypt would yield (0/1)."""
    if isinstance(content, str):
        content = content.encode("latin-1")
    re_ = parse(pattern)
    branches = []
    for i in range(len(content)):
        branches.extend(b for b, _ in build_branches(content, re_, i))
    ex = Execution()
    if len(branches) <= 1:
        res = branches[0](ex) if branches else ex.ct_false()
    else:
        res = branches[0](ex)
        for b in branches[1:]:
            br = b(ex)
            res = ex.ct_or(res, br)
    if return_exec:
        return res[0], ex, len(branches)
    return res[0]
