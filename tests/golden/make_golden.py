#!/usr/bin/env python3
"""Generate tests/golden/{parser_cases,engine_cases}.json from the reference's own unit tests.

Reads (never copies) /root/reference/src/regex/parser.rs (49 `#[test_case]` at :358-678) and
/root/reference/src/regex/engine.rs (25 `#[test_case]` at :256-280) and converts the Rust literals
into JSON.  Run in the build container only (the GPU box has no /root/reference); the JSON is
committed.

AST JSON encoding (mirrors RegExpr, parser.rs:9-41):
  "SOF" | "EOF" | "AnyChar"
  {"Char": c}  {"Between": [from, to]}  {"Range": [c, ...]}  {"Not": ast}  {"Either": [l, r]}
  {"Optional": ast}  {"Repeated": [ast, at_least|null, at_most|null]}  {"Seq": [ast, ...]}
with characters as integer byte values.
"""
import json
import os
import re
import sys

REF = "/root/reference/src/regex"
HERE = os.path.dirname(os.path.abspath(__file__))


def extract_test_cases(src: str):
    """yield the raw argument text of every #[test_case( ... )] (balanced parens, string aware)."""
    i = 0
    while True:
        i = src.find("#[test_case(", i)
        if i < 0:
            return
        j = i + len("#[test_case(")
        depth, k = 1, j
        while depth:
            ch = src[k]
            if ch == '"':
                k += 1
                while src[k] != '"':
                    k += 2 if src[k] == "\\" else 1
            elif ch == "'" and src[k - 1] == "b":
                k += 1
                while src[k] != "'":
                    k += 2 if src[k] == "\\" else 1
            elif ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            k += 1
        yield src[j:k - 1]
        i = k


class Tok:
    def __init__(self, s):
        self.s, self.i = s, 0

    def ws(self):
        while self.i < len(self.s) and self.s[self.i].isspace():
            self.i += 1

    def peek(self, lit):
        self.ws()
        return self.s.startswith(lit, self.i)

    def eat(self, lit):
        self.ws()
        assert self.s.startswith(lit, self.i), (lit, self.s[self.i:self.i + 40])
        self.i += len(lit)

    def try_eat(self, lit):
        if self.peek(lit):
            self.i += len(lit)
            return True
        return False

    def string(self):
        self.ws()
        assert self.s[self.i] == '"'
        self.i += 1
        out = []
        while self.s[self.i] != '"':
            ch = self.s[self.i]
            if ch == "\\":
                nx = self.s[self.i + 1]
                out.append({"\\": "\\", '"': '"', "n": "\n", "'": "'"}[nx])
                self.i += 2
            else:
                out.append(ch)
                self.i += 1
        self.i += 1
        return "".join(out)

    def byte(self):
        self.eat("b'")
        ch = self.s[self.i]
        if ch == "\\":
            ch = {"\\": "\\", "'": "'", '"': '"'}[self.s[self.i + 1]]
            self.i += 2
        else:
            self.i += 1
        self.eat("'")
        return ord(ch)

    def ident(self):
        self.ws()
        m = re.match(r"[A-Za-z_][A-Za-z_0-9]*", self.s[self.i:])
        assert m, self.s[self.i:self.i + 40]
        self.i += m.end()
        return m.group(0)

    def number(self):
        self.ws()
        m = re.match(r"[0-9]+", self.s[self.i:])
        self.i += m.end()
        return int(m.group(0))


def parse_opt(t: Tok):
    if t.try_eat("None"):
        return None
    t.eat("Some")
    t.eat("(")
    n = t.number()
    t.eat(")")
    return n


def parse_boxed(t: Tok):
    t.eat("Box::new(")
    a = parse_ast(t)
    t.eat(")")
    return a


def parse_fields(t: Tok, spec):
    """parse `{ name: value, ... }` with per-field value parsers, any order, optional trailing comma."""
    out = {}
    t.eat("{")
    while not t.peek("}"):
        name = t.ident()
        t.eat(":")
        out[name] = spec[name](t)
        t.try_eat(",")
    t.eat("}")
    return out


def parse_vec(t: Tok, elem):
    t.eat("vec![")
    out = []
    while not t.peek("]"):
        out.append(elem(t))
        t.try_eat(",")
    t.eat("]")
    return out


def parse_ast(t: Tok):
    t.eat("RegExpr::")
    kind = t.ident()
    if kind in ("SOF", "EOF", "AnyChar"):
        return kind
    if kind == "Char":
        return {"Char": parse_fields(t, {"c": Tok.byte})["c"]}
    if kind == "Between":
        f = parse_fields(t, {"from": Tok.byte, "to": Tok.byte})
        return {"Between": [f["from"], f["to"]]}
    if kind == "Range":
        return {"Range": parse_fields(t, {"cs": lambda tt: parse_vec(tt, Tok.byte)})["cs"]}
    if kind == "Not":
        return {"Not": parse_fields(t, {"not_re": parse_boxed})["not_re"]}
    if kind == "Either":
        f = parse_fields(t, {"l_re": parse_boxed, "r_re": parse_boxed})
        return {"Either": [f["l_re"], f["r_re"]]}
    if kind == "Optional":
        return {"Optional": parse_fields(t, {"opt_re": parse_boxed})["opt_re"]}
    if kind == "Repeated":
        f = parse_fields(t, {"repeat_re": parse_boxed, "at_least": parse_opt, "at_most": parse_opt})
        return {"Repeated": [f["repeat_re"], f["at_least"], f["at_most"]]}
    if kind == "Seq":
        return {"Seq": parse_fields(t, {"re_xs": lambda tt: parse_vec(tt, parse_ast)})["re_xs"]}
    raise AssertionError(kind)


def main():
    parser_src = open(os.path.join(REF, "parser.rs")).read()
    cases = []
    for raw in extract_test_cases(parser_src):
        t = Tok(raw)
        pattern = t.string()
        t.eat(",")
        ast = parse_ast(t)
        name = None
        if t.try_eat(";"):
            name = t.string()
        cases.append({"pattern": pattern, "ast": ast, "name": name})
    with open(os.path.join(HERE, "parser_cases.json"), "w") as f:
        json.dump({"source": "reference src/regex/parser.rs:358-678 (#[test_case] of test_parser)", "cases": cases}, f, indent=1)

    engine_src = open(os.path.join(REF, "engine.rs")).read()
    ecases = []
    for raw in extract_test_cases(engine_src):
        t = Tok(raw)
        content = t.string()
        t.eat(",")
        pattern = t.string()
        t.eat(",")
        exp = t.number()
        name = None
        if t.try_eat(";"):
            name = t.string()
        ecases.append({"content": content, "pattern": pattern, "expected": exp, "name": name})
    with open(os.path.join(HERE, "engine_cases.json"), "w") as f:
        json.dump({"source": "reference src/regex/engine.rs:256-280 (#[test_case] of test_has_match)", "cases": ecases}, f, indent=1)
    print(len(cases), "parser cases,", len(ecases), "engine cases")
    return 0


if __name__ == "__main__":
    sys.exit(main())
