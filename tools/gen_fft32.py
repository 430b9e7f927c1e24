#!/usr/bin/env python3
"""Emit fhe_regex_b200/csrc/fft32_gen.h: straight-line 32-point complex FFTs that live entirely in the
registers of one thread (constant indices only, so nvcc never spills the arrays to local memory).

All three transforms are decimation-in-time butterflies whose twiddle multiplications are done with the
"tangent" form: w*b = c*[(br - t*bi) + i(bi + t*br)], t = s/c (2 FMA instead of 2 MUL + 2 FMA), the real
factor c being carried as a compile-time pending magnitude of b and folded into the constant of the
following a +- rho*b (an FMA in place of an ADD).  In a DIT network every output inherits the pending
magnitude of input position 0, which is 1, so nothing is left to normalise at the end.  A generic
butterfly costs 6 FP64 issue slots instead of 8, and the input twist of the negacyclic transform
(exp(i*pi*r/64) on register r) is folded in as one more pending rotation: 2 slots instead of 4.

  fft32_fwd_twist(xr, xi)  in : xr[r] = d0, xi[r] = d1 (integer digits of coefficients 32r+lane, +1024)
                           out: register q holds X[brev5(q)] of the forward DFT (W = exp(-2*pi*i/32)) of
                                (d0 + i*d1) * exp(i*pi*r/64)
  fft32_fwd(xr, xi)        natural in, register q holds X[brev5(q)] out, forward DFT
  fft32_inv(xr, xi)        register q holds Y[brev5(q)] in, natural out, inverse DFT (unnormalised)

The generator checks every emitted routine numerically (numpy float64 replay against a direct DFT).
"""
import math
import os
from fractions import Fraction

import numpy as np

LD = np.longdouble
LD_PI = LD("3.141592653589793238462643383279502884")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "fhe_regex_b200", "csrc", "fft32_gen.h")


def brev5(v):
    return int("{:05b}".format(v)[::-1], 2)


def cs_frac(fr):
    """(cos, sin) of 2*pi*fr in long double, exact on the axes and diagonals."""
    fr = fr % 1
    table = {Fraction(0): (1, 0), Fraction(1, 4): (0, 1), Fraction(1, 2): (-1, 0), Fraction(3, 4): (0, -1)}
    if fr in table:
        c, s = table[fr]
        return LD(c), LD(s)
    if (fr * 8).denominator == 1:
        r = LD(0.5) ** LD(0.5)
        e = int(fr * 8)
        return {1: (r, r), 3: (-r, r), 5: (-r, -r), 7: (r, -r)}[e]
    a = 2 * LD_PI * LD(fr.numerator) / LD(fr.denominator)
    return np.cos(a), np.sin(a)


def lit(x):
    return repr(float(x))


KTAB = []  # distinct |constants| of the butterflies; referenced as FBK(i) so that the device code reads them
           # straight out of the constant bank as DFMA operands (literals cost two UMOVs per use)


def kref(x):
    x = float(x)
    a = abs(x)
    if a not in KTAB:
        KTAB.append(a)
    r = "FBK(%d)" % KTAB.index(a)
    return "-" + r if x < 0 else r


class Emitter:
    """Tracks, per array slot, stored value -> true value = sign * mag * stored (sign per component)."""

    def __init__(self):
        self.lines = []
        self.mag = [LD(1)] * 32
        self.sr = [1] * 32
        self.si = [1] * 32
        self.slots = 0  # FP64 issue slots emitted

    def emit(self, s):
        self.lines.append(s)

    def rotate(self, p, c, s):
        """slot p *= (c + i s), |c+is| = 1, by the tangent form; updates pending magnitude / signs."""
        if abs(s) < 1e-30:  # +-1
            if c < 0:
                self.sr[p] = -self.sr[p]
                self.si[p] = -self.si[p]
            return
        if abs(c) < 1e-30:  # +-i : (re, im) -> (-+im, +-re)
            self.emit(f"{{ const double t_ = xr[{p}]; xr[{p}] = xi[{p}]; xi[{p}] = t_; }}")
            nsr = -self.si[p] * (1 if s > 0 else -1)
            nsi = self.sr[p] * (1 if s > 0 else -1)
            self.sr[p], self.si[p] = nsr, nsi
            return
        if abs(c) >= abs(s):
            # true_re = c*Sr*m*[Br - (s*Si/(c*Sr))*Bi] ; true_im = c*Si*m*[Bi + (s*Sr/(c*Si))*Br]
            kr = -(s * self.si[p]) / (c * self.sr[p])
            ki = (s * self.sr[p]) / (c * self.si[p])
            self.emit(f"{{ const double t_ = xr[{p}]; xr[{p}] = {self.fma('xi[%d]' % p, kr, 't_')}; xi[{p}] = {self.fma('t_', ki, 'xi[%d]' % p)}; }}")
            f = c
        else:
            # true_re = -s*Si*m*[Bi - (c*Sr/(s*Si))*Br] -> stored re' = Bi + k*Br with sign -sgn(s)*Si
            # true_im =  s*Sr*m*[Br + (c*Si/(s*Sr))*Bi]
            kr = -(c * self.sr[p]) / (s * self.si[p])
            ki = (c * self.si[p]) / (s * self.sr[p])
            self.emit(f"{{ const double t_ = xr[{p}]; xr[{p}] = {self.fma('t_', kr, 'xi[%d]' % p)}; xi[{p}] = {self.fma('xi[%d]' % p, ki, 't_')}; }}")
            nsr = -self.si[p] * (1 if s > 0 else -1)
            nsi = self.sr[p] * (1 if s > 0 else -1)
            self.sr[p], self.si[p] = nsr, nsi
            self.mag[p] = self.mag[p] * abs(s)
            self.slots += 2
            return
        if f < 0:
            self.sr[p] = -self.sr[p]
            self.si[p] = -self.si[p]
        self.mag[p] = self.mag[p] * abs(f)
        self.slots += 2

    def fma(self, x, k, y):
        """C expression for x*k + y with the +-1 cases as plain add/sub."""
        kf = float(k)
        if kf == 1.0:
            return f"{y} + {x}"
        if kf == -1.0:
            return f"{y} - {x}"
        return f"fb_fma({x}, {kref(kf)}, {y})"

    def butterfly(self, a, b):
        """(a, b) <- (a + b, a - b) in true values; results inherit a's pending magnitude and signs."""
        rr = (self.sr[b] * self.mag[b]) / (self.sr[a] * self.mag[a])
        ri = (self.si[b] * self.mag[b]) / (self.si[a] * self.mag[a])
        self.emit(f"{{ const double tr_ = xr[{b}], ti_ = xi[{b}]; "
                  f"xr[{b}] = {self.fma('tr_', -rr, 'xr[%d]' % a)}; xi[{b}] = {self.fma('ti_', -ri, 'xi[%d]' % a)}; "
                  f"xr[{a}] = {self.fma('tr_', rr, 'xr[%d]' % a)}; xi[{a}] = {self.fma('ti_', ri, 'xi[%d]' % a)}; }}")
        self.mag[b], self.sr[b], self.si[b] = self.mag[a], self.sr[a], self.si[a]
        self.slots += 4

    def finish(self):
        for p in range(32):
            assert abs(float(self.mag[p]) - 1.0) < 1e-15 and self.sr[p] == 1 and self.si[p] == 1, (p, self.mag[p], self.sr[p], self.si[p])

    def mark(self):
        """index of the next emitted line (pieces of a routine are slices of self.lines)"""
        return len(self.lines)


def gen_dit(sign, slot_of_pos, twist):
    """DIT network over positions 0..31 (position p holds input x[brev5(p)]); array slot = slot_of_pos(p).
    sign = -1 forward, +1 inverse.  twist: pre-rotation exp(i*pi*r/64) of natural input index r."""
    E = Emitter()
    if twist:
        for p in range(32):
            r = brev5(p)  # natural input index held at position p
            c, s = cs_frac(Fraction(r, 128))
            E.rotate(slot_of_pos(p), c, s)
    half = 1
    while half <= 16:
        for g in range(0, 32, 2 * half):
            for j in range(half):
                a, b = g + j, g + j + half
                c, s = cs_frac(Fraction(sign * j, 2 * half))
                E.rotate(slot_of_pos(b), c, s)
                E.butterfly(slot_of_pos(a), slot_of_pos(b))
        half *= 2
    E.finish()
    return E



def dit_stages(sign, slot_of_pos):
    """butterflies of the 32-point DIT network as {stage: [(a_slot, b_slot, c, s)]}, stage = 1..5"""
    st = {}
    half, stage = 1, 1
    while half <= 16:
        ops = []
        for g in range(0, 32, 2 * half):
            for j in range(half):
                c, s = cs_frac(Fraction(sign * j, 2 * half))
                ops.append((slot_of_pos(g + j), slot_of_pos(g + j + half), c, s))
        st[stage] = ops
        half *= 2
        stage += 1
    return st


def emit_ops(E, ops):
    for a, b, c, s in ops:
        E.rotate(b, c, s)
        E.butterfly(a, b)


def gen_mid_pieces():
    """Forward pass 2 and inverse pass 1 cut into pieces so that the Fourier MAC can run block by block between
    them: the last three forward stages and the first three inverse stages stay inside aligned blocks of 8 slots
    (the last / first two inside blocks of 4).  Returns ([(name, lines)], slots)."""
    pieces = []
    F = Emitter()
    fs = dit_stages(-1, brev5)
    m = F.mark()
    emit_ops(F, fs[1] + fs[2])
    pieces.append(("fft32_fwd_s12", F.lines[m:]))
    for b in range(4):
        m = F.mark()
        emit_ops(F, [o for o in fs[3] if o[0] // 8 == b])
        assert all(o[1] // 8 == b for o in fs[3] if o[0] // 8 == b)
        pieces.append((f"fft32_fwd_s3_b{b}", F.lines[m:]))
        for q in (2 * b, 2 * b + 1):
            m = F.mark()
            sel = [o for o in fs[4] + fs[5] if o[0] // 4 == q]
            assert all(o[1] // 4 == q for o in sel) and len(sel) == 4
            emit_ops(F, sel)
            pieces.append((f"fft32_fwd_s45_q{q}", F.lines[m:]))
    F.finish()
    I = Emitter()
    isg = dit_stages(+1, lambda p: p)
    for b in range(4):
        for q in (2 * b, 2 * b + 1):
            m = I.mark()
            sel = [o for o in isg[1] + isg[2] if o[0] // 4 == q]
            assert all(o[1] // 4 == q for o in sel) and len(sel) == 4
            emit_ops(I, sel)
            pieces.append((f"fft32_inv_s12_q{q}", I.lines[m:]))
        m = I.mark()
        sel = [o for o in isg[3] if o[0] // 8 == b]
        assert all(o[1] // 8 == b for o in sel)
        emit_ops(I, sel)
        pieces.append((f"fft32_inv_s3_b{b}", I.lines[m:]))
    m = I.mark()
    emit_ops(I, isg[4] + isg[5])
    pieces.append(("fft32_inv_s45", I.lines[m:]))
    I.finish()
    return pieces, F.slots, I.slots


F1_ORDER = [brev5(n) for n in range(32)]   # slot whose digits are produced n-th: 0, 16, 8, 24, 4, ...


def gen_f1_progressive():
    """Forward pass 1 (with the input twist) as 32 steps: step n runs everything that becomes computable once the
    digits of slot F1_ORDER[n] are there, so that the decomposition (integer / shared-memory work) and the
    butterflies (FP64) interleave in program order."""
    E = Emitter()
    fs = dit_stages(-1, brev5)
    level = [0] * 32      # stages completed per slot
    loaded = [False] * 32
    todo = {s: list(fs[s]) for s in fs}
    steps = []
    for n in range(32):
        r = F1_ORDER[n]
        m = E.mark()
        c, s = cs_frac(Fraction(r, 128))
        E.rotate(r, c, s)
        loaded[r] = True
        progress = True
        while progress:
            progress = False
            for st in range(1, 6):
                rest = []
                for op in todo[st]:
                    a, b = op[0], op[1]
                    if loaded[a] and loaded[b] and level[a] == st - 1 and level[b] == st - 1:
                        emit_ops(E, [op])
                        level[a] = level[b] = st
                        progress = True
                    else:
                        rest.append(op)
                todo[st] = rest
        steps.append(E.lines[m:])
    assert all(not todo[s] for s in todo)
    E.finish()
    return steps, E.slots


def gen_i2_final():
    """Inverse pass 2 with the output untwist exp(-i*pi*r/64) folded in: stages 1-4 in one piece, then the 16
    butterflies of the last stage one by one, each followed by the untwist of its two outputs in tangent form.
    The real factor of the untwist stays pending: slot r's true value is kre[r] * xr[r] + i * kim[r] * xi[r];
    phase C folds kre / kim (and the 1/1024 of the inverse transform) into the FMAs of its torus rounding."""
    E = Emitter()
    isg = dit_stages(+1, lambda p: p)
    m = E.mark()
    emit_ops(E, isg[1] + isg[2] + isg[3] + isg[4])
    head = E.lines[m:]
    fins = []
    for op in isg[5]:
        a, b = op[0], op[1]
        assert b == a + 16
        m = E.mark()
        emit_ops(E, [op])
        for r in (a, b):
            c, s = cs_frac(Fraction(-r, 128))
            E.rotate(r, c, s)
        fins.append((a, E.lines[m:]))
    kre = [float(E.sr[r] * E.mag[r]) for r in range(32)]
    kim = [float(E.si[r] * E.mag[r]) for r in range(32)]
    return head, fins, kre, kim, E.slots


def replay(lines, xr, xi):
    """execute the emitted C statements with numpy float64 semantics (fma -> separate mul/add: fine for a check)"""
    env = {"xr": xr, "xi": xi, "fb_fma": lambda a, b, c: a * b + c, "FBK": lambda i: KTAB[i]}
    for ln in lines:
        body = ln.strip()
        assert body.startswith("{") and body.endswith("}")
        for st in body[1:-1].split(";"):
            st = st.strip()
            if not st:
                continue
            if st.startswith("const double"):
                for decl in st[len("const double"):].split(","):
                    name, expr = decl.split("=")
                    env[name.strip()] = eval(expr, {}, env)
            else:
                lhs, expr = st.split("=", 1)
                val = eval(expr, {}, env)
                arr, idx = lhs.strip().split("[")
                env[arr][int(idx[:-1])] = val


def check(E, sign, slot_of_pos, twist, out_slot_of_k):
    rng = np.random.default_rng(1)
    x = rng.standard_normal(32) + 1j * rng.standard_normal(32)
    xr, xi = np.zeros(32), np.zeros(32)
    for p in range(32):
        n = brev5(p)
        xr[slot_of_pos(p)], xi[slot_of_pos(p)] = x[n].real, x[n].imag
    replay(E.lines, xr, xi)
    xin = x * np.exp(1j * np.pi * np.arange(32) / 64) if twist else x
    ref = np.array([sum(xin[n] * np.exp(sign * 2j * np.pi * n * k / 32) for n in range(32)) for k in range(32)])
    got = np.array([xr[out_slot_of_k(k)] + 1j * xi[out_slot_of_k(k)] for k in range(32)])
    err = np.abs(got - ref).max()
    assert err < 1e-12, err
    return err


def main():
    out = ["// GENERATED by tools/gen_fft32.py -- do not edit.", "#pragma once", ""]
    out.append("// twist exp(i*pi*r/64) for register r (coefficient index 32*r + lane); switch on a constant folds to an immediate")
    for name, fn in (("cos", np.cos), ("sin", np.sin)):
        out.append(f"FB_HD constexpr double fb_twist_{name}(int r) {{")
        out.append("  switch (r) {")
        for r in range(32):
            out.append(f"    case {r}: return {lit(float(fn(LD_PI * LD(r) / LD(64))))};")
        out.append("  }")
        out.append("  return 0.0;")
        out.append("}")
    out.append("")
    # forward: natural input index r sits in array slot r; position p holds x[brev5(p)] -> slot brev5(p);
    # output X[k] ends at position k = slot brev5(k): "register q holds X[brev5(q)]"
    body = []
    fwd_slot = brev5
    # inverse: slot q holds Y[brev5(q)] = the input of position q; output natural: y[n] at slot n
    inv_slot = lambda p: p
    specs = [("fft32_fwd_twist", -1, fwd_slot, True, brev5), ("fft32_fwd", -1, fwd_slot, False, brev5), ("fft32_inv", +1, inv_slot, False, lambda k: k)]
    for name, sign, slot, twist, outslot in specs:
        E = gen_dit(sign, slot, twist)
        err = check(E, sign, slot, twist, outslot)
        print(f"{name}: {E.slots} FP64 slots, replay max err {err:.2e}")
        body.append(f"// {E.slots} FP64 issue slots")
        body.append(f"FB_HD void {name}(double (&xr)[32], double (&xi)[32]) {{")
        body += ["  " + l for l in E.lines]
        body.append("}")
        body.append("")
    # ---- pieces for the fused CMUX body (kernels.cu, variant "fused") ---------------------------------------
    def fn(name, lines, comment=None):
        if comment:
            body.append("// " + comment)
        body.append(f"FB_HD void {name}(double (&xr)[32], double (&xi)[32]) {{")
        body.extend("  " + l for l in lines)
        body.append("}")

    def dft(x, sign):
        return np.array([sum(x[n] * np.exp(sign * 2j * np.pi * n * k / 32) for n in range(32)) for k in range(32)])

    rng = np.random.default_rng(2)
    pieces, fslots, islots = gen_mid_pieces()
    # replay in the kernel's order: s12, then per block of 8 (s3, s45 x2 | MAC omitted | inv s12 x2, inv s3), inv s45
    x = rng.standard_normal(32) + 1j * rng.standard_normal(32)
    xr, xi = x.real.copy(), x.imag.copy()
    pd = dict(pieces)
    order_f = ["fft32_fwd_s12"] + [n for b in range(4) for n in (f"fft32_fwd_s3_b{b}", f"fft32_fwd_s45_q{2*b}", f"fft32_fwd_s45_q{2*b+1}")]
    for n in order_f:
        replay(pd[n], xr, xi)
    ref = dft(x, -1)
    errf = max(abs((xr[brev5(k)] + 1j * xi[brev5(k)]) - ref[k]) for k in range(32))
    y = rng.standard_normal(32) + 1j * rng.standard_normal(32)
    yr, yi = np.zeros(32), np.zeros(32)
    for q in range(32):
        yr[q], yi[q] = y[brev5(q)].real, y[brev5(q)].imag
    order_i = [n for b in range(4) for n in (f"fft32_inv_s12_q{2*b}", f"fft32_inv_s12_q{2*b+1}", f"fft32_inv_s3_b{b}")] + ["fft32_inv_s45"]
    for n in order_i:
        replay(pd[n], yr, yi)
    refi = dft(y, +1)
    erri = max(abs((yr[k] + 1j * yi[k]) - refi[k]) for k in range(32))
    assert errf < 1e-12 and erri < 1e-12, (errf, erri)
    print(f"mid pieces: fwd {fslots} + inv {islots} FP64 slots, replay max err {errf:.2e} / {erri:.2e}")
    body.append("// ---- forward pass 2 / inverse pass 1 in pieces: the Fourier MAC runs block by block between them")
    for name, lines in pieces:
        fn(name, lines)
    for stem, count in (("fft32_fwd_s3_b", 4), ("fft32_fwd_s45_q", 8), ("fft32_inv_s12_q", 8), ("fft32_inv_s3_b", 4)):
        body.append(f"template <int I> FB_HD void {stem[:-2]}(double (&xr)[32], double (&xi)[32]) {{")
        for i in range(count):
            body.append(f"  if constexpr (I == {i}) {stem}{i}(xr, xi);")
        body.append("}")
    body.append("")

    steps, f1slots = gen_f1_progressive()
    x = rng.standard_normal(32) + 1j * rng.standard_normal(32)
    xr, xi = np.zeros(32), np.zeros(32)
    for n in range(32):
        r = F1_ORDER[n]
        xr[r], xi[r] = x[r].real, x[r].imag      # "digits of slot r arrive"
        replay(steps[n], xr, xi)
    ref = dft(x * np.exp(1j * np.pi * np.arange(32) / 64), -1)
    err1 = max(abs((xr[brev5(k)] + 1j * xi[brev5(k)]) - ref[k]) for k in range(32))
    assert err1 < 1e-12, err1
    print(f"f1 progressive: {f1slots} FP64 slots, replay max err {err1:.2e}")
    body.append("// ---- forward pass 1 with the input twist, progressive: step n runs what becomes computable once slot")
    body.append("// fb_f1_order(n) holds its digits (decomposition and butterflies interleave in program order)")
    body.append("FB_HD constexpr int fb_f1_order(int n) {")
    body.append("  switch (n) {")
    for n in range(32):
        body.append(f"    case {n}: return {F1_ORDER[n]};")
    body.append("  }")
    body.append("  return 0;")
    body.append("}")
    for n in range(32):
        fn(f"fft32_f1_step{n}", steps[n])
    body.append("template <int N> FB_HD void fft32_f1_step(double (&xr)[32], double (&xi)[32]) {")
    for n in range(32):
        body.append(f"  if constexpr (N == {n}) fft32_f1_step{n}(xr, xi);")
    body.append("}")
    body.append("")

    head, fins, kre, kim, i2slots = gen_i2_final()
    y = rng.standard_normal(32) + 1j * rng.standard_normal(32)
    yr, yi = np.zeros(32), np.zeros(32)
    for q in range(32):
        yr[q], yi[q] = y[brev5(q)].real, y[brev5(q)].imag
    replay(head, yr, yi)
    for a, lines in fins:
        replay(lines, yr, yi)
    refi = dft(y, +1) * np.exp(-1j * np.pi * np.arange(32) / 64)
    err2 = max(abs((kre[k] * yr[k] + 1j * kim[k] * yi[k]) - refi[k]) for k in range(32))
    assert err2 < 1e-12, err2
    print(f"i2 final: {i2slots} FP64 slots (untwist folded), replay max err {err2:.2e}")
    body.append("// ---- inverse pass 2 with the output untwist folded in: head = stages 1-4, then the last stage butterfly by")
    body.append("// butterfly (slots a, a+16); true value of slot r = fb_i2_kre(r) * xr[r] + i * fb_i2_kim(r) * xi[r]")
    fn("fft32_i2_head", head)
    for a, lines in fins:
        fn(f"fft32_i2_fin{a}", lines)
    body.append("template <int A> FB_HD void fft32_i2_fin(double (&xr)[32], double (&xi)[32]) {")
    for a, _ in fins:
        body.append(f"  if constexpr (A == {a}) fft32_i2_fin{a}(xr, xi);")
    body.append("}")
    for nm, tab in (("kre", kre), ("kim", kim)):
        body.append(f"FB_HD constexpr double fb_i2_{nm}(int r) {{")
        body.append("  switch (r) {")
        for r in range(32):
            body.append(f"    case {r}: return {lit(tab[r])};")
        body.append("  }")
        body.append("  return 0.0;")
        body.append("}")
    body.append("")

    out.append("// butterfly constants (tangents and magnitude ratios); device code reads them from the constant bank")
    out.append("#define FB_KTAB_INIT { " + ", ".join(lit(k) for k in KTAB) + " }")
    out.append("#if defined(__CUDACC__)")
    out.append(f"static __constant__ double fb_ktab_d[{len(KTAB)}] = FB_KTAB_INIT;")
    out.append("#endif")
    out.append(f"static const double fb_ktab_h[{len(KTAB)}] = FB_KTAB_INIT;")
    out.append("#if defined(__CUDA_ARCH__)")
    out.append("#define FBK(i) fb_ktab_d[i]")
    out.append("#else")
    out.append("#define FBK(i) fb_ktab_h[i]")
    out.append("#endif")
    out.append("")
    out += body
    print(len(KTAB), "distinct constants")
    with open(OUT, "w") as f:
        f.write("\n".join(out) + "\n")
    print("wrote", os.path.normpath(OUT), len(out), "lines")


if __name__ == "__main__":
    main()
