// br_w16.cuh -- per-thread building blocks of the radix-16 LATENCY blind rotation (K2-K4): one sample on
// 128 threads, 64 per polynomial, 16 complex points per thread.  Same arithmetic as br_wide.cuh (32-bit torus accumulator,
// f64 negacyclic FFT, the same Fourier key in natural frequency order), fewer and fatter stages:
//
//   br_wide.cuh   1024 = 8 x 8 x 8 x 2:  7 barrier-separated stages per CMUX step, 3 exchanges through shared memory per
//                 transform -> 4 170 shared-memory wavefronts per step, which is what bounds that kernel (two samples on one
//                 SM take 1.9 x the time of one: profiles/r02_pair_kernel_probe.json);
//   here          1024 = 16 x 16 x 4:    5 stages, 2 exchanges per transform -> ~3 100 wavefronts per step.
//
// Index algebra (W_n = exp(-2 pi i / n), w = exp(i pi / 2048) the negacyclic twist, z[j] = (c[j] + i c[j+1024]) w^j):
//   j = u + 64 m          u < 64 (thread of forward stage 1), m < 16 (register)
//   u = 4 a + b           a < 16, b < 4
//   k = k1 + 16 c + 256 d k1 < 16, c < 16, d < 4 (natural frequency index)
//   Z[k] = sum_b W_4^{bd} W_64^{bc} sum_a W_16^{ac} [ w^u W_1024^{u k1} sum_m (z'[u + 64 m] e^{i pi m / 32}) W_16^{m k1} ]
//   stage F1  thread u:        16-point DFT over m (per-register twist e^{i pi m/32}), times T1[u][k1] = exp(i pi u (1 - 4 k1) / 2048)
//   stage F2  thread (k1, b):  16-point DFT over a, times W_64^{bc}
//   stage MAC thread (k1, c & 3) of output polynomial q: for c = (c & 3) + 4 e: 4-point DFT over b of BOTH input polynomials
//             (frequencies k = y + 64 (e + 4 d), y = k1 + 16 (c & 3): consecutive threads read consecutive key elements),
//             Fourier MAC with the staged GGSW, inverse 4-point DFT over d, times W_64^{-bc}
//   stage I2  thread (k1, b):  inverse 16-point DFT over c
//   stage I3  thread u:        times conj(T1), inverse 16-point DFT over k1, per-register untwist, rounding, accumulation
// A transform buffer holds element (k1, x) -- x = u, 4 c + b or 4 a + b depending on the stage -- at 65 k1 + x: rows padded by
// one element instead of an XOR swizzle, so that every stage works IN PLACE on slots only it touches, a thread's register
// index maps to a fixed address offset, and both access patterns (lanes along x, lanes along k1) are conflict-free 128-bit
// accesses.  One buffer of 2 x 1040 elements per sample is all the transform needs (the MAC stage reads both polynomials,
// then, behind a barrier, every thread overwrites the slots of its own).
//
// __host__ __device__ like br_wide.cuh: tests/emu/emu_w16.cpp runs these functions thread by thread.
#pragma once
#include "../../../fhe_regex_b200/csrc/br_core.cuh"

namespace fb {
namespace w16 {

constexpr int kThreads = 128;          // per sample
constexpr int kRow = 65;               // elements per buffer row (64 + 1 pad)
constexpr int kBufC2 = 16 * kRow;      // 1040 elements per polynomial
constexpr int kTwC2 = 48;              // twiddles per thread index: T1[16] | W_64^{bc}, c < 16 | W_64^{-b'(cl + 4 e)} at 32 + 4 e + b' - 1 (b' = 1..3, one pad per e)

FB_HD constexpr int pos(int k1, int x) { return kRow * k1 + x; }

FB_HD c2 mk(double x, double y) {
  c2 r;
  r.x = x;
  r.y = y;
  return r;
}
FB_HD c2 cadd(c2 a, c2 b) { return mk(a.x + b.x, a.y + b.y); }
FB_HD c2 csub(c2 a, c2 b) { return mk(a.x - b.x, a.y - b.y); }
FB_HD c2 cmul(c2 a, c2 b) { return mk(fb_fma(a.x, b.x, -(a.y * b.y)), fb_fma(a.x, b.y, a.y * b.x)); }
FB_HD c2 cmul_conj(c2 a, c2 b) { return mk(fb_fma(a.x, b.x, a.y * b.y), fb_fma(a.y, b.x, -(a.x * b.y))); }   // a * conj(b)
// a * (-i) forward, a * (+i) inverse
template <bool INV>
FB_HD c2 rot90(c2 a) {
  return INV ? mk(-a.y, a.x) : mk(a.y, -a.x);
}

// cos / sin of pi p / 32, p = 0 .. 16, as literals (folded at compile time in the unrolled stages)
FB_HD constexpr double cos32(int p) {
  return p == 0 ? 1.0 : p == 1 ? 0.99518472667219688624 : p == 2 ? 0.98078528040323044913 : p == 3 ? 0.95694033573220886494
       : p == 4 ? 0.92387953251128675613 : p == 5 ? 0.88192126434835502971 : p == 6 ? 0.83146961230254523708
       : p == 7 ? 0.77301045336273696081 : p == 8 ? 0.70710678118654752440 : p == 9 ? 0.63439328416364549822
       : p == 10 ? 0.55557023301960222474 : p == 11 ? 0.47139673682599764856 : p == 12 ? 0.38268343236508977173
       : p == 13 ? 0.29028467725446236764 : p == 14 ? 0.19509032201612826785 : p == 15 ? 0.09801714032956060199 : 0.0;
}
FB_HD constexpr double sin32(int p) { return cos32(16 - p); }

// 4-point DFT in place: y_d = sum_b x_b W_4^{bd} (forward) or W_4^{-bd} (inverse)
template <bool INV>
FB_HD void dft4(c2& x0, c2& x1, c2& x2, c2& x3) {
  const c2 t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3), t3 = rot90<INV>(csub(x1, x3));
  x0 = cadd(t0, t2);
  x2 = csub(t0, t2);
  x1 = cadd(t1, t3);
  x3 = csub(t1, t3);
}

// cos / sin of pi p / 32 for p in [0, 64)
FB_HD constexpr double cos64(int p) {
  return p <= 16 ? cos32(p) : p <= 32 ? -cos32(32 - p) : p <= 48 ? -cos32(p - 32) : cos32(64 - p);
}
FB_HD constexpr double sin64(int p) {
  return p <= 16 ? sin32(p) : p <= 32 ? sin32(32 - p) : p <= 48 ? -sin32(p - 32) : -sin32(64 - p);
}
// x * W_16^{P} (forward) or W_16^{-P} (inverse), P a compile-time constant; W_16^P = cos(pi 4P / 32) - i sin(pi 4P / 32)
template <bool INV, int P>
FB_HD c2 mul_w16(c2 a) {
  if (P == 0) return a;
  if (P == 4) return rot90<INV>(a);
  const double c = cos64(4 * P), s = sin64(4 * P);
  return INV ? mk(fb_fma(a.x, c, -(a.y * s)), fb_fma(a.y, c, a.x * s)) : mk(fb_fma(a.x, c, a.y * s), fb_fma(a.y, c, -(a.x * s)));
}

// 16-point DFT, natural order in and out: X[k] = sum_m x[m] W_16^{mk} (forward) / W_16^{-mk} (inverse)
// m = m1 + 4 m2, k = q1 + 4 q2:  X[q1 + 4 q2] = sum_m1 W_4^{m1 q2} W_16^{m1 q1} sum_m2 x[m1 + 4 m2] W_4^{m2 q1}
template <bool INV>
FB_HD void dft16(c2 (&x)[16]) {
#pragma unroll
  for (int m1 = 0; m1 < 4; m1++) dft4<INV>(x[m1], x[m1 + 4], x[m1 + 8], x[m1 + 12]);   // x[m1 + 4 q1] = inner sum
  x[1 + 4 * 1] = mul_w16<INV, 1>(x[1 + 4 * 1]);
  x[1 + 4 * 2] = mul_w16<INV, 2>(x[1 + 4 * 2]);
  x[1 + 4 * 3] = mul_w16<INV, 3>(x[1 + 4 * 3]);
  x[2 + 4 * 1] = mul_w16<INV, 2>(x[2 + 4 * 1]);
  x[2 + 4 * 2] = mul_w16<INV, 4>(x[2 + 4 * 2]);
  x[2 + 4 * 3] = mul_w16<INV, 6>(x[2 + 4 * 3]);
  x[3 + 4 * 1] = mul_w16<INV, 3>(x[3 + 4 * 1]);
  x[3 + 4 * 2] = mul_w16<INV, 6>(x[3 + 4 * 2]);
  x[3 + 4 * 3] = mul_w16<INV, 9>(x[3 + 4 * 3]);
  c2 y[16];
#pragma unroll
  for (int q1 = 0; q1 < 4; q1++) {
    c2 a = x[4 * q1], b = x[4 * q1 + 1], c = x[4 * q1 + 2], d = x[4 * q1 + 3];
    dft4<INV>(a, b, c, d);   // over m1 -> q2
    y[q1] = a;
    y[q1 + 4] = b;
    y[q1 + 8] = c;
    y[q1 + 12] = d;
  }
#pragma unroll
  for (int k = 0; k < 16; k++) x[k] = y[k];
}

// ---- stage F1 ---------------------------------------------------------------------------------------------------
// per-register twist e^{i pi m / 32}, 16-point DFT over m, twiddle T1[u][k1] (carries w^u), store row by row
FB_HD void fwd1_core(c2 (&x)[16], int u, const c2 (&t1)[16], c2* buf) {
#pragma unroll
  for (int m = 1; m < 16; m++) {
    const double cm = cos32(m), sm = sin32(m);
    x[m] = mk(fb_fma(x[m].x, cm, -(x[m].y * sm)), fb_fma(x[m].x, sm, x[m].y * cm));
  }
  dft16<false>(x);
#pragma unroll
  for (int k1 = 0; k1 < 16; k1++) buf[pos(k1, u)] = cmul(x[k1], t1[k1]);
}
// phase A + F1: digits of (acc X^a - acc) for the coefficients j = u + 64 m and j + 1024 this thread owns (own[2m], own[2m+1]:
// the thread that rounds a coefficient in stage I3 is the one that decomposes it here); accp: the shared copy for the rotated reads
FB_HD void fwd1(const uint32_t* accp, const uint32_t (&own)[32], uint32_t a, int u, const c2 (&t1)[16], c2* buf) {
  c2 x[16];
#pragma unroll
  for (int m = 0; m < 16; m++) {
    const uint32_t j = (uint32_t)u + 64u * m;
    x[m].x = pbs_digit32_cvt(rot_read32(accp, j, a) - own[2 * m]);
    x[m].y = pbs_digit32_cvt(rot_read32(accp, j + 1024u, a) - own[2 * m + 1]);
  }
  fwd1_core(x, u, t1, buf);
}

// ---- stages F2 and I2: thread v = k1 + 16 b, in place ---------------------------------------------------------------
// One body for both directions (the CMUX loop has to stay small: a lone warp per scheduler is fed from the instruction cache):
// the inverse 16-point DFT is conj o DFT o conj, exact and bit-identical to dft16<true> (IEEE rounding is sign-symmetric);
// sign = 0 forward (then times W_64^{bc}), sign = 0x80000000 inverse (no twiddle: T1 carries it in stage I3).
FB_HD double flip_sign(double x, uint32_t sign) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(__double2hiint(x) ^ (int)sign, __double2loint(x));
#else
  return sign ? -x : x;
#endif
}
FB_HD void stage2(c2* buf, int v, const c2 (&tw)[16], uint32_t sign) {
  const int k1 = v & 15, b = v >> 4;
  c2* row = buf + pos(k1, b);
  c2 x[16];
#pragma unroll
  for (int a = 0; a < 16; a++) {
    x[a] = row[4 * a];
    x[a].y = flip_sign(x[a].y, sign);
  }
  dft16<false>(x);
  if (sign == 0u) {
    row[0] = x[0];
#pragma unroll
    for (int c = 1; c < 16; c++) row[4 * c] = cmul(x[c], tw[c]);
  } else {
#pragma unroll
    for (int a = 0; a < 16; a++) row[4 * a] = mk(x[a].x, -x[a].y);
  }
}
FB_HD void fwd2(c2* buf, int v, const c2 (&tw)[16]) { stage2(buf, v, tw, 0u); }

// ---- stage MAC: thread y = k1 + 16 cl of output polynomial q --------------------------------------------------------
// in0 / in1: forward buffers of the two input polynomials; ggsw: [pin][q][1024] natural order; itw[3 e + b' - 1] = W_64^{-b'(cl + 4 e)}
// The stage works IN PLACE on the forward buffers, in two halves around a barrier: mac_fwd only reads (both polynomials: the
// thread (k1, cl) of the OTHER output polynomial reads the same slots), mac_store only writes the slots of this thread's own
// polynomial.  r[4 e + d]: MAC result at frequency y + 64 (e + 4 d).
FB_HD void mac_fwd(const c2* in0, const c2* in1, const c2* ggsw, int q, int y, c2 (&r)[16]) {
  const int k1 = y & 15, cl = y >> 4;
  const c2* g0 = ggsw + (size_t)(0 * 2 + q) * kHalfN + y;
  const c2* g1 = ggsw + (size_t)(1 * 2 + q) * kHalfN + y;
#pragma unroll
  for (int e = 0; e < 4; e++) {
    const int o = pos(k1, 4 * cl) + 16 * e;   // element (k1, 4 (cl + 4 e) + b) sits at o + b
    c2 z0[4], z1[4];
#pragma unroll
    for (int b = 0; b < 4; b++) {
      z0[b] = in0[o + b];
      z1[b] = in1[o + b];
    }
    dft4<false>(z0[0], z0[1], z0[2], z0[3]);
    dft4<false>(z1[0], z1[1], z1[2], z1[3]);
#pragma unroll
    for (int d = 0; d < 4; d++) {
      const c2 a = g0[64 * (e + 4 * d)], b = g1[64 * (e + 4 * d)];
      r[4 * e + d].x = fb_fma(-z1[d].y, b.y, fb_fma(z1[d].x, b.x, fb_fma(-z0[d].y, a.y, z0[d].x * a.x)));
      r[4 * e + d].y = fb_fma(z1[d].y, b.x, fb_fma(z1[d].x, b.y, fb_fma(z0[d].y, a.x, z0[d].x * a.y)));
    }
  }
}
// itw[4 e + b' - 1] = W_64^{-b'(cl + 4 e)}, b' = 1..3 (itw[4 e + 3] is padding)
FB_HD void mac_store(c2 (&r)[16], int y, const c2 (&itw)[16], c2* out) {
  const int k1 = y & 15, cl = y >> 4;
#pragma unroll
  for (int e = 0; e < 4; e++) {
    const int o = pos(k1, 4 * cl) + 16 * e;
    dft4<true>(r[4 * e], r[4 * e + 1], r[4 * e + 2], r[4 * e + 3]);
    out[o] = r[4 * e];
#pragma unroll
    for (int b = 1; b < 4; b++) out[o + b] = cmul(r[4 * e + b], itw[4 * e + b - 1]);
  }
}
// pointwise product variant for the negacyclic-product test (one polynomial, spectrum in natural order)
FB_HD void mul_inv1(const c2* in0, const c2* spec, int y, const c2 (&itw)[16], c2* out) {
  const int k1 = y & 15, cl = y >> 4;
#pragma unroll
  for (int e = 0; e < 4; e++) {
    const int o = pos(k1, 4 * cl) + 16 * e;
    c2 z[4];
#pragma unroll
    for (int b = 0; b < 4; b++) z[b] = in0[o + b];
    dft4<false>(z[0], z[1], z[2], z[3]);
#pragma unroll
    for (int d = 0; d < 4; d++) z[d] = cmul(z[d], spec[y + 64 * (e + 4 * d)]);
    dft4<true>(z[0], z[1], z[2], z[3]);
    out[o] = z[0];
#pragma unroll
    for (int b = 1; b < 4; b++) out[o + b] = cmul(z[b], itw[4 * e + b - 1]);
  }
}
// spectrum of one polynomial out of its forward buffer, natural order (tests: must equal the product's key conversion)
FB_HD void spectrum_of(const c2* in0, int y, c2* spec) {
  const int k1 = y & 15, cl = y >> 4;
  for (int e = 0; e < 4; e++) {
    const int o = pos(k1, 4 * cl) + 16 * e;
    c2 z[4];
    for (int b = 0; b < 4; b++) z[b] = in0[o + b];
    dft4<false>(z[0], z[1], z[2], z[3]);
    for (int d = 0; d < 4; d++) spec[y + 64 * (e + 4 * d)] = z[d];
  }
}

// ---- stage I2: thread v = k1 + 16 b, in place: inverse 16-point DFT over c -> a ---------------------------------------
FB_HD void inv2(c2* buf, int v) {
  c2 none[16];
#pragma unroll
  for (int c = 0; c < 16; c++) none[c] = mk(1.0, 0.0);
  stage2(buf, v, none, 0x80000000u);
}

// ---- stage I3 + phase C: thread u ------------------------------------------------------------------------------------
// times conj(T1), inverse 16-point DFT over k1 -> m, untwist e^{-i pi m / 32} / 1024, torus increments of coefficients
// j = u + 64 m (re) and j + 1024 (im)
FB_HD void inv3_increments(const c2* buf, int u, const c2 (&t1)[16], uint32_t (&inc)[32]) {
  c2 x[16];
#pragma unroll
  for (int k1 = 0; k1 < 16; k1++) x[k1] = cmul_conj(buf[pos(k1, u)], t1[k1]);
  dft16<true>(x);
#pragma unroll
  for (int m = 0; m < 16; m++) {
    const double cm = cos32(m) * (1.0 / 1024.0), sm = sin32(m) * (1.0 / 1024.0);
    const double re = (m == 0) ? x[m].x * cm : fb_fma(x[m].x, cm, x[m].y * sm);
    const double im = (m == 0) ? x[m].y * cm : fb_fma(x[m].y, cm, -(x[m].x * sm));
    inc[2 * m] = torus32_from_double(re);
    inc[2 * m + 1] = torus32_from_double(im);
  }
}
FB_HD void inv3_accumulate(const c2* buf, int u, const c2 (&t1)[16], uint32_t (&own)[32], uint32_t* accp) {
  uint32_t inc[32];
  inv3_increments(buf, u, t1, inc);
#pragma unroll
  for (int m = 0; m < 16; m++) {
    const int j = u + 64 * m;
    own[2 * m] += inc[2 * m];
    own[2 * m + 1] += inc[2 * m + 1];
    accp[j] = own[2 * m];
    accp[j + 1024] = own[2 * m + 1];
  }
}

// per-thread twiddles out of the table [entry][v] (entry < kTwC2), v = thread index within the polynomial
FB_HD void load_t1(c2 (&t1)[16], const c2* tab, int v) {
#pragma unroll
  for (int k = 0; k < 16; k++) t1[k] = tab[k * 64 + v];
}
FB_HD void load_f2(c2 (&tw)[16], const c2* tab, int v) {
#pragma unroll
  for (int k = 0; k < 16; k++) tw[k] = tab[(16 + k) * 64 + v];
}
FB_HD void load_i1(c2 (&tw)[16], const c2* tab, int v) {
#pragma unroll
  for (int k = 0; k < 16; k++) tw[k] = tab[(32 + k) * 64 + v];
}

// Host-side table: [kTwC2][64]
static inline void make_w16_table(c2* tab) {
  const long double pi = 3.141592653589793238462643383279502884L;
  auto e = [&](long double ang) {
    c2 v;
    v.x = (double)cosl(ang);
    v.y = (double)sinl(ang);
    return v;
  };
  for (int v = 0; v < 64; v++) {
    const int hi = v >> 4;   // b of stages F2 / I2, cl of the MAC stage
    for (int k1 = 0; k1 < 16; k1++) tab[k1 * 64 + v] = e(pi * (long double)(v * (1 - 4 * k1)) / 2048.0L);
    for (int c = 0; c < 16; c++) tab[(16 + c) * 64 + v] = e(-2.0L * pi * (long double)(hi * c) / 64.0L);
    for (int ee = 0; ee < 4; ee++) {
      for (int b = 1; b < 4; b++) tab[(32 + 4 * ee + b - 1) * 64 + v] = e(2.0L * pi * (long double)(b * (hi + 4 * ee)) / 64.0L);
      tab[(32 + 4 * ee + 3) * 64 + v] = e(0.0L);
    }
  }
}

}  // namespace w16
}  // namespace fb
