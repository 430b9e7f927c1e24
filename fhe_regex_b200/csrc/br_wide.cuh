// br_wide.cuh -- per-thread building blocks of the LATENCY variant of the blind rotation (K2-K4):
// one sample spread over a whole CTA of 256 threads, for the narrow DAG levels of a regex match
// (engine.rs:22-35: the levels near the final bitor fold hold a handful of PBS each, and a level costs one
// blind-rotation latency whatever its width).  The throughput kernel (br_core.cuh) gives a thread a 32-point
// transform, i.e. ~4 800 instructions per warp and CMUX step, and a lone sample per SM is bound by
// instruction supply; here a thread owns 8 complex points and a step is ~1 400 instructions per warp.
//
// Same transform as br_core.cuh (so both kernels share one Fourier bootstrapping key in natural frequency
// order):  z[j] = (c[j] + i c[j+1024]) * w^j, w = exp(i pi/2048);  Z[k] = sum_j z[j] exp(-2 pi i jk/1024).
// 1024 = 8 x 8 x 8 x 2 as Stockham autosort stages through two shared-memory buffers:
//   stage with sub-length n, stride s, radix R, m = n/R:  butterfly (p < m, q < s) reads x[q + s(p + m r)],
//   r < R, and writes W_n^{pk} * DFT_R(x)_k to y[q + s(R p + k)].
// Polynomial P of the sample is handled by threads 128P .. 128P+127 (t = thread & 127):
//   forward  stage 1 (n=1024, s=1):   p = t            fused with the decomposition (phase A)
//            stage 2 (n=128,  s=8):   q = t&7,  p = t>>3
//            stage 3 (n=16,   s=64):  q = t&63, p = t>>6
//            stage 4 (n=2,    s=512): twiddle-free; thread t owns k = t + 128 m -- fused into the Fourier MAC
//   inverse  the same four stages with conjugate twiddles: stage 1 fused with the MAC, stage 4 with the
//            rounding/accumulation (phase C).
// The negacyclic twist w^j splits as w^t (per thread, folded into the stage-1 twiddles) times
// exp(i pi m/16) (compile-time, per register); the untwist w^-j / 1024 likewise (per-thread parts folded into
// the inverse stage-2 and stage-3 twiddles).  Per-thread twiddles live in registers for the whole kernel.
// Buffers are XOR-swizzled on 16-byte elements (idx ^ ((idx >> 3) & 7)): every stage reads and writes
// conflict-free with 128-bit accesses.
//
// __host__ __device__ like br_core.cuh: tests/emu/emu_wide.cpp runs these functions thread by thread.
#pragma once
#include "br_core.cuh"

namespace fb {
namespace wide {

constexpr int kThreads = 256;
constexpr int kTwRegs = 30;                    // c2 per thread: w1f[8] w1i[7] w2f[7] w2i[8]
constexpr int kTabC2 = kTwRegs * 128;

struct Tw {
  c2 w1f[8];   // exp(i pi t (1 - 4k) / 2048)                      k = 0..7
  c2 w1i[7];   // exp(+2 pi i t k / 1024)                          k = 1..7
  c2 w2f[7];   // exp(-2 pi i p2 k / 128)                          k = 1..7
  c2 w2i[8];   // exp(+2 pi i p2 k / 128 - i pi (q2 + 8k) / 2048) / 1024   k = 0..7
};

FB_HD int swz(int idx) { return idx ^ ((idx >> 3) & 7); }
// Where element idx of a transform buffer lives.  Both layouts make every stage's 128-bit accesses conflict-free (a quarter-warp
// either walks the low three bits of idx inside one group of 8, or walks the groups at a fixed low part).
//   LSwz  XOR swizzle inside a buffer of 1024 elements: no extra memory (br_wide2.cu has none to spare), but an address is a
//         LOP3 / shift / add chain per access, recomputed every step (the register file cannot hold 150 addresses);
//   LPad  groups of 8 elements padded to 9 (idx + idx / 8, 1152 elements per polynomial): idx is (thread part) + (compile-time
//         part), and so is its position -- every access of a stage is [one base register + immediate].
#if defined(__CUDACC__)
#define FB_HDM static __host__ __device__ __forceinline__
#else
#define FB_HDM static inline
#endif
// at2(a, c): position of element a + c, a the per-thread part and c the compile-time part of the index (in every use either c is a
// multiple of 8 or a is and c < 8, so that (a + c) / 8 = a / 8 + c / 8)
struct LSwz {
  static constexpr int kBuf = 1024;
  FB_HDM int at(int idx) { return idx ^ ((idx >> 3) & 7); }
  FB_HDM int at2(int a, int c) { return at(a + c); }
};
struct LPad {
  static constexpr int kBuf = 1152;
  FB_HDM int at(int idx) { return idx + (idx >> 3); }
  FB_HDM int at2(int a, int c) { return at(a) + c + (c >> 3); }
};

FB_HD c2 mk(double x, double y) {
  c2 r;
  r.x = x;
  r.y = y;
  return r;
}
FB_HD c2 cadd(c2 a, c2 b) { return mk(a.x + b.x, a.y + b.y); }
FB_HD c2 csub(c2 a, c2 b) { return mk(a.x - b.x, a.y - b.y); }
FB_HD c2 cmul(c2 a, c2 b) { return mk(fb_fma(a.x, b.x, -(a.y * b.y)), fb_fma(a.x, b.y, a.y * b.x)); }
// a * (-i) forward, a * (+i) inverse
template <bool INV>
FB_HD c2 rot90(c2 a) {
  return INV ? mk(-a.y, a.x) : mk(a.y, -a.x);
}

// in-place 8-point DFT, natural order in and out; forward kernel exp(-2 pi i rk/8), inverse exp(+2 pi i rk/8)
template <bool INV>
FB_HD void dft8(c2 (&x)[8]) {
  const double h = 0.70710678118654752440;
  const c2 a0 = cadd(x[0], x[4]), a1 = csub(x[0], x[4]);
  const c2 a2 = cadd(x[2], x[6]), a3 = rot90<INV>(csub(x[2], x[6]));
  const c2 a4 = cadd(x[1], x[5]), a5 = csub(x[1], x[5]);
  const c2 a6 = cadd(x[3], x[7]), a7 = rot90<INV>(csub(x[3], x[7]));
  const c2 b0 = cadd(a0, a2), b2 = csub(a0, a2), b1 = cadd(a1, a3), b3 = csub(a1, a3);
  const c2 b4 = cadd(a4, a6), b6 = rot90<INV>(csub(a4, a6));
  const c2 o1 = cadd(a5, a7), o3 = csub(a5, a7);
  // o1 * W8^1, o3 * W8^3:  forward W8 = (1 - i)/sqrt2, W8^3 = (-1 - i)/sqrt2; inverse the conjugates
  const c2 b5 = INV ? mk((o1.x - o1.y) * h, (o1.x + o1.y) * h) : mk((o1.x + o1.y) * h, (o1.y - o1.x) * h);
  const c2 b7 = INV ? mk(-(o3.x + o3.y) * h, (o3.x - o3.y) * h) : mk((o3.y - o3.x) * h, -(o3.x + o3.y) * h);
  x[0] = cadd(b0, b4);
  x[4] = csub(b0, b4);
  x[2] = cadd(b2, b6);
  x[6] = csub(b2, b6);
  x[1] = cadd(b1, b5);
  x[5] = csub(b1, b5);
  x[3] = cadd(b3, b7);
  x[7] = csub(b3, b7);
}

// cos/sin(pi m / 16), m = 0..7: the per-register part of the negacyclic twist (j = t + 128 m)
#define FB_WIDE_C16(m)                                                                                        \
  ((m) == 0 ? 1.0 : (m) == 1 ? 0.98078528040323044913 : (m) == 2 ? 0.92387953251128675613                     \
   : (m) == 3 ? 0.83146961230254523708 : (m) == 4 ? 0.70710678118654752440 : (m) == 5 ? 0.55557023301960222474 \
   : (m) == 6 ? 0.38268343236508977173 : 0.19509032201612826785)
#define FB_WIDE_S16(m)                                                                                        \
  ((m) == 0 ? 0.0 : (m) == 1 ? 0.19509032201612826785 : (m) == 2 ? 0.38268343236508977173                     \
   : (m) == 3 ? 0.55557023301960222474 : (m) == 4 ? 0.70710678118654752440 : (m) == 5 ? 0.83146961230254523708 \
   : (m) == 6 ? 0.92387953251128675613 : 0.98078528040323044913)

// twist by exp(i pi m/16) per register, 8-point DFT, stage-1 twiddle (carries w^t), store
template <class L = LSwz>
FB_HD void fwd_stage1_core(c2 (&x)[8], int t, const c2 (&w1f)[8], c2* out) {
#pragma unroll
  for (int m = 1; m < 8; m++) {
    const double cm = FB_WIDE_C16(m), sm = FB_WIDE_S16(m);
    x[m] = mk(fb_fma(x[m].x, cm, -(x[m].y * sm)), fb_fma(x[m].x, sm, x[m].y * cm));
  }
  dft8<false>(x);
#pragma unroll
  for (int k = 0; k < 8; k++) out[L::at2(8 * t, k)] = cmul(x[k], w1f[k]);
}
template <class L = LSwz>
FB_HD void fwd_stage1_core(c2 (&x)[8], int t, const Tw& tw, c2* out) { fwd_stage1_core<L>(x, t, tw.w1f, out); }

// phase A + forward stage 1: digits of (acc X^a - acc) for coefficients j = t + 128 m and j + 1024.
// own[2m], own[2m+1]: this thread's accumulator words of those coefficients (registers; the thread that rounds
// coefficient j in phase C is the one that decomposes it here); accp: the shared copy, for the rotated reads
template <class L = LSwz>
FB_HD void fwd_stage1(const uint32_t* accp, const uint32_t (&own)[16], uint32_t a, int t, const c2 (&w1f)[8], c2* out) {
  c2 x[8];
#pragma unroll
  for (int m = 0; m < 8; m++) {
    const uint32_t j = (uint32_t)t + 128u * m;
    x[m].x = pbs_digit32(rot_read32(accp, j, a) - own[2 * m]);
    x[m].y = pbs_digit32(rot_read32(accp, j + 1024u, a) - own[2 * m + 1]);
  }
  fwd_stage1_core<L>(x, t, w1f, out);
}
template <class L = LSwz>
FB_HD void fwd_stage1(const uint32_t* accp, const uint32_t (&own)[16], uint32_t a, int t, const Tw& tw, c2* out) {
  fwd_stage1<L>(accp, own, a, t, tw.w1f, out);
}

template <class L = LSwz>
FB_HD void fwd_stage2(const c2* in, c2* out, int t, const c2 (&w2f)[7]) {
  const int q = t & 7, p = t >> 3;
  c2 x[8];
#pragma unroll
  for (int r = 0; r < 8; r++) x[r] = in[L::at2(q + 8 * p, 128 * r)];
  dft8<false>(x);
  out[L::at2(q + 64 * p, 0)] = x[0];
#pragma unroll
  for (int k = 1; k < 8; k++) out[L::at2(q + 64 * p, 8 * k)] = cmul(x[k], w2f[k - 1]);
}
template <class L = LSwz>
FB_HD void fwd_stage2(const c2* in, c2* out, int t, const Tw& tw) { fwd_stage2<L>(in, out, t, tw.w2f); }

// stage-3 twiddles depend on p3 = t >> 6 only, which is warp-uniform: compile-time constants behind a uniform branch
// forward W16^{p3 k} = exp(-i pi p3 k / 8); inverse exp(+i pi p3 k / 8) * w^{-64 (k & 1)} (the per-parity part of
// the untwist; see inv_stage4)
#define FB_W3_C(k) ((k) == 1 ? 0.923879532511286756128 : (k) == 2 ? 0.707106781186547524401 : (k) == 3 ? 0.382683432365089771728 \
                    : (k) == 5 ? -0.382683432365089771728 : (k) == 6 ? -0.707106781186547524401 : -0.923879532511286756128)
#define FB_W3_S(k) ((k) == 1 ? 0.382683432365089771728 : (k) == 2 ? 0.707106781186547524401 : (k) == 3 ? 0.923879532511286756128 \
                    : (k) == 5 ? 0.923879532511286756128 : (k) == 6 ? 0.707106781186547524401 : 0.382683432365089771728)
#define FB_W3I_C(k) ((k) == 1 ? 0.956940335732208864936 : (k) == 2 ? 0.707106781186547524401 : (k) == 3 ? 0.471396736825997648556 \
                     : (k) == 5 ? -0.290284677254462367636 : (k) == 6 ? -0.707106781186547524401 : -0.881921264348355029713)
#define FB_W3I_S(k) ((k) == 1 ? 0.290284677254462367636 : (k) == 2 ? 0.707106781186547524401 : (k) == 3 ? 0.881921264348355029713 \
                     : (k) == 5 ? 0.956940335732208864936 : (k) == 6 ? 0.707106781186547524401 : 0.471396736825997648556)

template <int P3, class L = LSwz>
FB_HD void fwd_stage3_p(const c2* in, c2* out, int q) {
  c2 x[8];
#pragma unroll
  for (int r = 0; r < 8; r++) x[r] = in[L::at2(q, 64 * P3 + 128 * r)];
  dft8<false>(x);
#pragma unroll
  for (int k = 0; k < 8; k++) {
    c2 y = x[k];
    if (P3 == 1 && k == 4) y = rot90<false>(x[k]);
    else if (P3 == 1 && k != 0) y = cmul(x[k], mk(FB_W3_C(k), -FB_W3_S(k)));
    out[L::at2(q, 512 * P3 + 64 * k)] = y;
  }
}
template <class L = LSwz>
FB_HD void fwd_stage3(const c2* in, c2* out, int t) {
  if ((t >> 6) == 0) fwd_stage3_p<0, L>(in, out, t & 63);
  else fwd_stage3_p<1, L>(in, out, t & 63);
}

// forward stage 4 (a+b, a-b) of both polynomials + Fourier MAC with the staged GGSW + inverse stage 1.
// Thread (column qo, t) produces out_qo[k] = X_0[k] G[0][qo][k] + X_1[k] G[1][qo][k] for k = t + 128 m.
// The GGSW values do not depend on the other threads: the first NPRE of the four groups (4 values each) may be
// fetched before the barrier that precedes this stage (mac_prefetch) and are then passed in gpre.
template <int NPRE>
FB_HD void mac_prefetch(const c2* ggsw, int qo, int t, c2 (&gpre)[4 * (NPRE > 0 ? NPRE : 1)]) {
  const c2* g0 = ggsw + (size_t)(0 * 2 + qo) * kHalfN;
  const c2* g1 = ggsw + (size_t)(1 * 2 + qo) * kHalfN;
#pragma unroll
  for (int u = 0; u < NPRE; u++) {
    const int k = t + 128 * u;
    gpre[4 * u] = g0[k];
    gpre[4 * u + 1] = g1[k];
    gpre[4 * u + 2] = g0[k + 512];
    gpre[4 * u + 3] = g1[k + 512];
  }
}
template <int NPRE, class L = LSwz>
FB_HD void mac_inv_stage1(const c2* in0, const c2* in1, const c2* ggsw, const c2* gpre, int qo, int t, const c2 (&w1i)[7], c2* out) {
  c2 o[8];
  const c2* g0 = ggsw + (size_t)(0 * 2 + qo) * kHalfN;
  const c2* g1 = ggsw + (size_t)(1 * 2 + qo) * kHalfN;
#pragma unroll
  for (int u = 0; u < 4; u++) {
    const int k = t + 128 * u;
    const c2 a0 = in0[L::at2(t, 128 * u)], b0 = in0[L::at2(t, 128 * u + 512)];
    const c2 a1 = in1[L::at2(t, 128 * u)], b1 = in1[L::at2(t, 128 * u + 512)];
    const c2 x0l = cadd(a0, b0), x0h = csub(a0, b0), x1l = cadd(a1, b1), x1h = csub(a1, b1);
    const c2 gl0 = u < NPRE ? gpre[4 * u] : g0[k], gl1 = u < NPRE ? gpre[4 * u + 1] : g1[k];
    const c2 gh0 = u < NPRE ? gpre[4 * u + 2] : g0[k + 512], gh1 = u < NPRE ? gpre[4 * u + 3] : g1[k + 512];
    o[u].x = fb_fma(-x1l.y, gl1.y, fb_fma(x1l.x, gl1.x, fb_fma(-x0l.y, gl0.y, x0l.x * gl0.x)));
    o[u].y = fb_fma(x1l.y, gl1.x, fb_fma(x1l.x, gl1.y, fb_fma(x0l.y, gl0.x, x0l.x * gl0.y)));
    o[u + 4].x = fb_fma(-x1h.y, gh1.y, fb_fma(x1h.x, gh1.x, fb_fma(-x0h.y, gh0.y, x0h.x * gh0.x)));
    o[u + 4].y = fb_fma(x1h.y, gh1.x, fb_fma(x1h.x, gh1.y, fb_fma(x0h.y, gh0.x, x0h.x * gh0.y)));
  }
  dft8<true>(o);
  out[L::at2(8 * t, 0)] = o[0];
#pragma unroll
  for (int k = 1; k < 8; k++) out[L::at2(8 * t, k)] = cmul(o[k], w1i[k - 1]);
}
template <int NPRE, class L = LSwz>
FB_HD void mac_inv_stage1(const c2* in0, const c2* in1, const c2* ggsw, const c2* gpre, int qo, int t, const Tw& tw, c2* out) {
  mac_inv_stage1<NPRE, L>(in0, in1, ggsw, gpre, qo, t, tw.w1i, out);
}

// pointwise product variant for the negacyclic-product test (one polynomial, spectrum b in natural order)
template <class L = LSwz>
FB_HD void mul_inv_stage1(const c2* in0, const c2* spec, int t, const Tw& tw, c2* out) {
  c2 o[8];
#pragma unroll
  for (int u = 0; u < 4; u++) {
    const int k = t + 128 * u;
    const c2 a0 = in0[L::at2(t, 128 * u)], b0 = in0[L::at2(t, 128 * u + 512)];
    o[u] = cmul(cadd(a0, b0), spec[k]);
    o[u + 4] = cmul(csub(a0, b0), spec[k + 512]);
  }
  dft8<true>(o);
  out[L::at2(8 * t, 0)] = o[0];
#pragma unroll
  for (int k = 1; k < 8; k++) out[L::at2(8 * t, k)] = cmul(o[k], tw.w1i[k - 1]);
}

template <class L = LSwz>
FB_HD void inv_stage2(const c2* in, c2* out, int t, const c2 (&w2i)[8]) {
  const int q = t & 7, p = t >> 3;
  c2 x[8];
#pragma unroll
  for (int r = 0; r < 8; r++) x[r] = in[L::at2(q + 8 * p, 128 * r)];
  dft8<true>(x);
#pragma unroll
  for (int k = 0; k < 8; k++) out[L::at2(q + 64 * p, 8 * k)] = cmul(x[k], w2i[k]);
}
template <class L = LSwz>
FB_HD void inv_stage2(const c2* in, c2* out, int t, const Tw& tw) { inv_stage2<L>(in, out, t, tw.w2i); }

template <int P3, class L = LSwz>
FB_HD void inv_stage3_p(const c2* in, c2* out, int q) {
  c2 x[8];
#pragma unroll
  for (int r = 0; r < 8; r++) x[r] = in[L::at2(q, 64 * P3 + 128 * r)];
  dft8<true>(x);
#pragma unroll
  for (int k = 0; k < 8; k++) {
    c2 y = x[k];
    if (P3 == 0) {
      if (k & 1) y = cmul(x[k], mk(0.995184726672196886245, -0.0980171403295606019942));   // w^-64
    } else {
      if (k == 4) y = rot90<true>(x[k]);
      else if (k != 0) y = cmul(x[k], mk(FB_W3I_C(k), FB_W3I_S(k)));
    }
    out[L::at2(q, 512 * P3 + 64 * k)] = y;
  }
}
template <class L = LSwz>
FB_HD void inv_stage3(const c2* in, c2* out, int t) {
  if ((t >> 6) == 0) inv_stage3_p<0, L>(in, out, t & 63);
  else inv_stage3_p<1, L>(in, out, t & 63);
}

// inverse stage 4 + per-register untwist exp(-i pi m/16): torus increments of coefficients j = t + 128 m (re)
// and j + 1024 (im); everything else of the untwist and the 1/1024 is already in the data
template <class L = LSwz>
FB_HD void inv_stage4(const c2* in, int t, uint32_t (&inc_re)[8], uint32_t (&inc_im)[8]) {
  c2 z[8];
#pragma unroll
  for (int u = 0; u < 4; u++) {
    const c2 a = in[L::at2(t, 128 * u)], b = in[L::at2(t, 128 * u + 512)];
    z[u] = cadd(a, b);
    z[u + 4] = csub(a, b);
  }
#pragma unroll
  for (int m = 0; m < 8; m++) {
    const double cm = FB_WIDE_C16(m), sm = FB_WIDE_S16(m);
    const double re = (m == 0) ? z[m].x : fb_fma(z[m].x, cm, z[m].y * sm);
    const double im = (m == 0) ? z[m].y : fb_fma(z[m].y, cm, -(z[m].x * sm));
    inc_re[m] = torus32_from_double(re);
    inc_im[m] = torus32_from_double(im);
  }
}

template <class L = LSwz>
FB_HD void phaseC_accumulate(const c2* in, int t, uint32_t (&own)[16], uint32_t* accp) {
  uint32_t ire[8], iim[8];
  inv_stage4<L>(in, t, ire, iim);
#pragma unroll
  for (int m = 0; m < 8; m++) {
    const int j = t + 128 * m;
    own[2 * m] += ire[m];
    own[2 * m + 1] += iim[m];
    accp[j] = own[2 * m];
    accp[j + 1024] = own[2 * m + 1];
  }
}

// per-thread twiddles out of the table [entry][t] (entry < kTwRegs), t = thread & 127
FB_HD void load_tw(Tw& tw, const c2* tab, int t) {
#pragma unroll
  for (int k = 0; k < 8; k++) tw.w1f[k] = tab[(0 + k) * 128 + t];
#pragma unroll
  for (int k = 0; k < 7; k++) tw.w1i[k] = tab[(8 + k) * 128 + t];
#pragma unroll
  for (int k = 0; k < 7; k++) tw.w2f[k] = tab[(15 + k) * 128 + t];
#pragma unroll
  for (int k = 0; k < 8; k++) tw.w2i[k] = tab[(22 + k) * 128 + t];
}

// Host-side table: [kTwRegs][128] per-thread twiddles
static inline void make_wide_table(c2* tab) {
  const long double pi = 3.141592653589793238462643383279502884L;
  auto e = [&](long double ang, long double scale) {
    c2 v;
    v.x = (double)(cosl(ang) * scale);
    v.y = (double)(sinl(ang) * scale);
    return v;
  };
  for (int t = 0; t < 128; t++) {
    const int q2 = t & 7, p2 = t >> 3;
    for (int k = 0; k < 8; k++) tab[(0 + k) * 128 + t] = e(pi * (long double)(t * (1 - 4 * k)) / 2048.0L, 1.0L);
    for (int k = 1; k < 8; k++) tab[(8 + k - 1) * 128 + t] = e(2.0L * pi * (long double)(t * k) / 1024.0L, 1.0L);
    for (int k = 1; k < 8; k++) tab[(15 + k - 1) * 128 + t] = e(-2.0L * pi * (long double)(p2 * k) / 128.0L, 1.0L);
    for (int k = 0; k < 8; k++)
      tab[(22 + k) * 128 + t] = e(2.0L * pi * (long double)(p2 * k) / 128.0L - pi * (long double)(q2 + 8 * k) / 2048.0L, 1.0L / 1024.0L);
  }
}

}  // namespace wide
}  // namespace fb
