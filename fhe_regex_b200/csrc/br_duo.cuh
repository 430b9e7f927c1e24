// br_duo.cuh -- per-thread building blocks of the CLUSTER variant of the blind rotation (K2-K4):
// one PBS on a pair of SMs (thread-block cluster of 2 CTAs x 256 threads), for DAG levels so narrow that half the
// GPU would idle under the one-PBS-per-CTA kernel (the last levels of has_match's bitor fold, engine.rs:30-33, and
// every level once the variants are sharded over 8 GPUs).  CTA r owns polynomial r: it decomposes and transforms
// accumulator polynomial r forward, and computes output polynomial r of the external product; the two CTAs only
// trade their spectra (16 KiB each way per CMUX step) through distributed shared memory before the Fourier MAC.
//
// Same transform as br_core.cuh / br_wide.cuh (one Fourier bootstrapping key for all three kernels):
//   z[j] = (c[j] + i c[j+1024]) w^j, w = exp(i pi/2048);  Z[k] = sum_j z[j] exp(-2 pi i jk/1024), natural order.
// 1024 = 4^5 as Stockham autosort stages; a thread owns 4 complex points.  In stage s (stride S = 4^s, sub-length
// n = 1024/S, m = n/4) butterfly (p, q), q = b mod S, p = b div S, reads x[q + S(p + m r)] = x[b + 256 r] -- the
// same four addresses in every stage -- and writes W_n^{pk} DFT4(x)_k to y[q + S(4p + k)].
//   forward  stage 0 fused with the decomposition (phase A), stage 4 (twiddle-free) with the Fourier MAC
//   inverse  stage 0 fused with the MAC, stage 4 with the rounding/accumulation (phase C): thread b ends up with
//            coefficients b + 256k and b + 256k + 1024, the ones it decomposes next -- its own accumulator
//            words live in registers, shared memory only holds the copy the rotation reads.
// Twist w^j = w^b (per thread, folded into the stage-0 twiddles) x exp(i pi r/8) (compile-time); the untwist and
// the 1/1024 likewise (per-thread part folded into the inverse stage-3 twiddles).  The whole CMUX step is ~700
// instructions per thread: the loop fits the per-scheduler instruction cache.
// Buffers are XOR-swizzled on 16-byte elements: sigma(idx) = idx ^ h(idx >> 3), h(x) = (x & 3) | ((x & 2) << 1).
//
// __host__ __device__ like br_core.cuh: tests/emu/emu_duo.cpp runs these functions thread by thread.
#pragma once
#include "br_core.cuh"

namespace fb {
namespace duo {

constexpr int kThreads = 256;
constexpr int kTwRegs = 26;                // c2 per thread: f0[4] f1[3] f2[3] f3[3] i0[3] i1[3] i2[3] i3[4]
constexpr int kTabC2 = kTwRegs * 256;

struct Tw {
  c2 f0[4];   // exp(i pi b (1 - 4k) / 2048)                                  k = 0..3
  c2 f1[3];   // exp(-2 pi i p k / 256), p = b >> 2                           k = 1..3
  c2 f2[3];   // exp(-2 pi i p k / 64),  p = b >> 4
  c2 f3[3];   // exp(-2 pi i p k / 16),  p = b >> 6
  c2 i0[3];   // exp(+2 pi i b k / 1024)                                      k = 1..3
  c2 i1[3];   // exp(+2 pi i p k / 256)
  c2 i2[3];   // exp(+2 pi i p k / 64)
  c2 i3[4];   // exp(+2 pi i p k / 16 - i pi (q + 64k) / 2048) / 1024, q = b & 63   k = 0..3
};

FB_HD int sig(int idx) {
  const int x = idx >> 3;
  return idx ^ ((x & 3) | ((x & 2) << 1));
}

FB_HD c2 mk(double x, double y) {
  c2 r;
  r.x = x;
  r.y = y;
  return r;
}
FB_HD c2 cadd(c2 a, c2 b) { return mk(a.x + b.x, a.y + b.y); }
FB_HD c2 csub(c2 a, c2 b) { return mk(a.x - b.x, a.y - b.y); }
FB_HD c2 cmul(c2 a, c2 b) { return mk(fb_fma(a.x, b.x, -(a.y * b.y)), fb_fma(a.x, b.y, a.y * b.x)); }

// in-place 4-point DFT, natural order; forward kernel exp(-2 pi i rk/4), inverse exp(+2 pi i rk/4)
template <bool INV>
FB_HD void dft4(c2 (&x)[4]) {
  const c2 a0 = cadd(x[0], x[2]), a1 = csub(x[0], x[2]), a2 = cadd(x[1], x[3]), d = csub(x[1], x[3]);
  const c2 a3 = INV ? mk(-d.y, d.x) : mk(d.y, -d.x);   // d * (+i) inverse, d * (-i) forward
  x[0] = cadd(a0, a2);
  x[2] = csub(a0, a2);
  x[1] = cadd(a1, a3);
  x[3] = csub(a1, a3);
}

// cos/sin(pi r / 8), r = 0..3: the per-register part of the negacyclic twist (j = b + 256 r)
#define FB_DUO_C8(r) ((r) == 0 ? 1.0 : (r) == 1 ? 0.92387953251128675613 : (r) == 2 ? 0.70710678118654752440 : 0.38268343236508977173)
#define FB_DUO_S8(r) ((r) == 0 ? 0.0 : (r) == 1 ? 0.38268343236508977173 : (r) == 2 ? 0.70710678118654752440 : 0.92387953251128675613)

// forward stage 0 on raw (untwisted) folded inputs x[r] = c[b + 256r] + i c[b + 256r + 1024]
FB_HD void fwd_stage0_core(c2 (&x)[4], int b, const Tw& tw, c2* out) {
#pragma unroll
  for (int r = 1; r < 4; r++) {
    const double cr = FB_DUO_C8(r), sr = FB_DUO_S8(r);
    x[r] = mk(fb_fma(x[r].x, cr, -(x[r].y * sr)), fb_fma(x[r].x, sr, x[r].y * cr));
  }
  dft4<false>(x);
#pragma unroll
  for (int k = 0; k < 4; k++) out[sig(4 * b + k)] = cmul(x[k], tw.f0[k]);
}

// phase A + forward stage 0.  own[2r], own[2r+1]: this thread's accumulator words of coefficients b + 256r and
// b + 256r + 1024 (registers); accp: the shared copy of the whole polynomial, for the rotated reads
FB_HD void fwd_stage0(const uint32_t* accp, const uint32_t (&own)[8], uint32_t a, int b, const Tw& tw, c2* out) {
  c2 x[4];
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const uint32_t j = (uint32_t)b + 256u * r;
    x[r].x = pbs_digit32(rot_read32(accp, j, a) - own[2 * r]);
    x[r].y = pbs_digit32(rot_read32(accp, j + 1024u, a) - own[2 * r + 1]);
  }
  fwd_stage0_core(x, b, tw, out);
}

// middle stages s = 1, 2, 3 (stride S = 4^s): in -> out with twiddles w[k-1], k = 1..3
template <int S, bool INV>
FB_HD void mid_stage(const c2* in, c2* out, int b, const c2 (&w)[3]) {
  constexpr int SH = (S == 4) ? 2 : (S == 16) ? 4 : 6;
  const int q = b & (S - 1), p = b >> SH;
  c2 x[4];
#pragma unroll
  for (int r = 0; r < 4; r++) x[r] = in[sig(b + 256 * r)];
  dft4<INV>(x);
  const int o = q + 4 * S * p;
  out[sig(o)] = x[0];
#pragma unroll
  for (int k = 1; k < 4; k++) out[sig(o + S * k)] = cmul(x[k], w[k - 1]);
}
// inverse stage 3: all four outputs carry a factor (untwist part + 1/1024)
FB_HD void inv_stage3(const c2* in, c2* out, int b, const Tw& tw) {
  const int q = b & 63, p = b >> 6;
  c2 x[4];
#pragma unroll
  for (int r = 0; r < 4; r++) x[r] = in[sig(b + 256 * r)];
  dft4<true>(x);
#pragma unroll
  for (int k = 0; k < 4; k++) out[sig(q + 256 * p + 64 * k)] = cmul(x[k], tw.i3[k]);
}

// forward stage 4 (twiddle-free): spectrum values X[b + 256k], k = 0..3, into registers
FB_HD void fwd_stage4(const c2* in, int b, c2 (&X)[4]) {
#pragma unroll
  for (int r = 0; r < 4; r++) X[r] = in[sig(b + 256 * r)];
  dft4<false>(X);
}

// Fourier MAC for output polynomial `me` at frequencies b + 256k: out = X_own G[me][me] + X_peer G[1-me][me], then
// inverse stage 0.  g_own / g_peer: the two staged GGSW polynomials (natural order, unswizzled).
FB_HD void mac_inv_stage0(const c2 (&Xown)[4], const c2 (&Xpeer)[4], const c2* g_own, const c2* g_peer, int b, const Tw& tw, c2* out) {
  c2 o[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const c2 go = g_own[b + 256 * k], gp = g_peer[b + 256 * k];
    o[k].x = fb_fma(-Xpeer[k].y, gp.y, fb_fma(Xpeer[k].x, gp.x, fb_fma(-Xown[k].y, go.y, Xown[k].x * go.x)));
    o[k].y = fb_fma(Xpeer[k].y, gp.x, fb_fma(Xpeer[k].x, gp.y, fb_fma(Xown[k].y, go.x, Xown[k].x * go.y)));
  }
  dft4<true>(o);
  out[sig(4 * b)] = o[0];
#pragma unroll
  for (int k = 1; k < 4; k++) out[sig(4 * b + k)] = cmul(o[k], tw.i0[k - 1]);
}

// inverse stage 4 + per-register untwist exp(-i pi k/8) + rounding: accumulates into own[] and returns the new words
FB_HD void inv_stage4_accumulate(const c2* in, int b, uint32_t (&own)[8]) {
  c2 z[4];
#pragma unroll
  for (int r = 0; r < 4; r++) z[r] = in[sig(b + 256 * r)];
  dft4<true>(z);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const double ck = FB_DUO_C8(k), sk = FB_DUO_S8(k);
    const double re = (k == 0) ? z[k].x : fb_fma(z[k].x, ck, z[k].y * sk);
    const double im = (k == 0) ? z[k].y : fb_fma(z[k].y, ck, -(z[k].x * sk));
    own[2 * k] += torus32_from_double(re);
    own[2 * k + 1] += torus32_from_double(im);
  }
}

// per-thread twiddles out of the table [entry][b]
FB_HD void load_tw(Tw& tw, const c2* tab, int b) {
  int e = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) tw.f0[k] = tab[(e++) * 256 + b];
#pragma unroll
  for (int k = 0; k < 3; k++) tw.f1[k] = tab[(e++) * 256 + b];
#pragma unroll
  for (int k = 0; k < 3; k++) tw.f2[k] = tab[(e++) * 256 + b];
#pragma unroll
  for (int k = 0; k < 3; k++) tw.f3[k] = tab[(e++) * 256 + b];
#pragma unroll
  for (int k = 0; k < 3; k++) tw.i0[k] = tab[(e++) * 256 + b];
#pragma unroll
  for (int k = 0; k < 3; k++) tw.i1[k] = tab[(e++) * 256 + b];
#pragma unroll
  for (int k = 0; k < 3; k++) tw.i2[k] = tab[(e++) * 256 + b];
#pragma unroll
  for (int k = 0; k < 4; k++) tw.i3[k] = tab[(e++) * 256 + b];
}

// Host-side table: [kTwRegs][256]
static inline void make_duo_table(c2* tab) {
  const long double pi = 3.141592653589793238462643383279502884L;
  auto e = [&](long double ang, long double scale) {
    c2 v;
    v.x = (double)(cosl(ang) * scale);
    v.y = (double)(sinl(ang) * scale);
    return v;
  };
  for (int b = 0; b < 256; b++) {
    int en = 0;
    for (int k = 0; k < 4; k++) tab[(en++) * 256 + b] = e(pi * (long double)(b * (1 - 4 * k)) / 2048.0L, 1.0L);
    for (int k = 1; k < 4; k++) tab[(en++) * 256 + b] = e(-2.0L * pi * (long double)((b >> 2) * k) / 256.0L, 1.0L);
    for (int k = 1; k < 4; k++) tab[(en++) * 256 + b] = e(-2.0L * pi * (long double)((b >> 4) * k) / 64.0L, 1.0L);
    for (int k = 1; k < 4; k++) tab[(en++) * 256 + b] = e(-2.0L * pi * (long double)((b >> 6) * k) / 16.0L, 1.0L);
    for (int k = 1; k < 4; k++) tab[(en++) * 256 + b] = e(2.0L * pi * (long double)(b * k) / 1024.0L, 1.0L);
    for (int k = 1; k < 4; k++) tab[(en++) * 256 + b] = e(2.0L * pi * (long double)((b >> 2) * k) / 256.0L, 1.0L);
    for (int k = 1; k < 4; k++) tab[(en++) * 256 + b] = e(2.0L * pi * (long double)((b >> 4) * k) / 64.0L, 1.0L);
    for (int k = 0; k < 4; k++)
      tab[(en++) * 256 + b] = e(2.0L * pi * (long double)((b >> 6) * k) / 16.0L - pi * (long double)((b & 63) + 64 * k) / 2048.0L, 1.0L / 1024.0L);
  }
}

}  // namespace duo
}  // namespace fb
