// br_core.cuh -- per-thread building blocks of the blind-rotation kernel (K2/K3/K4 of SURVEY.md 2).
//
// Replaces, for the hot path under execution.rs:76,93,110,143,173,190, what tfhe-rs 0.2.0 does in
// blind_rotate_assign / add_external_product_assign (GGSW x GLWE external product with an f64
// negacyclic FFT).  Nothing here is translated from tfhe-rs: the transform is a 32x32 two-pass
// FFT whose passes run entirely in the registers of one thread, laid out for a warp per polynomial.
//
// The functions are __host__ __device__ so that tests/emu/ can run the exact index/twiddle logic
// lane by lane on the CPU (there is no GPU in the build container); the product only ever runs
// them inside the CUDA kernels of kernels.cu.
//
// Layout (one "sample" = one PBS in flight = 2 warps = 64 threads):
//   polynomial p (0 = GLWE mask, 1 = body), coefficient j in [0,2048)
//   folded complex point z[j'] = (c[j'] + i c[j'+1024]) * exp(i*pi*j'/2048), j' = 32*r + lane
//     -> thread `lane` of warp p holds register r = 0..31           (phase A / phase C)
//   pass 1: in-register DFT over r (fft32_fwd_twist, which also applies exp(i*pi*32r/2048)) -> k1
//   twiddle exp(i*pi*lane*(1-4*k1)/2048) rebuilt from 4 "lo" x 8 "hi" factors, transpose through shared
//   memory (XOR swizzle; the real and the imaginary plane one after the other)
//   pass 2: thread (p', k1) holds c = 0..31, in-register DFT over c -> k2   (phase B)
//     lanes 0-15 of warp w: p'=0, k1 = 16w + lane ; lanes 16-31: p'=1, k1 = 16w + lane - 16
//   frequency index k = k1 + 32*k2, the Fourier bootstrapping key is stored in that natural order.
//   The inverse runs the same two passes backwards with the conjugate twiddles.
// Measured on B200: reading the twiddles from a 16 KiB shared-memory table or from
// tensor memory is slower than the lo x hi rebuild -- shared-memory wavefronts are the scarce resource.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define FB_HD static __host__ __device__ __forceinline__
#else
#define FB_HD static inline
#endif

FB_HD double fb_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return fma(a, b, c);
#endif
}

#include "fft32_gen.h"

template <int... Ns> struct fb_iseq {};
template <int N, int... Ns> struct fb_make_iseq_impl : fb_make_iseq_impl<N - 1, N - 1, Ns...> {};
template <int... Ns> struct fb_make_iseq_impl<0, Ns...> { typedef fb_iseq<Ns...> type; };
template <int N> using fb_make_iseq = typename fb_make_iseq_impl<N>::type;

namespace fb {

constexpr int kLweN = 742;       // small LWE dimension
constexpr int kN = 2048;         // GLWE polynomial size (k = 1)
constexpr int kHalfN = 1024;     // complex points per transform
constexpr int kBig = kN + 1;     // words per big LWE ciphertext
constexpr int kSmall = kLweN + 1;
constexpr int kKsLevels = 5;
constexpr int kKsBaseLog = 3;
constexpr int kTabEntries = 12;  // 4 "lo" + 8 "hi" twiddle factors per lane

#if defined(__CUDACC__)
typedef double2 c2;
#else
struct alignas(16) c2 { double x, y; };
#endif

FB_HD constexpr int brev5(int v) {
  return ((v & 1) << 4) | ((v & 2) << 2) | (v & 4) | ((v & 8) >> 2) | ((v & 16) >> 4);
}

// T3: PBS modulus switch to Z_{2N} (round), result in [0, 4096]
FB_HD uint32_t modswitch(uint64_t x) {
  uint64_t t = x >> 51;
  t += t & 1ull;
  return (uint32_t)(t >> 1);
}

// coefficient j of acc * X^a, a in [0, 4096)
FB_HD uint64_t rot_read(const uint64_t* accp, uint32_t j, uint32_t a) {
  uint32_t idx = (j - a) & 4095u;
  uint64_t v = accp[idx & 2047u];
  return (idx & 2048u) ? (uint64_t)0 - v : v;
}

// ---- 32-bit accumulator -------------------------------------------------------------------------
// The accumulator of the blind rotation is kept on the top 32 torus bits.  Its increments come out of an
// f64 inverse transform whose values have magnitude ~2^25 (digits up to 2^22 times 4096 key terms), i.e. a
// granularity of ~2^-27 on the torus: the low word of a 64-bit accumulator would only ever collect zeros
// and rounding dust.  Rounding every increment to 2^-32 adds at most 742 * 2^-66 / 12 of variance
// (sigma < 2^-29) to an output whose FFT noise alone is ~2^-15.  With a 32-bit
// accumulator the base-2^23 decomposition of acc*X^a - acc below is exact.
FB_HD uint32_t rot_read32(const uint32_t* accp, uint32_t j, uint32_t a) {
  const uint32_t idx = (j - a) & 4095u;
  const uint32_t v = accp[idx & 2047u];
  return (idx & 2048u) ? 0u - v : v;
}

FB_HD double hilo_to_double(uint32_t hi, uint32_t lo) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double((int)hi, (int)lo);
#else
  const uint64_t bits = ((uint64_t)hi << 32) | lo;
  double d;
  memcpy(&d, &bits, sizeof d);
  return d;
#endif
}
FB_HD uint32_t double_lo32(double d) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__double2loint(d);
#else
  uint64_t bits;
  memcpy(&bits, &d, sizeof d);
  return (uint32_t)bits;
#endif
}

// balanced base-2^23 digit of a 32-bit torus difference, round(diff / 2^9) in [-2^22, 2^22], as a double.
// int -> double without the conversion unit and without the shift: d = diff + 2^8 + 2^31 (the rounding offset and a flip of the
// sign bit in one add) with its low 9 bits cleared is 512 (q + 2^22), q = floor((diff + 2^8) / 2^9); as the low mantissa word of
// a double with exponent 2^43 (ulp 2^-9) it reads 2^43 + 2^22 + q.  One add, one AND and one DADD per digit.
FB_HD double pbs_digit32(uint32_t diff) {
  const uint32_t lo = (diff + 0x80000100u) & 0xFFFFFE00u;
  return hilo_to_double(0x42A00000u, lo) - 8796097216512.0;  // 2^43 + 2^22
}

// decompose (acc*X^a - acc) of polynomial accp into the folded FFT input (digits; the twist is in fft32_fwd_twist)
FB_HD void phaseA_load32(double (&xr)[32], double (&xi)[32], const uint32_t* accp, uint32_t a, int lane) {
#pragma unroll
  for (int r = 0; r < 32; r++) {
    const uint32_t j = 32u * r + lane;
    xr[r] = pbs_digit32(rot_read32(accp, j, a) - accp[j]);
    xi[r] = pbs_digit32(rot_read32(accp, j + 1024u, a) - accp[j + 1024u]);
  }
}

// fractional part of t as a 32-bit torus word, round(frac(t) * 2^32) mod 2^32, |t| < 2^51; magic-number
// rounding on the FP64 pipe only (no F2I/FRND through the conversion unit)
FB_HD uint32_t torus32_from_double(double t) {
  const double kMagic = 6755399441055744.0;  // 1.5 * 2^52
  const double ti = (t + kMagic) - kMagic;   // rint(t)
  const double f = t - ti;                   // exact, in [-1/2, 1/2]
  return double_lo32(fb_fma(f, 4294967296.0, kMagic));
}

// phase C for one register: torus increments of coefficients 32r+lane (re) and 32r+lane+1024 (im)
FB_HD void phaseC_increments32(const double (&xr)[32], const double (&xi)[32], int r, uint32_t& inc0, uint32_t& inc1) {
  const double cr = fb_twist_cos(r) * (1.0 / 1024.0), sr = fb_twist_sin(r) * (1.0 / 1024.0);
  inc0 = torus32_from_double(fb_fma(xr[r], cr, xi[r] * sr));
  inc1 = torus32_from_double(fb_fma(xi[r], cr, -(xr[r] * sr)));
}

// Fourier MAC of one frequency point: out = x * b_own + partner * b_in, where `partner` is the other
// polynomial's spectrum value at the same frequency (lane ^ 16) and b_in = GGSW[other row][my column]
FB_HD void mac_point2(double& xr, double& xi, double pr, double pi, c2 b_own, c2 b_in) {
  double orr = xr * b_own.x;
  orr = fb_fma(-xi, b_own.y, orr);
  orr = fb_fma(pr, b_in.x, orr);
  orr = fb_fma(-pi, b_in.y, orr);
  double oi = xr * b_own.y;
  oi = fb_fma(xi, b_own.x, oi);
  oi = fb_fma(pr, b_in.y, oi);
  oi = fb_fma(pi, b_in.x, oi);
  xr = orr;
  xi = oi;
}

// same folding (untwisted) for a standard-domain key polynomial read as a signed torus value in [-1/2, 1/2)
// (key conversion K7; tfhe-rs forward_as_torus)
FB_HD void load_torus_poly(double (&xr)[32], double (&xi)[32], const uint64_t* poly, int lane) {
#pragma unroll
  for (int r = 0; r < 32; r++) {
    const uint32_t j = 32u * r + lane;
    xr[r] = (double)(int64_t)poly[j] * (1.0 / 18446744073709551616.0);
    xi[r] = (double)(int64_t)poly[j + 1024u] * (1.0 / 18446744073709551616.0);
  }
}

// index (in complex elements) of Fourier GGSW element: [i][row pin][poly jout][k]
FB_HD size_t fbsk_index(int i, int pin, int jout, int k) {
  return (((size_t)i * 2 + pin) * 2 + jout) * kHalfN + k;
}

// ---- split transposes: the real and the imaginary planes go through one [2][kPlaneDoubles] double buffer one
// after the other (half the shared memory of a complex buffer; 4 two-warp barriers per transpose).
// Element (row k1, column c) of polynomial p lives at p*kPlaneDoubles + k1*34 + c: rows padded to 34 doubles instead of an
// XOR swizzle.  Every access is then [one base register + immediate] (the swizzle cost a LOP3 + LEA per access, ~250
// instructions per CMUX step, in a loop whose instruction supply is the scarce resource), a row is 16-byte aligned so the
// row side moves two columns per 128-bit access, and both sides stay conflict-free: a column access touches 32 consecutive
// doubles; in a row access a quarter-warp holds 8 consecutive rows, 8 x 272 bytes apart = 8 different 16-byte bank groups.
constexpr int kPlaneRow = 34;
constexpr int kPlaneDoubles = 32 * kPlaneRow;   // 1088 per polynomial

// forward inter-pass twiddle in place: register q (row k1 = brev5(q)) *= exp(i*pi*lane*(1-4*k1)/2048)
FB_HD void fwd_twiddle_inplace(double (&xr)[32], double (&xi)[32], const c2* tab_f, int lane) {
  c2 lo[4];
#pragma unroll
  for (int l = 0; l < 4; l++) lo[l] = tab_f[l * 32 + lane];
#pragma unroll
  for (int h = 0; h < 8; h++) {
    const c2 hi = tab_f[(4 + h) * 32 + lane];
#pragma unroll
    for (int l = 0; l < 4; l++) {
      const int q = brev5(4 * h + l);
      double tr = lo[l].x, ti = lo[l].y;
      if (h != 0) {
        tr = fb_fma(lo[l].x, hi.x, -(lo[l].y * hi.y));
        ti = fb_fma(lo[l].x, hi.y, lo[l].y * hi.x);
      }
      const double yr = fb_fma(xr[q], tr, -(xi[q] * ti));
      xi[q] = fb_fma(xr[q], ti, xi[q] * tr);
      xr[q] = yr;
    }
  }
}
// column writer (thread = column `lane` of polynomial p, register q = row brev5(q))
FB_HD void col_store_brev(const double (&x)[32], double* plane_p, int lane) {
#pragma unroll
  for (int q = 0; q < 32; q++) {
    const int k1 = brev5(q);
    plane_p[k1 * kPlaneRow + lane] = x[q];
  }
}
// row reader (thread = row k1 of polynomial pp, register c = column c)
FB_HD void row_load(double (&x)[32], const double* plane_pp, int k1) {
  const c2* row = reinterpret_cast<const c2*>(plane_pp + k1 * kPlaneRow);
#pragma unroll
  for (int h = 0; h < 16; h++) {
    const c2 v = row[h];
    x[2 * h] = v.x;
    x[2 * h + 1] = v.y;
  }
}
// inverse inter-pass twiddle in place: register c *= exp(-i*pi*c*(1-4*k1)/2048)
FB_HD void inv_twiddle_inplace(double (&xr)[32], double (&xi)[32], const c2* tab_i, int k1) {
  c2 lo[4];
#pragma unroll
  for (int l = 0; l < 4; l++) lo[l] = tab_i[l * 32 + k1];
#pragma unroll
  for (int h = 0; h < 8; h++) {
    const c2 hi = tab_i[(4 + h) * 32 + k1];
#pragma unroll
    for (int l = 0; l < 4; l++) {
      const int c = 4 * h + l;
      double tr = lo[l].x, ti = lo[l].y;
      if (h != 0) {
        tr = fb_fma(lo[l].x, hi.x, -(lo[l].y * hi.y));
        ti = fb_fma(lo[l].x, hi.y, lo[l].y * hi.x);
      }
      const double yr = fb_fma(xr[c], tr, -(xi[c] * ti));
      xi[c] = fb_fma(xr[c], ti, xi[c] * tr);
      xr[c] = yr;
    }
  }
}
// row writer (thread = row k1, register c = column c)
FB_HD void row_store(const double (&x)[32], double* plane_pp, int k1) {
  c2* row = reinterpret_cast<c2*>(plane_pp + k1 * kPlaneRow);
#pragma unroll
  for (int h = 0; h < 16; h++) {
    c2 v;
    v.x = x[2 * h];
    v.y = x[2 * h + 1];
    row[h] = v;
  }
}
// column reader into the bit-reversed register order fft32_dit_inv wants (register q = row brev5(q))
FB_HD void col_load_brev(double (&x)[32], const double* plane_p, int lane) {
#pragma unroll
  for (int q = 0; q < 32; q++) {
    const int k1 = brev5(q);
    x[q] = plane_p[k1 * kPlaneRow + lane];
  }
}

// ---- planes aliased into the accumulator copies (blind_rotate_fused_kernel with more than 4 PBS per SM) --------
// Between phase A (last rotated read of a step) and phase C (which rewrites every word) the 8 KiB shared-memory copy of a
// polynomial's accumulator holds nothing that is read again, and the plane of polynomial p is written by the warp that owns
// polynomial p only (column side) before anybody reads it.  The plane of polynomial p therefore lives INSIDE the accumulator
// copy of polynomial p: rows 0..29 (30 x 272 = 8160 bytes) in the copy itself, rows 30 and 31 in a 640-byte overflow block,
// at +96 and +368 of a 128-byte aligned address -- the 16-byte bank groups 6 and 7 they have in the contiguous layout, so the
// quarter-warp that holds rows 24..31 stays conflict-free.  Shared memory per PBS drops from 33 KiB to 17.25 KiB.
// Pointer arguments: `main` = the polynomial's accumulator copy as doubles, `ovf` = its overflow block + 96 bytes.
constexpr int kPlaneMainRows = 30;
constexpr int kPlaneOvfBytes = 640;
constexpr int kPlaneOvfLead = 96;
FB_HD double* plane_al_row(double* main, double* ovf, int k1) {
  return k1 < kPlaneMainRows ? main + k1 * kPlaneRow : ovf + (k1 - kPlaneMainRows) * kPlaneRow;
}
FB_HD void col_store_brev_al(const double (&x)[32], double* main_l, double* ovf_l) {   // both with the lane folded in
#pragma unroll
  for (int q = 0; q < 32; q++) {
    const int k1 = brev5(q);
    if (k1 < kPlaneMainRows) main_l[k1 * kPlaneRow] = x[q];
    else ovf_l[(k1 - kPlaneMainRows) * kPlaneRow] = x[q];
  }
}
FB_HD void col_load_brev_al(double (&x)[32], const double* main_l, const double* ovf_l) {
#pragma unroll
  for (int q = 0; q < 32; q++) {
    const int k1 = brev5(q);
    x[q] = k1 < kPlaneMainRows ? main_l[k1 * kPlaneRow] : ovf_l[(k1 - kPlaneMainRows) * kPlaneRow];
  }
}
// row side: row_load(x, row, 0) / row_store(x, row, 0) with row = plane_al_row(main of pp, ovf of pp, k1)

// ---- full inter-pass twiddle tables (blind_rotate_fused_kernel, V & 64) ---------------------------------------------------
// The lo x hi rebuild costs 28 complex products = 112 FP64 instructions per direction, per thread and per CMUX step.  With the
// planes inside the accumulator copies there is room for both full tables (2 x 16 KiB): the kernel builds them ONCE per launch
// with the very same two-FMA products (fb_full_twiddle), so the values -- and therefore the outputs -- are bit-identical.
FB_HD c2 fb_full_twiddle(const c2* tab, int e, int x) {   // entry e = 4h + l of tab_f (x = lane) or tab_i (x = k1)
  const int h = e >> 2, l = e & 3;
  const c2 lo = tab[l * 32 + x];
  if (h == 0) return lo;
  const c2 hi = tab[(4 + h) * 32 + x];
  c2 t;
  t.x = fb_fma(lo.x, hi.x, -(lo.y * hi.y));
  t.y = fb_fma(lo.x, hi.y, lo.y * hi.x);
  return t;
}
// forward: register q (row k1 = brev5(q)) *= full_f[k1][lane]
FB_HD void fwd_twiddle_full(double (&xr)[32], double (&xi)[32], const c2* full_f_lane) {
#pragma unroll
  for (int k1 = 0; k1 < 32; k1++) {
    const int q = brev5(k1);
    const c2 t = full_f_lane[k1 * 32];
    const double yr = fb_fma(xr[q], t.x, -(xi[q] * t.y));
    xi[q] = fb_fma(xr[q], t.y, xi[q] * t.x);
    xr[q] = yr;
  }
}
// inverse: register c *= full_i[c][k1]
FB_HD void inv_twiddle_full(double (&xr)[32], double (&xi)[32], const c2* full_i_k1) {
#pragma unroll
  for (int c = 0; c < 32; c++) {
    const c2 t = full_i_k1[c * 32];
    const double yr = fb_fma(xr[c], t.x, -(xi[c] * t.y));
    xi[c] = fb_fma(xr[c], t.y, xi[c] * t.x);
    xr[c] = yr;
  }
}

// ---- both planes at once ("dual": the real plane inside the accumulator copy, the imaginary plane in a buffer of its own) ----
// With a plane per component a transpose needs ONE two-warp barrier (stores | loads) instead of three, and the stores of an
// element can be issued as soon as its inter-pass twiddle product exists: they drain while the FP64 pipe works on the next
// products (the store side of shared memory, 128 B/cycle/SM, is what the split transposes waited for).
// forward: fwd_twiddle_inplace + col_store_brev_al(re) + col_store_brev(im)
FB_HD void fwd_twiddle_col_store(double (&xr)[32], double (&xi)[32], const c2* tab_f, int lane, double* re_main_l, double* re_ovf_l, double* im_l) {
  c2 lo[4];
#pragma unroll
  for (int l = 0; l < 4; l++) lo[l] = tab_f[l * 32 + lane];
#pragma unroll
  for (int h = 0; h < 8; h++) {
    const c2 hi = tab_f[(4 + h) * 32 + lane];
#pragma unroll
    for (int l = 0; l < 4; l++) {
      const int q = brev5(4 * h + l);
      const int k1 = 4 * h + l;   // = brev5(q)
      double tr = lo[l].x, ti = lo[l].y;
      if (h != 0) {
        tr = fb_fma(lo[l].x, hi.x, -(lo[l].y * hi.y));
        ti = fb_fma(lo[l].x, hi.y, lo[l].y * hi.x);
      }
      const double yr = fb_fma(xr[q], tr, -(xi[q] * ti));
      const double yi = fb_fma(xr[q], ti, xi[q] * tr);
      if (k1 < kPlaneMainRows) re_main_l[k1 * kPlaneRow] = yr;
      else re_ovf_l[(k1 - kPlaneMainRows) * kPlaneRow] = yr;
      im_l[k1 * kPlaneRow] = yi;
    }
  }
}
// inverse: inv_twiddle_inplace + row_store(re) + row_store(im); rows as 128-bit pairs of columns
FB_HD void inv_twiddle_row_store(double (&xr)[32], double (&xi)[32], const c2* tab_i, int k1, double* re_row, double* im_row) {
  c2 lo[4];
#pragma unroll
  for (int l = 0; l < 4; l++) lo[l] = tab_i[l * 32 + k1];
  c2* rr = reinterpret_cast<c2*>(re_row);
  c2* ri = reinterpret_cast<c2*>(im_row);
#pragma unroll
  for (int h = 0; h < 8; h++) {
    const c2 hi = tab_i[(4 + h) * 32 + k1];
    double yr[4], yi[4];
#pragma unroll
    for (int l = 0; l < 4; l++) {
      const int c = 4 * h + l;
      double tr = lo[l].x, ti = lo[l].y;
      if (h != 0) {
        tr = fb_fma(lo[l].x, hi.x, -(lo[l].y * hi.y));
        ti = fb_fma(lo[l].x, hi.y, lo[l].y * hi.x);
      }
      yr[l] = fb_fma(xr[c], tr, -(xi[c] * ti));
      yi[l] = fb_fma(xr[c], ti, xi[c] * tr);
    }
#pragma unroll
    for (int u = 0; u < 2; u++) {
      c2 a, b;
      a.x = yr[2 * u]; a.y = yr[2 * u + 1];
      b.x = yi[2 * u]; b.y = yi[2 * u + 1];
      rr[2 * h + u] = a;
      ri[2 * h + u] = b;
    }
  }
}

// ---- fused CMUX body (blind_rotate_fused_kernel, br_fused.cu) ------------------------------------------
// The same arithmetic as above cut into pieces that interleave in program order: the decomposition of slot r
// (shared-memory reads + integer work) with the butterflies of pass 1 that its digits make computable; the
// Fourier MAC (shuffles + key reads) block by block between the last stages of forward pass 2 and the first
// stages of inverse pass 1; the torus rounding / accumulation with the last-stage butterflies of inverse pass 2.
// A warp then has FP64 work to issue while its shared-memory requests are in flight.

// digits of slot R: coefficients 32R+lane (re) and 32R+lane+1024 (im) of acc*X^a - acc.
// sm = base of the shared memory block, shp_off = byte offset of this polynomial's 8 KiB accumulator copy in it
// (a multiple of 8192, so the masked rotation offset is OR-ed in); t0 = 4*((lane - a) mod 4096);
// own = &acc[lane] (this thread's own words, 32 apart)
// the same digit through the integer-to-double conversion unit (one instruction instead of XOR + move + DADD)
FB_HD double pbs_digit32_cvt(uint32_t diff) {
  const int32_t q = (int32_t)(diff + 256u) >> 9;
#if defined(__CUDA_ARCH__)
  return __int2double_rn(q);
#else
  return (double)q;
#endif
}
template <int R, bool CVT = false>
FB_HD void phaseA_slot(double& dr, double& di, const unsigned char* sm, uint32_t shp_off, uint32_t t0, const uint32_t* own) {
  const uint32_t ta = t0 + 128u * R, tb = t0 + 128u * (R + 32);
#if defined(__CUDA_ARCH__)
  // the kernel places the block so that sm + shp_off is 8 KiB aligned as an ABSOLUTE shared-memory address: the masked rotation
  // offset is OR-ed into it (one LOP3 in front of the load instead of a LOP3 and an add)
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + shp_off;
  uint32_t xa = *reinterpret_cast<const uint32_t*>(__cvta_shared_to_generic((ta & 8188u) | base));
  uint32_t xb = *reinterpret_cast<const uint32_t*>(__cvta_shared_to_generic((tb & 8188u) | base));
#else
  uint32_t xa = *reinterpret_cast<const uint32_t*>(sm + ((ta & 8188u) | shp_off));
  uint32_t xb = *reinterpret_cast<const uint32_t*>(sm + ((tb & 8188u) | shp_off));
#endif
  if (ta & 8192u) xa = 0u - xa;
  if (tb & 8192u) xb = 0u - xb;
  if (CVT) {
    dr = pbs_digit32_cvt(xa - own[32 * R]);
    di = pbs_digit32_cvt(xb - own[32 * (R + 32)]);
  } else {
    dr = pbs_digit32(xa - own[32 * R]);
    di = pbs_digit32(xb - own[32 * (R + 32)]);
  }
}
template <int N, bool CVT>
FB_HD void phaseA_f1_step(double (&xr)[32], double (&xi)[32], const unsigned char* sm, uint32_t shp_off, uint32_t t0, const uint32_t* own) {
  constexpr int R = fb_f1_order(N);
  phaseA_slot<R, CVT>(xr[R], xi[R], sm, shp_off, t0, own);
  fft32_f1_step<N>(xr, xi);
}
template <bool CVT, int... Ns>
FB_HD void phaseA_f1_seq(double (&xr)[32], double (&xi)[32], const unsigned char* sm, uint32_t shp_off, uint32_t t0, const uint32_t* own,
                         fb_iseq<Ns...>) {
  (phaseA_f1_step<Ns, CVT>(xr, xi, sm, shp_off, t0, own), ...);
}
// phase A + forward pass 1 (twist folded in): replaces phaseA_load32 + fft32_fwd_twist
template <bool CVT = false>
FB_HD void phaseA_f1(double (&xr)[32], double (&xi)[32], const unsigned char* sm, uint32_t shp_off, uint32_t a, int lane) {
  const uint32_t t0 = (4u * (uint32_t)lane - 4u * a) & 16380u;
  const uint32_t* own = reinterpret_cast<const uint32_t*>(sm + shp_off) + lane;
  phaseA_f1_seq<CVT>(xr, xi, sm, shp_off, t0, own, fb_make_iseq<32>{});
}

// round(frac(k * ts) * 2^32) mod 2^32 in three FP64 instructions, the pending factor k of the untwist folded in.  Everything is
// done at the scale 2^32 so that ONE per-slot constant (k32 = k * 2^32) is needed:
//   s1 = ts * k32 + 1.5 * 2^84        the sum rounds to a multiple of 2^32: s1 = 1.5 * 2^84 + 2^32 rint(k ts)
//   g  = (1.5 * 2^84 + 1.5 * 2^52) - s1 = 1.5 * 2^52 - 2^32 rint(k ts)          (exact)
//   r  = ts * k32 + g = 2^32 (k ts - rint(k ts)) + 1.5 * 2^52                  (low word: the torus value); |k ts| < 2^51
FB_HD uint32_t torus32_round_scaled32(double ts, double k32) {
  const double kMp = 29014219670751100192948224.0;                         // 1.5 * 2^84
  const double kC = 29014219670751100192948224.0 + 6755399441055744.0;     // + 1.5 * 2^52 (exact: 34 significant bits)
  const double s1 = fb_fma(ts, k32, kMp);
  const double g = kC - s1;
  return double_lo32(fb_fma(ts, k32, g));
}
FB_HD uint32_t torus32_round_scaled(double ts, double k) { return torus32_round_scaled32(ts, k * 4294967296.0); }
// torus increments of slot R after fft32_i2_fin (pending untwist magnitudes and the 1/1024 of the transform); the 64 scale factors
// come from a constant-bank table (as literals each costs two 32-bit moves per CMUX step)
#define FB_PC1(R) fb_i2_kre(R) * (4294967296.0 / 1024.0), fb_i2_kim(R) * (4294967296.0 / 1024.0)
#define FB_PC4(R) FB_PC1(R), FB_PC1(R + 1), FB_PC1(R + 2), FB_PC1(R + 3)
#define FB_PCTAB_INIT { FB_PC4(0), FB_PC4(4), FB_PC4(8), FB_PC4(12), FB_PC4(16), FB_PC4(20), FB_PC4(24), FB_PC4(28) }
#if defined(__CUDACC__)
static __constant__ double fb_pctab_d[64] = FB_PCTAB_INIT;
#endif
static const double fb_pctab_h[64] = FB_PCTAB_INIT;
#if defined(__CUDA_ARCH__)
#define FBPC(i) fb_pctab_d[i]
#else
#define FBPC(i) fb_pctab_h[i]
#endif
template <int R>
FB_HD void phaseC_slot(const double (&xr)[32], const double (&xi)[32], uint32_t& inc0, uint32_t& inc1) {
  inc0 = torus32_round_scaled32(xr[R], FBPC(2 * R));
  inc1 = torus32_round_scaled32(xi[R], FBPC(2 * R + 1));
}

// Host-side: twiddle tables.  tab_f[e][lane], tab_i[e][k1], e < 4: "lo" l = e; e >= 4: "hi" h = e-4.
static inline void make_twiddle_tables(c2* tab_f, c2* tab_i) {
  const long double pi = 3.141592653589793238462643383279502884L;
  for (int x = 0; x < 32; x++) {
    for (int l = 0; l < 4; l++) {
      long double af = pi * (long double)x * (long double)(1 - 4 * l) / 2048.0L;   // lane c = x
      tab_f[l * 32 + x].x = (double)cosl(af);
      tab_f[l * 32 + x].y = (double)sinl(af);
      long double ai = -pi * (long double)l * (long double)(1 - 4 * x) / 2048.0L;  // row k1 = x
      tab_i[l * 32 + x].x = (double)cosl(ai);
      tab_i[l * 32 + x].y = (double)sinl(ai);
    }
    for (int h = 0; h < 8; h++) {
      long double af = -pi * (long double)x * (long double)(16 * h) / 2048.0L;
      tab_f[(4 + h) * 32 + x].x = (double)cosl(af);
      tab_f[(4 + h) * 32 + x].y = (double)sinl(af);
      long double ai = -pi * (long double)(4 * h) * (long double)(1 - 4 * x) / 2048.0L;
      tab_i[(4 + h) * 32 + x].x = (double)cosl(ai);
      tab_i[(4 + h) * 32 + x].y = (double)sinl(ai);
    }
  }
}

}  // namespace fb
