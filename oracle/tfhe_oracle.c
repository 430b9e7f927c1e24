/*
 * tfhe_oracle.c -- CPU restatement of the TFHE arithmetic under fhe-regex's hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (fhe_regex_b200/) links, loads or
 * calls this file.  It is used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs as the checker and as the timed CPU baseline ("port").
 *
 * PARITY STATUS
 *   - The arithmetic of the reference lives in an un-vendored dependency:
 *       tfhe 0.2.0 @ 13ad7d5468411c88f9f01e1cb175a2ac28ed2f72   (/root/reference/Cargo.lock:602-604)
 *       concrete-fft 0.1.0 (Cargo.lock:111), concrete-csprng 0.3.0 (Cargo.lock:100)
 *     whose sources are NOT in /root/reference and cannot be fetched or built here (no cargo/rustc,
 *     no network).  This file restates the PUBLISHED algorithms of that crate (signed gadget
 *     decomposition, LWE keyswitch, PBS modulus switch, CMUX blind rotation with an f64
 *     negacyclic FFT, sample extraction, shortint accumulator generation).
 *   - The reference holds NO known-answer vectors for keyswitch / bootstrap, so at
 *     the ciphertext level this oracle is "PARITY UNPINNED".  It is pinned only (a) against the
 *     parameter set and secret keys decoded from the reference fixture test_data/client_key,
 *     (b) by self-consistency (decrypt correctness, exact-integer vs FFT blind rotation, noise
 *     variance vs the parameter set's bound) and (c) at the regex level by the reference's own
 *     25 engine vectors / 49 parser vectors (oracle/regex_plain.py, tests/golden/).
 *
 * Call sites in the reference that this arithmetic sits under:
 *   execution.rs:76,93,110,143,173,190 (smart_eq/gt/le/bitand/bitor/bitxor),
 *   ciphertext.rs:26 (create_trivial), ciphertext.rs:38 (encrypt), mod.rs:17 (decrypt),
 *   ciphertext.rs:42-45 (gen_keys_radix(PARAM_MESSAGE_2_CARRY_2, 4)).
 *
 * Parameter set (decoded from test_data/client_key, offsets in SURVEY.md 8c):
 *   n=742, k=1, N=2048, pbs base 2^23 x 1 level, ks base 2^3 x 5 levels, q=2^64,
 *   message modulus 4, carry modulus 4 => delta = 2^59.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LWE_N 742
#define GLWE_N 2048
#define HALF_N 1024
#define KS_LEVELS 5
#define KS_BASE_LOG 3
#define PBS_BASE_LOG 23
#define BIG_LWE (GLWE_N + 1)
#define SMALL_LWE (LWE_N + 1)

typedef uint64_t u64;
typedef int64_t i64;

/* ------------------------------------------------------------------------------------------ */
/* Deterministic test RNG (xoshiro256** seeded through splitmix64).  The reference uses the   */
/* AES-CTR CSPRNG of concrete-csprng; only the distributions matter for a statistical oracle. */
/* ------------------------------------------------------------------------------------------ */
typedef struct { u64 s[4]; int have_spare; double spare; } orc_rng;

static u64 splitmix64(u64 *x) {
  u64 z = (*x += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static void rng_seed(orc_rng *r, u64 seed, u64 stream) {
  u64 x = seed * 0xD1342543DE82EF95ull + stream * 0x9E3779B97F4A7C15ull + 0x1234567ull;
  for (int i = 0; i < 4; i++) r->s[i] = splitmix64(&x);
  r->have_spare = 0;
}
static inline u64 rotl64(u64 x, int k) { return (x << k) | (x >> (64 - k)); }
static inline u64 rng_u64(orc_rng *r) {
  u64 *s = r->s;
  u64 result = rotl64(s[1] * 5, 7) * 9;
  u64 t = s[1] << 17;
  s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
  s[2] ^= t; s[3] = rotl64(s[3], 45);
  return result;
}
static inline double rng_unit(orc_rng *r) { return ((rng_u64(r) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
static double rng_gauss(orc_rng *r) {
  if (r->have_spare) { r->have_spare = 0; return r->spare; }
  double u1 = rng_unit(r), u2 = rng_unit(r);
  double m = sqrt(-2.0 * log(u1));
  r->spare = m * sin(2.0 * M_PI * u2);
  r->have_spare = 1;
  return m * cos(2.0 * M_PI * u2);
}
/* torus Gaussian with standard deviation sigma (as a fraction of the torus) */
static inline u64 rng_torus_gauss(orc_rng *r, double sigma) {
  double e = rng_gauss(r) * sigma * 18446744073709551616.0;
  return (u64)(i64)llround(e);
}

/* ------------------------------------------------------------------------------------------ */
/* T2: signed gadget decomposition (tfhe-rs SignedDecomposer / decompose_one_level).           */
/* ------------------------------------------------------------------------------------------ */
/* closest representable: round x to the top base_log*levels bits (round half up on the bit below). */
static inline u64 closest_representable(u64 x, int base_log, int levels) {
  int non_rep = 64 - base_log * levels;
  u64 r = x >> (non_rep - 1);
  r += 1;
  r >>= 1;
  return r << non_rep;
}
/* one level of the balanced decomposition; digits come out least-significant first */
static inline i64 decompose_one_level(int base_log, u64 *state, u64 mod_b_mask) {
  u64 res = *state & mod_b_mask;
  *state >>= base_log;
  u64 carry = ((res - 1ull) | *state) & res;
  carry >>= (base_log - 1);
  *state += carry;
  return (i64)(res - (carry << base_log));
}
/* exported for known-answer tests: digits[levels], least-significant (highest level index) first */
void orc_decompose(u64 x, int base_log, int levels, i64 *digits) {
  u64 state = closest_representable(x, base_log, levels) >> (64 - base_log * levels);
  u64 mask = (1ull << base_log) - 1ull;
  for (int l = 0; l < levels; l++) digits[l] = decompose_one_level(base_log, &state, mask);
}
u64 orc_closest_representable(u64 x, int base_log, int levels) { return closest_representable(x, base_log, levels); }

/* ------------------------------------------------------------------------------------------ */
/* T1 / T8: LWE encryption, phase; server key generation from the fixture's secret keys.       */
/* ------------------------------------------------------------------------------------------ */
/* out[dim+1]: mask uniform, body = <a,s> + pt + e */
void orc_lwe_encrypt(const u64 *key, int dim, u64 pt, double sigma, u64 seed, u64 stream, u64 *out) {
  orc_rng r; rng_seed(&r, seed, stream);
  u64 acc = 0;
  for (int i = 0; i < dim; i++) { u64 a = rng_u64(&r); out[i] = a; acc += a * key[i]; }
  out[dim] = acc + pt + rng_torus_gauss(&r, sigma);
}
void orc_lwe_trivial(int dim, u64 pt, u64 *out) { memset(out, 0, sizeof(u64) * dim); out[dim] = pt; }
u64 orc_lwe_phase(const u64 *key, int dim, const u64 *ct) {
  u64 acc = 0;
  for (int i = 0; i < dim; i++) acc += ct[i] * key[i];
  return ct[dim] - acc;
}
/* shortint decode: round at bit 58, >> 59 (message and carry, 4 bits) */
u64 orc_decode(u64 phase) { return ((phase + (1ull << 58)) >> 59) & 15ull; }

/* KSK[i][l][743], l = 0 is level 1 (most significant, factor 2^(64-3)), encrypted under small key. */
void orc_keygen_ksk(const u64 *big_key, const u64 *small_key, double sigma, u64 seed, u64 *ksk) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < GLWE_N; i++) {
    for (int l = 0; l < KS_LEVELS; l++) {
      u64 pt = big_key[i] << (64 - KS_BASE_LOG * (l + 1));
      orc_lwe_encrypt(small_key, LWE_N, pt, sigma, seed, 0x100000ull + (u64)i * KS_LEVELS + l,
                      ksk + ((size_t)i * KS_LEVELS + l) * SMALL_LWE);
    }
  }
}

/* out += a (*) s  negacyclic, s binary */
static void negacyclic_mul_binary_add(const u64 *a, const u64 *s, u64 *out) {
  for (int t = 0; t < GLWE_N; t++) {
    if (!s[t]) continue;
    for (int j = 0; j < GLWE_N - t; j++) out[j + t] += a[j];
    for (int j = GLWE_N - t; j < GLWE_N; j++) out[j + t - GLWE_N] -= a[j];
  }
}
/* GLWE (k=1) encryption of plaintext polynomial pt: ct = (A, B = A*S + E + pt) */
static void glwe_encrypt(const u64 *glwe_key, const u64 *pt, double sigma, orc_rng *r, u64 *ct) {
  u64 *A = ct, *B = ct + GLWE_N;
  for (int j = 0; j < GLWE_N; j++) A[j] = rng_u64(r);
  for (int j = 0; j < GLWE_N; j++) B[j] = pt[j] + rng_torus_gauss(r, sigma);
  negacyclic_mul_binary_add(A, glwe_key, B);
}
/* BSK standard domain: [i < 742][level = 1][row < 2][poly < 2][2048].
 * Row 0 encrypts -(s_i * 2^41) * S(X) (mask row), row 1 encrypts s_i * 2^41 (body row). */
void orc_keygen_bsk(const u64 *small_key, const u64 *glwe_key, double sigma, u64 seed, u64 *bsk) {
#pragma omp parallel for schedule(dynamic, 4)
  for (int i = 0; i < LWE_N; i++) {
    orc_rng r; rng_seed(&r, seed, 0x200000ull + (u64)i);
    u64 factor = small_key[i] << (64 - PBS_BASE_LOG);
    u64 pt[GLWE_N];
    for (int j = 0; j < GLWE_N; j++) pt[j] = (u64)0 - factor * glwe_key[j];
    glwe_encrypt(glwe_key, pt, sigma, &r, bsk + ((size_t)i * 2 + 0) * 2 * GLWE_N);
    memset(pt, 0, sizeof(pt));
    pt[0] = factor;
    glwe_encrypt(glwe_key, pt, sigma, &r, bsk + ((size_t)i * 2 + 1) * 2 * GLWE_N);
  }
}

/* ------------------------------------------------------------------------------------------ */
/* T2: LWE keyswitch, exact mod 2^64.                                                          */
/* ------------------------------------------------------------------------------------------ */
void orc_keyswitch(const u64 *ksk, const u64 *in, u64 *out) {
  memset(out, 0, sizeof(u64) * SMALL_LWE);
  out[LWE_N] = in[GLWE_N];
  const u64 mask = (1ull << KS_BASE_LOG) - 1ull;
  for (int i = 0; i < GLWE_N; i++) {
    u64 state = closest_representable(in[i], KS_BASE_LOG, KS_LEVELS) >> (64 - KS_BASE_LOG * KS_LEVELS);
    /* first digit out is the least significant one = level KS_LEVELS = last stored row */
    for (int l = KS_LEVELS - 1; l >= 0; l--) {
      u64 d = (u64)decompose_one_level(KS_BASE_LOG, &state, mask);
      if (!d) continue;
      const u64 *row = ksk + ((size_t)i * KS_LEVELS + l) * SMALL_LWE;
      for (int c = 0; c < SMALL_LWE; c++) out[c] -= d * row[c];
    }
  }
}
void orc_keyswitch_batch(const u64 *ksk, const u64 *in, u64 *out, int count, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < count; b++) orc_keyswitch(ksk, in + (size_t)b * BIG_LWE, out + (size_t)b * SMALL_LWE);
}

/* ------------------------------------------------------------------------------------------ */
/* T3: PBS modulus switch to Z_{2N}: round x * 2N / 2^64, result in [0, 2N].                   */
/* ------------------------------------------------------------------------------------------ */
u64 orc_modswitch(u64 x) {
  u64 t = x >> 51; /* 64 - log2(2N) - 1 */
  t += t & 1ull;
  return t >> 1;
}

/* ------------------------------------------------------------------------------------------ */
/* T4: shortint accumulator: body[i*128..(i+1)*128) = f(i)*delta; negate first 64; rotate left 64. */
/* ------------------------------------------------------------------------------------------ */
void orc_make_lut(const u64 *f16, u64 *lut) {
  u64 tmp[GLWE_N];
  const int box = GLWE_N / 16, half = box / 2;
  for (int i = 0; i < 16; i++)
    for (int j = 0; j < box; j++) tmp[i * box + j] = f16[i] << 59;
  for (int j = 0; j < half; j++) tmp[j] = (u64)0 - tmp[j];
  for (int j = 0; j < GLWE_N; j++) lut[j] = tmp[(j + half) % GLWE_N];
}

/* ------------------------------------------------------------------------------------------ */
/* Negacyclic FFT: fold N reals into N/2 complex, twist by exp(i*pi*j/N), radix-2 FFT.         */
/* Frequency order is the bit-reversed order produced by the DIF network (internal to oracle). */
/* ------------------------------------------------------------------------------------------ */
static double tw_re[HALF_N], tw_im[HALF_N];        /* twist: exp(i*pi*j/N), j < N/2 */
static double w_re[HALF_N / 2], w_im[HALF_N / 2];  /* exp(-2*pi*i*j/(N/2)), j < N/4 */
static int fft_ready = 0;
static void fft_init(void) {
  if (fft_ready) return;
#pragma omp critical
  {
    if (!fft_ready) {
      for (int j = 0; j < HALF_N; j++) {
        long double a = M_PIl * (long double)j / (long double)GLWE_N;
        tw_re[j] = (double)cosl(a); tw_im[j] = (double)sinl(a);
      }
      for (int j = 0; j < HALF_N / 2; j++) {
        long double a = -2.0L * M_PIl * (long double)j / (long double)HALF_N;
        w_re[j] = (double)cosl(a); w_im[j] = (double)sinl(a);
      }
      fft_ready = 1;
    }
  }
}
/* in-place DIF, natural in -> bit-reversed out */
static void fft_dif(double *re, double *im) {
  for (int half = HALF_N / 2, step = 1; half >= 1; half >>= 1, step <<= 1) {
    for (int g = 0; g < HALF_N; g += 2 * half) {
      for (int j = 0; j < half; j++) {
        int a = g + j, b = a + half;
        double wr = w_re[j * step], wi = w_im[j * step];
        double xr = re[a] - re[b], xi = im[a] - im[b];
        re[a] += re[b]; im[a] += im[b];
        re[b] = xr * wr - xi * wi;
        im[b] = xr * wi + xi * wr;
      }
    }
  }
}
/* in-place DIT with conjugate twiddles, bit-reversed in -> natural out (unnormalised inverse) */
static void ifft_dit(double *re, double *im) {
  for (int half = 1, step = HALF_N / 2; half < HALF_N; half <<= 1, step >>= 1) {
    for (int g = 0; g < HALF_N; g += 2 * half) {
      for (int j = 0; j < half; j++) {
        int a = g + j, b = a + half;
        double wr = w_re[j * step], wi = -w_im[j * step];
        double xr = re[b] * wr - im[b] * wi;
        double xi = re[b] * wi + im[b] * wr;
        re[b] = re[a] - xr; im[b] = im[a] - xi;
        re[a] += xr; im[a] += xi;
      }
    }
  }
}
/* forward of a real polynomial given as doubles p[2048] */
static void nfft_forward(const double *p, double *re, double *im) {
  for (int j = 0; j < HALF_N; j++) {
    double a = p[j], b = p[j + HALF_N];
    re[j] = a * tw_re[j] - b * tw_im[j];
    im[j] = a * tw_im[j] + b * tw_re[j];
  }
  fft_dif(re, im);
}
/* backward into real coefficients (unrounded doubles) */
static void nfft_backward(double *re, double *im, double *p) {
  ifft_dit(re, im);
  const double inv = 1.0 / HALF_N;
  for (int j = 0; j < HALF_N; j++) {
    double a = re[j] * inv, b = im[j] * inv;
    p[j] = a * tw_re[j] + b * tw_im[j];
    p[j + HALF_N] = b * tw_re[j] - a * tw_im[j];
  }
}
/* tfhe-rs from_torus: fractional part of t, times 2^64, rounded, as a wrapping u64 */
static inline u64 from_torus(double t) {
  double f = t - nearbyint(t);
  double v = nearbyint(f * 18446744073709551616.0);
  if (v >= 9223372036854775808.0) return 0x8000000000000000ull;
  if (v < -9223372036854775808.0) return 0x8000000000000000ull;
  return (u64)(i64)v;
}

/* K7: standard BSK -> Fourier BSK.  Layout out[i][row][poly][2][1024] doubles (re block, im block),
 * coefficients read as signed torus in [-1/2, 1/2) (forward_as_torus). */
void orc_bsk_to_fourier(const u64 *bsk, double *fbsk) {
  fft_init();
#pragma omp parallel for schedule(static)
  for (int t = 0; t < LWE_N * 4; t++) {
    const u64 *src = bsk + (size_t)t * GLWE_N;
    double p[GLWE_N];
    for (int j = 0; j < GLWE_N; j++) p[j] = (double)(i64)src[j] * (1.0 / 18446744073709551616.0);
    nfft_forward(p, fbsk + (size_t)t * GLWE_N, fbsk + (size_t)t * GLWE_N + HALF_N);
  }
}

/* multiply by X^a (0 <= a < 2N), negacyclic */
static void poly_rotate(const u64 *in, unsigned a, u64 *out) {
  for (unsigned j = 0; j < GLWE_N; j++) {
    unsigned idx = (j + 2 * GLWE_N - a) & (2 * GLWE_N - 1);
    out[j] = idx < GLWE_N ? in[idx] : (u64)0 - in[idx - GLWE_N];
  }
}
static inline double pbs_digit(u64 x) {
  u64 state = closest_representable(x, PBS_BASE_LOG, 1) >> (64 - PBS_BASE_LOG);
  return (double)decompose_one_level(PBS_BASE_LOG, &state, (1ull << PBS_BASE_LOG) - 1ull);
}

/* T5: blind rotation, f64 FFT external products.  acc[2][2048] out (mask poly, body poly). */
void orc_blind_rotate_fft(const double *fbsk, const u64 *lwe, const u64 *lut, u64 *acc) {
  fft_init();
  u64 *am = acc, *ab = acc + GLWE_N;
  unsigned bt = (unsigned)orc_modswitch(lwe[LWE_N]);
  memset(am, 0, sizeof(u64) * GLWE_N);
  /* acc = lut * X^{-b} = lut * X^{2N - b} */
  poly_rotate(lut, (2 * GLWE_N - bt) & (2 * GLWE_N - 1), ab);
  u64 rot[GLWE_N];
  double d[GLWE_N], f_re[2][HALF_N], f_im[2][HALF_N], o_re[HALF_N], o_im[HALF_N], p[GLWE_N];
  for (int i = 0; i < LWE_N; i++) {
    if (lwe[i] == 0) continue;
    unsigned at = (unsigned)orc_modswitch(lwe[i]) & (2 * GLWE_N - 1);
    for (int r = 0; r < 2; r++) {
      u64 *poly = acc + r * GLWE_N;
      poly_rotate(poly, at, rot);
      for (int j = 0; j < GLWE_N; j++) d[j] = pbs_digit(rot[j] - poly[j]);
      nfft_forward(d, f_re[r], f_im[r]);
    }
    const double *g = fbsk + (size_t)i * 4 * GLWE_N; /* [row][poly][re|im][1024] */
    for (int jp = 0; jp < 2; jp++) {
      const double *g0r = g + (0 * 2 + jp) * GLWE_N, *g0i = g0r + HALF_N;
      const double *g1r = g + (1 * 2 + jp) * GLWE_N, *g1i = g1r + HALF_N;
      for (int k = 0; k < HALF_N; k++) {
        o_re[k] = f_re[0][k] * g0r[k] - f_im[0][k] * g0i[k] + f_re[1][k] * g1r[k] - f_im[1][k] * g1i[k];
        o_im[k] = f_re[0][k] * g0i[k] + f_im[0][k] * g0r[k] + f_re[1][k] * g1i[k] + f_im[1][k] * g1r[k];
      }
      nfft_backward(o_re, o_im, p);
      u64 *poly = acc + jp * GLWE_N;
      for (int j = 0; j < GLWE_N; j++) poly[j] += from_torus(p[j]);
    }
  }
}

/* exact negacyclic product accumulate: out += d (*) g  mod 2^64, d small signed */
static void negacyclic_mul_exact_add(const i64 *d, const u64 *g, u64 *out) {
  for (int t = 0; t < GLWE_N; t++) {
    u64 dt = (u64)d[t];
    if (!dt) continue;
    for (int j = 0; j < GLWE_N - t; j++) out[j + t] += dt * g[j];
    for (int j = GLWE_N - t; j < GLWE_N; j++) out[j + t - GLWE_N] -= dt * g[j];
  }
}
/* T5 (differential): same blind rotation with exact integer external products (standard-domain BSK) */
void orc_blind_rotate_exact(const u64 *bsk, const u64 *lwe, const u64 *lut, u64 *acc) {
  u64 *am = acc, *ab = acc + GLWE_N;
  unsigned bt = (unsigned)orc_modswitch(lwe[LWE_N]);
  memset(am, 0, sizeof(u64) * GLWE_N);
  poly_rotate(lut, (2 * GLWE_N - bt) & (2 * GLWE_N - 1), ab);
  u64 rot[GLWE_N];
  static __thread i64 d[2][GLWE_N];
  for (int i = 0; i < LWE_N; i++) {
    if (lwe[i] == 0) continue;
    unsigned at = (unsigned)orc_modswitch(lwe[i]) & (2 * GLWE_N - 1);
    for (int r = 0; r < 2; r++) {
      u64 *poly = acc + r * GLWE_N;
      poly_rotate(poly, at, rot);
      for (int j = 0; j < GLWE_N; j++) d[r][j] = (i64)pbs_digit(rot[j] - poly[j]);
    }
    const u64 *g = bsk + (size_t)i * 4 * GLWE_N;
    for (int jp = 0; jp < 2; jp++)
      for (int r = 0; r < 2; r++) negacyclic_mul_exact_add(d[r], g + (r * 2 + jp) * GLWE_N, acc + jp * GLWE_N);
  }
}

/* T6: sample extract of the constant coefficient. */
void orc_sample_extract(const u64 *acc, u64 *out) {
  const u64 *am = acc, *ab = acc + GLWE_N;
  out[0] = am[0];
  for (int j = 1; j < GLWE_N; j++) out[j] = (u64)0 - am[GLWE_N - j];
  out[GLWE_N] = ab[0];
}

/* KS -> PBS, the shortint keyswitch_programmable_bootstrap order of tfhe-rs 0.2 */
void orc_pbs(const u64 *ksk, const double *fbsk, const u64 *in, const u64 *lut, u64 *out) {
  u64 small[SMALL_LWE];
  u64 *acc = (u64 *)malloc(sizeof(u64) * 2 * GLWE_N);
  orc_keyswitch(ksk, in, small);
  orc_blind_rotate_fft(fbsk, small, lut, acc);
  orc_sample_extract(acc, out);
  free(acc);
}
/* luts[n_luts][2048]; lut_idx[count] */
void orc_pbs_batch(const u64 *ksk, const double *fbsk, const u64 *in, const u64 *luts, const uint32_t *lut_idx,
                   u64 *out, int count, int nthreads) {
  fft_init();
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < count; b++)
    orc_pbs(ksk, fbsk, in + (size_t)b * BIG_LWE, luts + (size_t)lut_idx[b] * GLWE_N, out + (size_t)b * BIG_LWE);
}
/* blind rotate + sample extract on an already keyswitched input (for stage-wise parity) */
void orc_bootstrap_small(const double *fbsk, const u64 *small, const u64 *lut, u64 *out) {
  u64 *acc = (u64 *)malloc(sizeof(u64) * 2 * GLWE_N);
  orc_blind_rotate_fft(fbsk, small, lut, acc);
  orc_sample_extract(acc, out);
  free(acc);
}
int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* test helper: out = round(a (*) b) negacyclic through the oracle's own f64 FFT; a small signed ints, b torus */
void orc_negacyclic_mul_fft(const i64 *a, const u64 *b, u64 *out) {
  fft_init();
  double pa[GLWE_N], pb[GLWE_N], ar[HALF_N], ai[HALF_N], br[HALF_N], bi[HALF_N], p[GLWE_N];
  for (int j = 0; j < GLWE_N; j++) { pa[j] = (double)a[j]; pb[j] = (double)(i64)b[j] * (1.0 / 18446744073709551616.0); }
  nfft_forward(pa, ar, ai);
  nfft_forward(pb, br, bi);
  for (int k = 0; k < HALF_N; k++) {
    double r = ar[k] * br[k] - ai[k] * bi[k], im = ar[k] * bi[k] + ai[k] * br[k];
    ar[k] = r; ai[k] = im;
  }
  nfft_backward(ar, ai, p);
  for (int j = 0; j < GLWE_N; j++) out[j] = from_torus(p[j]);
}
