// client.cpp -- client-side glue of the C ABI (include/fhe_b200.h): key import, test keygen,
// encrypt / decrypt, accumulator generation.  Host only; NOT on the server hot path.
//
// Mirrors the reference's use of tfhe-rs on the client side:
//   gen_keys / ServerKey::new      ciphertext.rs:42-45, engine.rs:252
//   encrypt_str / create_trivial   ciphertext.rs:8-40
//   RadixClientKey::decrypt        mod.rs:17, engine.rs:289
//   bincode RadixClientKey         engine.rs:238-254 (fixture test_data/client_key)
// The PRNG is a seeded xoshiro256** (tests/bench need reproducible keys); tfhe-rs uses an AES-CTR
// CSPRNG -- keys made here are statistically equivalent to, not bit-identical with, tfhe-rs keys.
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>
#include "../../include/fhe_b200.h"

namespace {

constexpr int kLweN = 742, kN = 2048, kBig = 2049, kSmall = 743, kKsLevels = 5, kKsBaseLog = 3, kPbsBaseLog = 23;
constexpr double kSigmaLwe = 7.069849454709433e-06;   // lwe_modular_std_dev of PARAM_MESSAGE_2_CARRY_2 (fixture offset 38760)
constexpr double kSigmaGlwe = 2.9403601535432533e-16;  // glwe_modular_std_dev (fixture offset 38768)

struct Rng {
  uint64_t s[4];
  bool spare_ok = false;
  double spare = 0;
  static uint64_t mix(uint64_t& x) {
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  Rng(uint64_t seed, uint64_t stream) {
    uint64_t x = seed * 0xA24BAED4963EE407ull + stream * 0x9FB21C651E98DF25ull + 0x5851F42Dull;
    for (auto& v : s) v = mix(x);
  }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
  }
  double unit() { return ((next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
  double gauss() {
    if (spare_ok) { spare_ok = false; return spare; }
    double u1 = unit(), u2 = unit(), m = std::sqrt(-2.0 * std::log(u1));
    spare = m * std::sin(6.283185307179586476925 * u2);
    spare_ok = true;
    return m * std::cos(6.283185307179586476925 * u2);
  }
  uint64_t torus_gauss(double sigma) { return (uint64_t)(int64_t)std::llround(gauss() * sigma * 18446744073709551616.0); }
};

void lwe_encrypt(const uint64_t* key, int dim, uint64_t pt, double sigma, Rng& r, uint64_t* out) {
  uint64_t acc = 0;
  for (int i = 0; i < dim; i++) {
    uint64_t a = r.next();
    out[i] = a;
    acc += a * key[i];
  }
  out[dim] = acc + pt + r.torus_gauss(sigma);
}

// B += A (*) S negacyclic for a binary key S
void add_mul_binary(const uint64_t* A, const uint64_t* S, uint64_t* B) {
  for (int t = 0; t < kN; t++) {
    if (!S[t]) continue;
    for (int j = 0; j < kN - t; j++) B[j + t] += A[j];
    for (int j = kN - t; j < kN; j++) B[j + t - kN] -= A[j];
  }
}

template <class F>
void parallel_for(int n, F f) {
  unsigned hw = std::thread::hardware_concurrency();
  int nt = (int)(hw ? hw : 4);
  if (nt > 32) nt = 32;
  std::vector<std::thread> th;
  for (int t = 0; t < nt; t++)
    th.emplace_back([=]() {
      for (int i = t; i < n; i += nt) f(i);
    });
  for (auto& x : th) x.join();
}

}  // namespace

extern "C" int fb_client_key_from_bincode(const uint8_t* buf, size_t len, uint64_t* big_key, uint64_t* small_key) {
  // layout measured on the reference fixture: Vec<u64> big (len 2048), Vec<u64> glwe (2048),
  // polynomial_size, Vec<u64> small (742), 16 x 8 B parameters, num_blocks
  if (!buf || !big_key || !small_key) return FB_ERR_ARG;
  size_t o = 0;
  auto rd = [&](uint64_t& v) -> bool {
    if (o + 8 > len) return false;
    std::memcpy(&v, buf + o, 8);
    o += 8;
    return true;
  };
  uint64_t n = 0;
  if (!rd(n) || n != (uint64_t)kN || o + 8 * n > len) return FB_ERR_FORMAT;
  std::memcpy(big_key, buf + o, 8 * n);
  o += 8 * n;
  if (!rd(n) || n != (uint64_t)kN || o + 8 * n > len) return FB_ERR_FORMAT;
  if (std::memcmp(big_key, buf + o, 8 * n) != 0) return FB_ERR_FORMAT;  // GLWE key must equal the big LWE key (k = 1)
  o += 8 * n;
  uint64_t poly = 0;
  if (!rd(poly) || poly != (uint64_t)kN) return FB_ERR_FORMAT;
  if (!rd(n) || n != (uint64_t)kLweN || o + 8 * n > len) return FB_ERR_FORMAT;
  std::memcpy(small_key, buf + o, 8 * n);
  o += 8 * n;
  uint64_t p[16];
  for (auto& v : p)
    if (!rd(v)) return FB_ERR_FORMAT;
  uint64_t blocks = 0;
  if (!rd(blocks) || o != len) return FB_ERR_FORMAT;
  // lwe n, glwe k, N, pbs base/level, ks base/level, message/carry modulus, blocks
  if (p[0] != 742 || p[1] != 1 || p[2] != 2048 || p[5] != 23 || p[6] != 1 || p[7] != 3 || p[8] != 5 || p[14] != 4 ||
      p[15] != 4 || blocks != 4)
    return FB_ERR_FORMAT;
  for (int i = 0; i < kN; i++)
    if (big_key[i] > 1) return FB_ERR_FORMAT;
  for (int i = 0; i < kLweN; i++)
    if (small_key[i] > 1) return FB_ERR_FORMAT;
  return FB_OK;
}

extern "C" int fb_client_keygen_server(const uint64_t* big_key, const uint64_t* small_key, uint64_t seed, uint64_t* h_ksk,
                                       uint64_t* h_bsk_std) {
  if (!big_key || !small_key || !h_ksk || !h_bsk_std) return FB_ERR_ARG;
  // KSK: bit i of the big key times 2^(64 - 3*level), level = 1..5, under the small key
  parallel_for(kN, [&](int i) {
    for (int l = 0; l < kKsLevels; l++) {
      Rng r(seed, 0x10000000ull + (uint64_t)i * kKsLevels + l);
      lwe_encrypt(small_key, kLweN, big_key[i] << (64 - kKsBaseLog * (l + 1)), kSigmaLwe, r,
                  h_ksk + ((size_t)i * kKsLevels + l) * kSmall);
    }
  });
  // BSK: GGSW(s_i), one level: row 0 encrypts -(s_i 2^41) S(X), row 1 encrypts s_i 2^41
  parallel_for(kLweN, [&](int i) {
    Rng r(seed, 0x20000000ull + (uint64_t)i);
    const uint64_t factor = small_key[i] << (64 - kPbsBaseLog);
    for (int row = 0; row < 2; row++) {
      uint64_t* A = h_bsk_std + (((size_t)i * 2 + row) * 2 + 0) * kN;
      uint64_t* B = h_bsk_std + (((size_t)i * 2 + row) * 2 + 1) * kN;
      for (int j = 0; j < kN; j++) A[j] = r.next();
      for (int j = 0; j < kN; j++) {
        uint64_t pt = row == 0 ? (uint64_t)0 - factor * big_key[j] : (j == 0 ? factor : 0);
        B[j] = pt + r.torus_gauss(kSigmaGlwe);
      }
      add_mul_binary(A, big_key, B);
    }
  });
  return FB_OK;
}

extern "C" int fb_client_encrypt_block(const uint64_t* big_key, uint64_t m, uint64_t seed, uint64_t stream, uint64_t* h_out) {
  if (!big_key || !h_out) return FB_ERR_ARG;
  Rng r(seed, 0x30000000ull + stream);
  lwe_encrypt(big_key, kN, (m & 15ull) << 59, kSigmaGlwe, r, h_out);
  return FB_OK;
}

extern "C" int fb_client_encrypt_str(const uint64_t* big_key, const uint8_t* bytes, size_t n, uint64_t seed, uint64_t* h_out) {
  if (!big_key || (!bytes && n) || (!h_out && n)) return FB_ERR_ARG;
  for (size_t i = 0; i < n; i++)
    if (bytes[i] >= 128) return FB_ERR_ARG;  // "content contains non-ascii characters" (ciphertext.rs:33-35)
  for (size_t i = 0; i < n; i++)
    for (int b = 0; b < 4; b++)
      fb_client_encrypt_block(big_key, (bytes[i] >> (2 * b)) & 3, seed, i * 4 + b, h_out + (i * 4 + b) * kBig);
  return FB_OK;
}

extern "C" int fb_client_trivial_str(const uint8_t* bytes, size_t n, uint64_t* h_out) {
  if ((!bytes && n) || (!h_out && n)) return FB_ERR_ARG;
  std::memset(h_out, 0, n * 4 * kBig * sizeof(uint64_t));
  for (size_t i = 0; i < n; i++)
    for (int b = 0; b < 4; b++) h_out[(i * 4 + b) * kBig + kN] = (uint64_t)((bytes[i] >> (2 * b)) & 3) << 59;
  return FB_OK;
}

extern "C" uint64_t fb_client_phase(const uint64_t* key, size_t dim, const uint64_t* ct) {
  uint64_t acc = 0;
  for (size_t i = 0; i < dim; i++) acc += ct[i] * key[i];
  return ct[dim] - acc;
}

extern "C" uint64_t fb_client_decrypt_block(const uint64_t* big_key, const uint64_t* ct) {
  return ((fb_client_phase(big_key, kN, ct) + (1ull << 58)) >> 59) & 15ull;
}

extern "C" uint64_t fb_client_decrypt_radix(const uint64_t* big_key, const uint64_t* ct) {
  uint64_t v = 0, shift = 1;
  for (int b = 0; b < 4; b++) {
    v += fb_client_decrypt_block(big_key, ct + (size_t)b * kBig) * shift;
    shift *= 4;
  }
  return v % 256;
}

extern "C" int fb_make_lut(const uint64_t* f16, uint64_t* lut) {
  if (!f16 || !lut) return FB_ERR_ARG;
  const int box = kN / 16, half = box / 2;
  for (int j = 0; j < kN; j++) {
    const int src = (j + half) % kN;          // rotate left by half a box
    uint64_t v = (f16[src / box] & 15ull) << 59;
    if (src < half) v = (uint64_t)0 - v;      // the first half box was negated before the rotation
    lut[j] = v;
  }
  return FB_OK;
}
