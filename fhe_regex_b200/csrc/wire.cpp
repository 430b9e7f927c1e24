// wire.cpp -- tfhe-rs 0.2.0 wire formats at the drop-in boundary (host only): bincode readers / writers for
//   tfhe::integer::ServerKey        what `ServerKey::new(&client_key)` / `gen_keys_radix` produce and the reference hands to
//                                   has_match (/root/reference/src/regex/engine.rs:8-12, :20, :248-254; ciphertext.rs:42-45)
//   tfhe::integer::RadixCiphertext  one encrypted character (ciphertext.rs:6, :29, :32-40), and Vec<RadixCiphertext> =
//                                   StringCiphertext (ciphertext.rs:6; src/regex/mod.rs:13-17)
// The reference serializes with bincode 1.3.3 (Cargo.lock; engine.rs:242-251 for the client key fixture): default
// options = little endian, fixed-width integers, u64 sequence lengths, struct fields in declaration order, no framing.
//
// [UPSTREAM-MEMORY] tfhe-rs is not in /root/reference (git dependency tfhe 0.2.0 @13ad7d5, Cargo.lock:602-604), so the
// struct layouts below are restated from its published sources; every length and parameter is checked against the
// PARAM_MESSAGE_2_CARRY_2 constants and a mismatch is FB_ERR_FORMAT, never a guess.  Layouts:
//
//   integer::ServerKey { key: shortint::ServerKey }                                  (newtype: same bytes)
//   shortint::ServerKey {
//     key_switching_key: LweKeyswitchKey<Vec<u64>> {
//       data: Vec<u64>                  u64 len = 2048*5*743, words [input bit][level, most significant first][743]
//       decomp_base_log: usize          u64 = 3
//       decomp_level_count: usize       u64 = 5
//       output_lwe_size: usize          u64 = 743 }
//     bootstrapping_key: FourierLweBootstrapKey<ABox<[c64]>> {
//       fourier: FourierPolynomialList  custom Serialize: seq of 2 + count elements = u64 (2 + count), polynomial_size
//                                       u64 = 2048, count u64 = 742*1*2*2 = 2968, then per polynomial a seq: u64 len =
//                                       1024, 1024 x (f64 re, f64 im) in the PLAN-INDEPENDENT standard frequency order
//                                       (concrete-fft Plan::serialize_fourier_buffer) -- in memory tfhe-rs keeps the
//                                       plan's own permuted order, which depends on the machine the plan was measured
//                                       on, so the serialized order is the only portable one; it is the natural order
//                                       k = 0..1023 of X[k] = sum_j (a_j + i a_{j+1024}) e^{i pi j/2048} e^{-2 pi i jk/1024},
//                                       key coefficients read as signed torus fractions in [-1/2, 1/2): exactly the
//                                       layout of this library's resident Fourier key
//       input_lwe_dimension: usize      u64 = 742
//       glwe_size: usize                u64 = 2
//       decomposition_base_log: usize   u64 = 23
//       decomposition_level_count: usize u64 = 1 }
//     message_modulus: usize u64 = 4, carry_modulus: usize u64 = 4, max_degree: usize u64 = 15 }
//
//   RadixCiphertext { blocks: Vec<shortint::Ciphertext> }          u64 len = 4, then per block
//   shortint::Ciphertext { ct: LweCiphertext<Vec<u64>> { data: Vec<u64> }   u64 len = 2049, 2049 words (mask, body)
//                          degree: Degree(usize)  u64,  message_modulus: u64 = 4,  carry_modulus: u64 = 4 }
#include <cmath>
#include <cstring>
#include "../../include/fhe_b200.h"

namespace {

constexpr uint64_t kLweN = 742, kN = 2048, kHalfN = 1024, kBig = 2049, kSmall = 743;
constexpr uint64_t kPolys = kLweN * 4;   // 742 GGSW x 1 level x 2 rows x 2 polynomials

struct Reader {
  const uint8_t* p;
  size_t len, o = 0;
  bool u64(uint64_t& v) {
    if (o + 8 > len) return false;
    std::memcpy(&v, p + o, 8);
    o += 8;
    return true;
  }
  bool expect(uint64_t want) {
    uint64_t v;
    return u64(v) && v == want;
  }
  bool bytes(void* dst, size_t n) {
    if (n > len - o || o > len) return false;
    std::memcpy(dst, p + o, n);
    o += n;
    return true;
  }
};

struct Writer {
  uint8_t* p;
  size_t cap, o = 0;
  void u64(uint64_t v) {
    if (p && o + 8 <= cap) std::memcpy(p + o, &v, 8);
    o += 8;
  }
  void bytes(const void* src, size_t n) {
    if (p && o + n <= cap) std::memcpy(p + o, src, n);
    o += n;
  }
};

void write_block(Writer& w, const uint64_t* lwe, uint64_t degree) {
  w.u64(kBig);
  w.bytes(lwe, kBig * 8);
  w.u64(degree);
  w.u64(4);
  w.u64(4);
}

int read_block(Reader& r, uint64_t* lwe, uint64_t* degree) {
  uint64_t d = 0;
  if (!r.expect(kBig) || !r.bytes(lwe, kBig * 8) || !r.u64(d) || !r.expect(4) || !r.expect(4)) return FB_ERR_FORMAT;
  if (d > 15) return FB_ERR_FORMAT;   // max_degree of PARAM_MESSAGE_2_CARRY_2
  if (degree) *degree = d;
  return FB_OK;
}

}  // namespace

extern "C" size_t fb_server_key_bincode_size(void) {
  return 8 + FB_KSK_WORDS * 8 + 3 * 8            // key_switching_key
         + 3 * 8 + kPolys * (8 + kHalfN * 16)    // fourier polynomial list
         + 4 * 8                                  // lwe dimension, glwe size, base log, level count
         + 3 * 8;                                 // message modulus, carry modulus, max degree
}

extern "C" int fb_server_key_from_bincode(const uint8_t* buf, size_t len, uint64_t* h_ksk, double* h_fbsk) {
  if (!buf || !h_ksk || !h_fbsk) return FB_ERR_ARG;
  if (len != fb_server_key_bincode_size()) return FB_ERR_FORMAT;
  Reader r{buf, len};
  if (!r.expect(FB_KSK_WORDS) || !r.bytes(h_ksk, FB_KSK_WORDS * 8)) return FB_ERR_FORMAT;
  if (!r.expect(3) || !r.expect(5) || !r.expect(kSmall)) return FB_ERR_FORMAT;
  if (!r.expect(2 + kPolys) || !r.expect(kN) || !r.expect(kPolys)) return FB_ERR_FORMAT;
  for (uint64_t t = 0; t < kPolys; t++)
    if (!r.expect(kHalfN) || !r.bytes(h_fbsk + t * kHalfN * 2, kHalfN * 16)) return FB_ERR_FORMAT;
  if (!r.expect(kLweN) || !r.expect(2) || !r.expect(23) || !r.expect(1)) return FB_ERR_FORMAT;
  if (!r.expect(4) || !r.expect(4) || !r.expect(15) || r.o != len) return FB_ERR_FORMAT;
  for (uint64_t t = 0; t < kPolys * kHalfN * 2; t++)
    if (!std::isfinite(h_fbsk[t])) return FB_ERR_FORMAT;
  return FB_OK;
}

extern "C" int fb_server_key_to_bincode(const uint64_t* h_ksk, const double* h_fbsk, uint8_t* out, size_t cap, size_t* written) {
  if (!h_ksk || !h_fbsk || !written) return FB_ERR_ARG;
  Writer w{out, cap};
  w.u64(FB_KSK_WORDS);
  w.bytes(h_ksk, FB_KSK_WORDS * 8);
  w.u64(3);
  w.u64(5);
  w.u64(kSmall);
  w.u64(2 + kPolys);
  w.u64(kN);
  w.u64(kPolys);
  for (uint64_t t = 0; t < kPolys; t++) {
    w.u64(kHalfN);
    w.bytes(h_fbsk + t * kHalfN * 2, kHalfN * 16);
  }
  w.u64(kLweN);
  w.u64(2);
  w.u64(23);
  w.u64(1);
  w.u64(4);
  w.u64(4);
  w.u64(15);
  *written = w.o;
  return (out && w.o <= cap) ? FB_OK : FB_ERR_ARG;   // out == NULL: size query
}

// ---- ciphertexts ---------------------------------------------------------------------------------------------

extern "C" size_t fb_radix_bincode_size(void) { return 8 + 4 * (8 + kBig * 8 + 3 * 8); }

extern "C" int fb_radix_from_bincode(const uint8_t* buf, size_t len, uint64_t* h_ct, uint64_t* degrees) {
  if (!buf || !h_ct) return FB_ERR_ARG;
  if (len != fb_radix_bincode_size()) return FB_ERR_FORMAT;
  Reader r{buf, len};
  if (!r.expect(FB_RADIX_BLOCKS)) return FB_ERR_FORMAT;
  for (int b = 0; b < FB_RADIX_BLOCKS; b++) {
    int rc = read_block(r, h_ct + (size_t)b * kBig, degrees ? degrees + b : nullptr);
    if (rc != FB_OK) return rc;
  }
  return r.o == len ? FB_OK : FB_ERR_FORMAT;
}

extern "C" int fb_radix_to_bincode(const uint64_t* h_ct, const uint64_t* degrees, uint8_t* out, size_t cap, size_t* written) {
  if (!h_ct || !written) return FB_ERR_ARG;
  Writer w{out, cap};
  w.u64(FB_RADIX_BLOCKS);
  for (int b = 0; b < FB_RADIX_BLOCKS; b++) write_block(w, h_ct + (size_t)b * kBig, degrees ? degrees[b] : 3);
  *written = w.o;
  return (out && w.o <= cap) ? FB_OK : FB_ERR_ARG;
}

extern "C" int fb_string_ciphertext_from_bincode(const uint8_t* buf, size_t len, uint64_t* h_content, size_t cap_chars, size_t* n_chars) {
  if (!buf || !n_chars) return FB_ERR_ARG;
  Reader r{buf, len};
  uint64_t n = 0;
  if (!r.u64(n)) return FB_ERR_FORMAT;
  const size_t per = fb_radix_bincode_size();
  if (n > (len - 8) / per || 8 + n * per != len) return FB_ERR_FORMAT;
  *n_chars = (size_t)n;
  if (!h_content) return FB_OK;              // length query
  if (n > cap_chars) return FB_ERR_ARG;
  for (uint64_t i = 0; i < n; i++) {
    uint64_t deg[4];
    int rc = fb_radix_from_bincode(buf + 8 + i * per, per, h_content + i * 4 * kBig, deg);
    if (rc != FB_OK) return rc;
    // has_match assumes what encrypt_str produces (ciphertext.rs:32-40): fresh blocks, carries empty
    for (int b = 0; b < 4; b++)
      if (deg[b] > 3) return FB_ERR_FORMAT;
  }
  return FB_OK;
}

extern "C" int fb_string_ciphertext_to_bincode(const uint64_t* h_content, size_t n_chars, uint8_t* out, size_t cap, size_t* written) {
  if ((!h_content && n_chars) || !written) return FB_ERR_ARG;
  const size_t per = fb_radix_bincode_size();
  *written = 8 + n_chars * per;
  if (!out) return FB_OK;
  if (cap < *written) return FB_ERR_ARG;
  const uint64_t n = n_chars;
  std::memcpy(out, &n, 8);
  for (size_t i = 0; i < n_chars; i++) {
    size_t w = 0;
    int rc = fb_radix_to_bincode(h_content + i * 4 * kBig, nullptr, out + 8 + i * per, per, &w);
    if (rc != FB_OK) return rc;
  }
  return FB_OK;
}
