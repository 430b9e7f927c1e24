// CPU thread-by-thread emulation of the radix-16 latency blind-rotation kernel, built from the SAME __host__ __device__
// stage functions the CUDA kernel uses (fhe_regex_b200/csrc/br_w16.cuh).  Test infrastructure: index / twiddle / layout
// logic checked in the build container, which has no GPU.  Not part of the product.
#include <cstdlib>
#include <cstring>
#include <vector>
#include "br_w16.cuh"

using namespace fb;
using namespace fb::w16;

namespace {
struct Cta {
  std::vector<c2> tab;          // [kTwC2][64]
  std::vector<c2> bufX;         // [2][kBufC2]: the transform buffer, every stage in place
  std::vector<uint32_t> acc;    // [2][2048] shared copy
  uint32_t own[128][32];        // registers: thread (P, u) owns coefficients u + 64m (+1024) of polynomial P
  c2 t1[64][16], f2[64][16], i1[64][16];
  Cta() : tab(kTwC2 * 64), bufX(2 * kBufC2), acc(2 * kN) {
    make_w16_table(tab.data());
    for (int v = 0; v < 64; v++) {
      load_t1(t1[v], tab.data(), v);
      load_f2(f2[v], tab.data(), v);
      load_i1(i1[v], tab.data(), v);
    }
  }
};
void forward_tail(Cta& c) {
  for (int tid = 0; tid < 128; tid++) fwd2(c.bufX.data() + (tid >> 6) * kBufC2, tid & 63, c.f2[tid & 63]);
}
void inverse_tail(Cta& c) {
  for (int tid = 0; tid < 128; tid++) inv2(c.bufX.data() + (tid >> 6) * kBufC2, tid & 63);
}
}  // namespace

// spectrum (natural frequency order) of two standard-domain torus polynomials: must equal the product's key conversion
extern "C" void emu_w16_forward_torus(const uint64_t* polys /* [2][2048] */, c2* spec /* [2][1024] */) {
  Cta& c = *new Cta();
  for (int tid = 0; tid < 128; tid++) {
    const int P = tid >> 6, u = tid & 63;
    c2 x[16];
    for (int m = 0; m < 16; m++) {
      const int j = u + 64 * m;
      x[m].x = (double)(int64_t)polys[P * kN + j] * (1.0 / 18446744073709551616.0);
      x[m].y = (double)(int64_t)polys[P * kN + j + 1024] * (1.0 / 18446744073709551616.0);
    }
    fwd1_core(x, u, c.t1[u], c.bufX.data() + P * kBufC2);
  }
  forward_tail(c);
  for (int tid = 0; tid < 128; tid++) spectrum_of(c.bufX.data() + (tid >> 6) * kBufC2, tid & 63, spec + (tid >> 6) * kHalfN);
  delete &c;
}

// negacyclic product check: out = round(a_int (*) b_torus), forward / pointwise / inverse through the emulated stages
extern "C" void emu_w16_negacyclic_mul(const int64_t* a_int, const uint64_t* b_torus, uint64_t* out) {
  std::vector<uint64_t> bp(2 * kN, 0);
  memcpy(bp.data(), b_torus, sizeof(uint64_t) * kN);
  std::vector<c2> spec(2 * kHalfN);
  emu_w16_forward_torus(bp.data(), spec.data());
  Cta& c = *new Cta();
  for (int tid = 0; tid < 128; tid++) {
    const int P = tid >> 6, u = tid & 63;
    c2 x[16];
    for (int m = 0; m < 16; m++) {
      const int j = u + 64 * m;
      x[m].x = P == 0 ? (double)a_int[j] : 0.0;
      x[m].y = P == 0 ? (double)a_int[j + 1024] : 0.0;
    }
    fwd1_core(x, u, c.t1[u], c.bufX.data() + P * kBufC2);
  }
  forward_tail(c);
  for (int tid = 0; tid < 128; tid++) {
    const int P = tid >> 6, y = tid & 63;
    mul_inv1(c.bufX.data() + P * kBufC2, spec.data(), y, c.i1[y], c.bufX.data() + P * kBufC2);
  }
  inverse_tail(c);
  std::fill(c.acc.begin(), c.acc.end(), 0u);
  memset(c.own, 0, sizeof c.own);
  for (int tid = 0; tid < 128; tid++) inv3_accumulate(c.bufX.data() + (tid >> 6) * kBufC2, tid & 63, c.t1[tid & 63], c.own[tid], c.acc.data() + (tid >> 6) * kN);
  for (int j = 0; j < kN; j++) out[j] = (uint64_t)c.acc[j] << 32;
  delete &c;
}

// small[743], lut[2048] -> acc[2][2048] (top 32 bits); max_steps < 0 means all 742
extern "C" void emu_w16_blind_rotate(const c2* fbsk, const uint64_t* small, const uint64_t* lut, uint64_t* acc_out, int max_steps) {
  Cta& c = *new Cta();
  const uint32_t bt = modswitch(small[kLweN]);
  const uint32_t rot = (4096u - bt) & 4095u;
  for (int j = 0; j < kN; j++) {
    c.acc[j] = 0;
    c.acc[kN + j] = (uint32_t)(rot_read(lut, j, rot) >> 32);
  }
  for (int tid = 0; tid < 128; tid++)
    for (int m = 0; m < 16; m++) {
      c.own[tid][2 * m] = c.acc[(tid >> 6) * kN + (tid & 63) + 64 * m];
      c.own[tid][2 * m + 1] = c.acc[(tid >> 6) * kN + (tid & 63) + 64 * m + 1024];
    }
  const int steps = max_steps < 0 ? kLweN : max_steps;
  for (int i = 0; i < steps; i++) {
    const uint32_t a = modswitch(small[i]) & 4095u;
    if (small[i] == 0 || a == 0) continue;
    const c2* ggsw = fbsk + (size_t)i * 4 * kHalfN;
    for (int tid = 0; tid < 128; tid++) fwd1(c.acc.data() + (tid >> 6) * kN, c.own[tid], a, tid & 63, c.t1[tid & 63], c.bufX.data() + (tid >> 6) * kBufC2);
    forward_tail(c);
    // the MAC stage in place: every thread reads (both polynomials), barrier, every thread writes its own slots
    static c2 r[128][16];
    for (int tid = 0; tid < 128; tid++) mac_fwd(c.bufX.data(), c.bufX.data() + kBufC2, ggsw, tid >> 6, tid & 63, r[tid]);
    for (int tid = 0; tid < 128; tid++) mac_store(r[tid], tid & 63, c.i1[tid & 63], c.bufX.data() + (tid >> 6) * kBufC2);
    inverse_tail(c);
    for (int tid = 0; tid < 128; tid++) inv3_accumulate(c.bufX.data() + (tid >> 6) * kBufC2, tid & 63, c.t1[tid & 63], c.own[tid], c.acc.data() + (tid >> 6) * kN);
  }
  for (int j = 0; j < 2 * kN; j++) acc_out[j] = (uint64_t)c.acc[j] << 32;
  delete &c;
}
