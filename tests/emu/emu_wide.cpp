// CPU thread-by-thread emulation of the latency (one sample per CTA) blind-rotation kernel, built from the
// SAME __host__ __device__ stage functions the CUDA kernel uses (fhe_regex_b200/csrc/br_wide.cuh).
// Test infrastructure: index / twiddle / swizzle logic checked in the build container, which has no GPU.
// Not part of the product.
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../fhe_regex_b200/csrc/br_wide.cuh"

using namespace fb;
using namespace fb::wide;
using WL = fb::wide::LPad;   // the layout br_wide.cu uses (br_wide2.cu keeps LSwz: the GPU test compares the two bit for bit)

namespace {
struct Cta {
  std::vector<c2> tab;          // [kTabC2]
  std::vector<c2> bufA, bufB;   // [2][WL::kBuf]
  std::vector<uint32_t> acc;    // [2][2048] shared copy
  uint32_t own[256][16];        // registers: thread (P, t) owns coefficients t + 128m (+1024) of polynomial P
  Tw tw[256];
  Cta() : tab(kTabC2), bufA(2 * WL::kBuf), bufB(2 * WL::kBuf), acc(2 * kN) {
    make_wide_table(tab.data());
    for (int tid = 0; tid < 256; tid++) load_tw(tw[tid], tab.data(), tid & 127);
  }
};

// stages 2 and 3 of the forward transform of both polynomials: bufA -> bufB -> bufA (barriers between the loops)
void forward_tail(Cta& c) {
  for (int tid = 0; tid < 256; tid++) fwd_stage2<WL>(c.bufA.data() + (tid >> 7) * WL::kBuf, c.bufB.data() + (tid >> 7) * WL::kBuf, tid & 127, c.tw[tid]);
  for (int tid = 0; tid < 256; tid++) fwd_stage3<WL>(c.bufB.data() + (tid >> 7) * WL::kBuf, c.bufA.data() + (tid >> 7) * WL::kBuf, tid & 127);
}
void inverse_tail(Cta& c) {
  for (int tid = 0; tid < 256; tid++) inv_stage2<WL>(c.bufB.data() + (tid >> 7) * WL::kBuf, c.bufA.data() + (tid >> 7) * WL::kBuf, tid & 127, c.tw[tid]);
  for (int tid = 0; tid < 256; tid++) inv_stage3<WL>(c.bufA.data() + (tid >> 7) * WL::kBuf, c.bufB.data() + (tid >> 7) * WL::kBuf, tid & 127);
}
}  // namespace

// spectrum (natural frequency order) of two standard-domain torus polynomials: must equal the product's key conversion
extern "C" void emu_wide_forward_torus(const uint64_t* polys /* [2][2048] */, c2* spec /* [2][1024] */) {
  Cta& c = *new Cta();
  for (int tid = 0; tid < 256; tid++) {
    const int P = tid >> 7, t = tid & 127;
    c2 x[8];
    for (int m = 0; m < 8; m++) {
      const int j = t + 128 * m;
      x[m].x = (double)(int64_t)polys[P * kN + j] * (1.0 / 18446744073709551616.0);
      x[m].y = (double)(int64_t)polys[P * kN + j + 1024] * (1.0 / 18446744073709551616.0);
    }
    fwd_stage1_core<WL>(x, t, c.tw[tid], c.bufA.data() + P * WL::kBuf);
  }
  forward_tail(c);
  for (int P = 0; P < 2; P++)
    for (int k = 0; k < 512; k++) {
      const c2 a = c.bufA[P * WL::kBuf + WL::at(k)], b = c.bufA[P * WL::kBuf + WL::at(k + 512)];
      spec[P * kHalfN + k] = cadd(a, b);
      spec[P * kHalfN + k + 512] = csub(a, b);
    }
  delete &c;
}

// negacyclic product check: out = round(a_int (*) b_torus), forward / pointwise / inverse through the emulated stages
extern "C" void emu_wide_negacyclic_mul(const int64_t* a_int, const uint64_t* b_torus, uint64_t* out) {
  std::vector<uint64_t> bp(2 * kN, 0);
  memcpy(bp.data(), b_torus, sizeof(uint64_t) * kN);
  std::vector<c2> spec(2 * kHalfN);
  emu_wide_forward_torus(bp.data(), spec.data());
  Cta& c = *new Cta();
  for (int tid = 0; tid < 256; tid++) {
    const int P = tid >> 7, t = tid & 127;
    c2 x[8];
    for (int m = 0; m < 8; m++) {
      const int j = t + 128 * m;
      x[m].x = P == 0 ? (double)a_int[j] : 0.0;
      x[m].y = P == 0 ? (double)a_int[j + 1024] : 0.0;
    }
    fwd_stage1_core<WL>(x, t, c.tw[tid], c.bufA.data() + P * WL::kBuf);
  }
  forward_tail(c);
  for (int tid = 0; tid < 256; tid++) {
    const int P = tid >> 7, t = tid & 127;
    mul_inv_stage1<WL>(c.bufA.data() + P * WL::kBuf, spec.data(), t, c.tw[tid], c.bufB.data() + P * WL::kBuf);
  }
  inverse_tail(c);
  std::fill(c.acc.begin(), c.acc.end(), 0u);
  memset(c.own, 0, sizeof c.own);
  for (int tid = 0; tid < 256; tid++) phaseC_accumulate<WL>(c.bufB.data() + (tid >> 7) * WL::kBuf, tid & 127, c.own[tid], c.acc.data() + (tid >> 7) * kN);
  for (int j = 0; j < kN; j++) out[j] = (uint64_t)c.acc[j] << 32;
  delete &c;
}

// small[743], lut[2048] -> acc[2][2048] (top 32 bits); max_steps < 0 means all 742
extern "C" void emu_wide_blind_rotate(const c2* fbsk, const uint64_t* small, const uint64_t* lut, uint64_t* acc_out, int max_steps) {
  Cta& c = *new Cta();
  const uint32_t bt = modswitch(small[kLweN]);
  const uint32_t rot = (4096u - bt) & 4095u;
  for (int j = 0; j < kN; j++) {
    c.acc[j] = 0;
    c.acc[kN + j] = (uint32_t)(rot_read(lut, j, rot) >> 32);
  }
  for (int tid = 0; tid < 256; tid++)
    for (int m = 0; m < 8; m++) {
      c.own[tid][2 * m] = c.acc[(tid >> 7) * kN + (tid & 127) + 128 * m];
      c.own[tid][2 * m + 1] = c.acc[(tid >> 7) * kN + (tid & 127) + 128 * m + 1024];
    }
  const int steps = max_steps < 0 ? kLweN : max_steps;
  for (int i = 0; i < steps; i++) {
    const uint32_t a = modswitch(small[i]) & 4095u;
    if (small[i] == 0 || a == 0) continue;
    const c2* ggsw = fbsk + (size_t)i * 4 * kHalfN;
    for (int tid = 0; tid < 256; tid++) fwd_stage1<WL>(c.acc.data() + (tid >> 7) * kN, c.own[tid], a, tid & 127, c.tw[tid], c.bufA.data() + (tid >> 7) * WL::kBuf);
    forward_tail(c);
    for (int tid = 0; tid < 256; tid++)
      mac_inv_stage1<0, WL>(c.bufA.data(), c.bufA.data() + WL::kBuf, ggsw, nullptr, tid >> 7, tid & 127, c.tw[tid], c.bufB.data() + (tid >> 7) * WL::kBuf);
    inverse_tail(c);
    for (int tid = 0; tid < 256; tid++) phaseC_accumulate<WL>(c.bufB.data() + (tid >> 7) * WL::kBuf, tid & 127, c.own[tid], c.acc.data() + (tid >> 7) * kN);
  }
  for (int j = 0; j < 2 * kN; j++) acc_out[j] = (uint64_t)c.acc[j] << 32;
  delete &c;
}
