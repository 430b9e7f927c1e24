// CPU thread-by-thread emulation of the cluster (one sample per pair of CTAs) blind-rotation kernel, built from
// the SAME __host__ __device__ stage functions the CUDA kernel uses (fhe_regex_b200/csrc/br_duo.cuh).
// Test infrastructure: index / twiddle / swizzle logic checked in the build container, which has no GPU.
// Not part of the product.
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../fhe_regex_b200/csrc/br_duo.cuh"

using namespace fb;
using namespace fb::duo;

namespace {
struct Cta {                      // one CTA of the pair: polynomial `me`
  std::vector<c2> bufA, bufB, spec;
  std::vector<uint32_t> acc;      // shared copy [2048]
  uint32_t own[256][8];           // registers
  c2 X[256][4];                   // registers
  Cta() : bufA(kHalfN), bufB(kHalfN), spec(kHalfN), acc(kN) {}
};
struct Pair {
  std::vector<c2> tab;
  Tw tw[256];
  Cta cta[2];
  Pair() : tab(kTabC2) {
    make_duo_table(tab.data());
    for (int b = 0; b < 256; b++) load_tw(tw[b], tab.data(), b);
  }
  // forward stages 1..4 of CTA c: bufA -> bufB -> bufA -> bufB -> registers X (barriers between the loops)
  void forward_tail(Cta& c) {
    for (int b = 0; b < 256; b++) mid_stage<4, false>(c.bufA.data(), c.bufB.data(), b, tw[b].f1);
    for (int b = 0; b < 256; b++) mid_stage<16, false>(c.bufB.data(), c.bufA.data(), b, tw[b].f2);
    for (int b = 0; b < 256; b++) mid_stage<64, false>(c.bufA.data(), c.bufB.data(), b, tw[b].f3);
    for (int b = 0; b < 256; b++) {
      fwd_stage4(c.bufB.data(), b, c.X[b]);
      for (int k = 0; k < 4; k++) c.spec[b + 256 * k] = c.X[b][k];   // what the peer reads through DSMEM
    }
  }
  // inverse stages 1..3 of CTA c: bufA -> bufB -> bufA -> bufB
  void inverse_tail(Cta& c) {
    for (int b = 0; b < 256; b++) mid_stage<4, true>(c.bufA.data(), c.bufB.data(), b, tw[b].i1);
    for (int b = 0; b < 256; b++) mid_stage<16, true>(c.bufB.data(), c.bufA.data(), b, tw[b].i2);
    for (int b = 0; b < 256; b++) inv_stage3(c.bufA.data(), c.bufB.data(), b, tw[b]);
  }
};
}  // namespace

// spectrum (natural frequency order) of a standard-domain torus polynomial: must equal the product's key conversion
extern "C" void emu_duo_forward_torus(const uint64_t* poly /* [2048] */, c2* spec /* [1024] */) {
  Pair& P = *new Pair();
  Cta& c = P.cta[0];
  for (int b = 0; b < 256; b++) {
    c2 x[4];
    for (int r = 0; r < 4; r++) {
      const int j = b + 256 * r;
      x[r].x = (double)(int64_t)poly[j] * (1.0 / 18446744073709551616.0);
      x[r].y = (double)(int64_t)poly[j + 1024] * (1.0 / 18446744073709551616.0);
    }
    fwd_stage0_core(x, b, P.tw[b], c.bufA.data());
  }
  P.forward_tail(c);
  memcpy(spec, c.spec.data(), sizeof(c2) * kHalfN);
  delete &P;
}

// negacyclic product check: out = round(a_int (*) b_torus)
extern "C" void emu_duo_negacyclic_mul(const int64_t* a_int, const uint64_t* b_torus, uint64_t* out) {
  std::vector<c2> spec(kHalfN), zero(kHalfN);
  emu_duo_forward_torus(b_torus, spec.data());
  for (auto& z : zero) z = mk(0, 0);
  Pair& P = *new Pair();
  Cta& c = P.cta[0];
  for (int b = 0; b < 256; b++) {
    c2 x[4];
    for (int r = 0; r < 4; r++) x[r] = mk((double)a_int[b + 256 * r], (double)a_int[b + 256 * r + 1024]);
    fwd_stage0_core(x, b, P.tw[b], c.bufA.data());
  }
  P.forward_tail(c);
  c2 Xz[4] = {mk(0, 0), mk(0, 0), mk(0, 0), mk(0, 0)};
  for (int b = 0; b < 256; b++) mac_inv_stage0(c.X[b], Xz, spec.data(), zero.data(), b, P.tw[b], c.bufA.data());
  P.inverse_tail(c);
  for (int b = 0; b < 256; b++) {
    for (int k = 0; k < 8; k++) c.own[b][k] = 0;
    inv_stage4_accumulate(c.bufB.data(), b, c.own[b]);
    for (int k = 0; k < 4; k++) {
      out[b + 256 * k] = (uint64_t)c.own[b][2 * k] << 32;
      out[b + 256 * k + 1024] = (uint64_t)c.own[b][2 * k + 1] << 32;
    }
  }
  delete &P;
}

// small[743], lut[2048] -> acc[2][2048] (top 32 bits); max_steps < 0 means all 742
extern "C" void emu_duo_blind_rotate(const c2* fbsk, const uint64_t* small, const uint64_t* lut, uint64_t* acc_out, int max_steps) {
  Pair& P = *new Pair();
  const uint32_t bt = modswitch(small[kLweN]);
  const uint32_t rot = (4096u - bt) & 4095u;
  for (int me = 0; me < 2; me++)
    for (int b = 0; b < 256; b++)
      for (int r = 0; r < 4; r++)
        for (int h = 0; h < 2; h++) {
          const int j = b + 256 * r + 1024 * h;
          const uint32_t v = me == 0 ? 0u : (uint32_t)(rot_read(lut, (uint32_t)j, rot) >> 32);
          P.cta[me].own[b][2 * r + h] = v;
          P.cta[me].acc[j] = v;
        }
  const int steps = max_steps < 0 ? kLweN : max_steps;
  for (int i = 0; i < steps; i++) {
    const uint32_t a = modswitch(small[i]) & 4095u;
    if (small[i] == 0 || a == 0) continue;
    const c2* ggsw = fbsk + (size_t)i * 4 * kHalfN;
    for (int me = 0; me < 2; me++) {
      Cta& c = P.cta[me];
      for (int b = 0; b < 256; b++) fwd_stage0(c.acc.data(), c.own[b], a, b, P.tw[b], c.bufA.data());
      P.forward_tail(c);
    }
    // cluster barrier; every CTA reads the peer's spectrum
    for (int me = 0; me < 2; me++) {
      Cta& c = P.cta[me];
      const c2* g_own = ggsw + (size_t)(me * 2 + me) * kHalfN;          // GGSW[row me][column me]
      const c2* g_peer = ggsw + (size_t)((1 - me) * 2 + me) * kHalfN;   // GGSW[row 1-me][column me]
      for (int b = 0; b < 256; b++) {
        c2 Xp[4];
        for (int k = 0; k < 4; k++) Xp[k] = P.cta[1 - me].spec[b + 256 * k];
        mac_inv_stage0(c.X[b], Xp, g_own, g_peer, b, P.tw[b], c.bufA.data());
      }
      P.inverse_tail(c);
      for (int b = 0; b < 256; b++) {
        inv_stage4_accumulate(c.bufB.data(), b, c.own[b]);
        for (int k = 0; k < 4; k++) {
          c.acc[b + 256 * k] = c.own[b][2 * k];
          c.acc[b + 256 * k + 1024] = c.own[b][2 * k + 1];
        }
      }
    }
  }
  for (int me = 0; me < 2; me++)
    for (int j = 0; j < kN; j++) acc_out[me * kN + j] = (uint64_t)P.cta[me].acc[j] << 32;
  delete &P;
}
