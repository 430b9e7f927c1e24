// br_wide.cu -- K2-K4, latency variant: one PBS per CTA (256 threads), for DAG levels narrower than two
// waves of SMs.  Same arithmetic as blind_rotate_kernel (modulus switch, accumulator init, CMUX steps with the
// f64 negacyclic FFT external product on a 32-bit torus accumulator, sample extract) and the same Fourier
// bootstrapping key; the per-thread stages are in br_wide.cuh.
//
// Replaces, like kernels.cu, the blind rotation under /root/reference/src/regex/execution.rs:76,93,110,143,173,190;
// this variant exists because has_match (engine.rs:22-35) ends in levels of a few PBS whose cost is pure latency.
//
// Shared memory (219 KiB): two GGSW stages (2 x 64 KiB, cp.async.bulk; the copy of the next needed step is
// issued at the top of the current one, a full step ahead) + two transform buffers [2][1152] complex (2 x 36 KiB,
// Stockham ping-pong, groups of 8 elements padded to 9: br_wide.cuh::LPad) + the accumulator [2][2048] u32 (16 KiB) + small tables.  Steps whose mask element
// switches to 0 (all of them for trivial inputs) are skipped through a compacted step list.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "br_wide.cuh"
#include "kernels.h"
#include "ptx_sync.cuh"

namespace fb {

namespace {
constexpr int kStageBytes = 4 * kHalfN * (int)sizeof(c2);  // 65536
// kMacPrefetch (template parameter of the kernel): groups of 4 GGSW values a thread fetches before the pre-MAC barrier
using WL = wide::LPad;                                     // padded transform buffers: every access [base + immediate]
constexpr int kBufBytes = 2 * WL::kBuf * (int)sizeof(c2);  // 36864
constexpr size_t kOffBufA = 2 * (size_t)kStageBytes;
constexpr size_t kOffBufB = kOffBufA + kBufBytes;
constexpr size_t kOffAcc = kOffBufB + kBufBytes;
constexpr size_t kOffAt = kOffAcc + 2 * kN * sizeof(uint32_t);
constexpr size_t kOffSteps = kOffAt + 768 * sizeof(uint16_t);
constexpr size_t kOffBars = kOffSteps + 768 * sizeof(uint16_t);
constexpr size_t kWideSmem = kOffBars + 2 * sizeof(uint64_t) + 16;
}  // namespace

__device__ __forceinline__ void half_sync(int P) { asm volatile("bar.sync %0, 128;" ::"r"(1 + P) : "memory"); }

template <int kMacPrefetch>
__global__ void __launch_bounds__(wide::kThreads, 1)
blind_rotate_wide_kernel(const c2* __restrict__ fbsk, const uint64_t* __restrict__ small, const uint64_t* __restrict__ luts,
                         const uint32_t* __restrict__ lut_idx, uint64_t* __restrict__ out, const int32_t* __restrict__ out_rows,
                         const c2* __restrict__ wtab, int count, int skew_cycles) {
  extern __shared__ __align__(128) unsigned char smem[];
  c2* bufA = reinterpret_cast<c2*>(smem + kOffBufA);
  c2* bufB = reinterpret_cast<c2*>(smem + kOffBufB);
  uint32_t* acc = reinterpret_cast<uint32_t*>(smem + kOffAcc);
  uint16_t* at = reinterpret_cast<uint16_t*>(smem + kOffAt);      // mod-switched ciphertext, bit 15: step needed
  uint16_t* steps = reinterpret_cast<uint16_t*>(smem + kOffSteps);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kOffBars);
  int* n_steps_p = reinterpret_cast<int*>(full_bar + 2);

  const int tid = threadIdx.x, P = tid >> 7, t = tid & 127;
  const int sample = blockIdx.x;
  if (sample >= count) return;

  wide::Tw tw;
  wide::load_tw(tw, wtab, t);
  for (int i = tid; i < 768; i += wide::kThreads) {
    uint32_t a = 0;
    if (i < kSmall) {
      const uint64_t x = small[(size_t)sample * kSmall + i];
      a = modswitch(x);
      if (i < kLweN) a = (a & 4095u) | ((x != 0 && (a & 4095u) != 0) ? 0x8000u : 0u);
    }
    at[i] = (uint16_t)a;
  }
  __syncthreads();

  auto issue_ggsw = [&](int i, int b) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive_expect_tx(full_bar + b, (uint32_t)kStageBytes);
    const unsigned char* src = reinterpret_cast<const unsigned char*>(fbsk + (size_t)i * 4 * kHalfN);
    unsigned char* dst = smem + (size_t)b * kStageBytes;
#pragma unroll
    for (int c = 0; c < 4; c++) bulk_g2s(dst + c * (kStageBytes / 4), src + c * (kStageBytes / 4), kStageBytes / 4, full_bar + b);
  };

  if (tid == 0) {
    int n = 0;
    for (int i = 0; i < kLweN; i++)
      if (at[i] & 0x8000u) steps[n++] = (uint16_t)i;
    *n_steps_p = n;
    mbar_init(full_bar, 1);
    mbar_init(full_bar + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (n > 0) issue_ggsw(steps[0], 0);
  }
  // accumulator init: (0, lut * X^{-b}), top words; thread (P, t) owns coefficients t + 128m and t + 128m + 1024 of polynomial P
  uint32_t own[16];
  {
    const uint64_t* lut = luts + (size_t)lut_idx[sample] * kN;
    const uint32_t rot = (4096u - (uint32_t)at[kLweN]) & 4095u;
#pragma unroll
    for (int m = 0; m < 8; m++) {
      const uint32_t j = (uint32_t)t + 128u * m;
      own[2 * m] = (P == 0) ? 0u : (uint32_t)(rot_read(lut, j, rot) >> 32);
      own[2 * m + 1] = (P == 0) ? 0u : (uint32_t)(rot_read(lut, j + 1024u, rot) >> 32);
      acc[P * kN + j] = own[2 * m];
      acc[P * kN + j + 1024u] = own[2 * m + 1];
    }
  }
  __syncthreads();
  const int n_steps = *n_steps_p;

  uint32_t* accp = acc + P * kN;
  c2* bufA_p = bufA + P * WL::kBuf;
  c2* bufB_p = bufB + P * WL::kBuf;

#pragma unroll 1
  for (int n = 0; n < n_steps; n++) {
    const int i = steps[n];
    const uint32_t a = (uint32_t)at[i] & 4095u;
    // the other stage was last read by the MAC of step n-1, four barriers ago
    if (tid == 0 && n + 1 < n_steps) issue_ggsw(steps[n + 1], (n + 1) & 1);
    // The two polynomials only meet in the Fourier MAC: everywhere else a half (128 threads) synchronises on its
    // own named barrier, so one half's transform arithmetic overlaps the other half's shared-memory traffic.
    wide::fwd_stage1<WL>(accp, own, a, t, tw, bufA_p);
    half_sync(P);
    wide::fwd_stage2<WL>(bufA_p, bufB_p, t, tw);
    half_sync(P);
    wide::fwd_stage3<WL>(bufB_p, bufA_p, t);
    // the staged GGSW does not depend on the other warps: wait for it and fetch half of this thread's values before the
    // barrier, so that their shared-memory latency falls into the barrier wait
    mbar_wait(full_bar + (n & 1), (uint32_t)(n >> 1) & 1u);
    const c2* ggsw = reinterpret_cast<const c2*>(smem + (size_t)(n & 1) * kStageBytes);
    c2 gpre[4 * (kMacPrefetch > 0 ? kMacPrefetch : 1)];
    wide::mac_prefetch<kMacPrefetch>(ggsw, P, t, gpre);
    __syncthreads();                              // both spectra complete
    wide::mac_inv_stage1<kMacPrefetch, WL>(bufA, bufA + WL::kBuf, ggsw, gpre, P, t, tw, bufB_p);
    __syncthreads();                              // nobody reads bufA (or this GGSW stage) any more
    // the halves leave this barrier in phase; holding one back by a fraction of a stage makes its shared-memory
    // bursts fall into the other half's arithmetic for the six per-half stages until they meet again
    if (P == 1 && skew_cycles > 0) {
      const long long t0 = clock64();
      while (clock64() - t0 < (long long)skew_cycles) {}
    }
    wide::inv_stage2<WL>(bufB_p, bufA_p, t, tw);
    half_sync(P);
    wide::inv_stage3<WL>(bufA_p, bufB_p, t);
    half_sync(P);
    wide::phaseC_accumulate<WL>(bufB_p, t, own, accp);
    half_sync(P);
  }
  __syncthreads();

  // K4: sample extract of the constant coefficient: mask_0 = a_0, mask_j = -a_{N-j}; body = b_0
  {
    const size_t row = out_rows ? (size_t)out_rows[sample] : (size_t)sample;
    uint64_t* o = out + row * kBig;
    for (int j = tid; j < kN; j += wide::kThreads) {
      const uint32_t v = (j == 0) ? acc[0] : 0u - acc[kN - j];
      o[j] = (uint64_t)v << 32;
    }
    if (tid == 0) o[kN] = (uint64_t)acc[kN] << 32;
  }
}

size_t br_wide_table_bytes() { return (size_t)wide::kTabC2 * sizeof(c2); }
void br_wide_make_table(c2* host_tab) { wide::make_wide_table(host_tab); }

template <int NPRE>
static cudaError_t launch_wide_n(const c2* fbsk, const uint64_t* small, const uint64_t* luts, const uint32_t* lut_idx, uint64_t* out,
                                 const int32_t* out_rows, const c2* wtab, int count, int skew, cudaStream_t st) {
  static PerDeviceOnce once;
  bool& configured = *once.slot();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(blind_rotate_wide_kernel<NPRE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWideSmem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  blind_rotate_wide_kernel<NPRE><<<count, wide::kThreads, kWideSmem, st>>>(fbsk, small, luts, lut_idx, out, out_rows, wtab, count, skew);
  return cudaGetLastError();
}

cudaError_t launch_blind_rotate_wide(const c2* fbsk, const uint64_t* small, const uint64_t* luts, const uint32_t* lut_idx,
                                     uint64_t* out, const int32_t* out_rows, const c2* wtab, int count, int skew, int npre, cudaStream_t st) {
  if (count <= 0) return cudaSuccess;
  // npre = GGSW groups fetched before the pre-MAC barrier: 0 -> 2.44 ms, 1/2 -> 2.40, 3 -> 2.30, 4 -> 2.39 (254 registers);
  // skew (cycles): measured on B200: 0 -> 2.46 ms, 100..300 -> 2.38-2.40 ms per 148-PBS wave
  switch (npre) {
    case 0: return launch_wide_n<0>(fbsk, small, luts, lut_idx, out, out_rows, wtab, count, skew, st);
    case 1: return launch_wide_n<1>(fbsk, small, luts, lut_idx, out, out_rows, wtab, count, skew, st);
    case 4: return launch_wide_n<4>(fbsk, small, luts, lut_idx, out, out_rows, wtab, count, skew, st);
    case 2: return launch_wide_n<2>(fbsk, small, luts, lut_idx, out, out_rows, wtab, count, skew, st);
    default: return launch_wide_n<3>(fbsk, small, luts, lut_idx, out, out_rows, wtab, count, skew, st);
  }
}

}  // namespace fb
