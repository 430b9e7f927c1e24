// br_wide2.cu -- K2-K4, latency variant for levels between one and two waves of SMs: TWO PBS per CTA (512 threads), each spread
// over 256 threads with the stages of br_wide.cuh.  A level of 149 .. 296 PBS costs two launches' worth of one-PBS-per-SM
// waves with br_wide.cu (2 x 2.3 ms); here both samples of an SM run at the same time and fill each other's barrier and
// shared-memory waits.
//
// Replaces, like br_wide.cu, the blind rotation under /root/reference/src/regex/execution.rs:76,93,110,143,173,190 for the
// levels of has_match (engine.rs:22-35) that are a little wider than the GPU.
//
// What changes against br_wide.cu so that two samples fit one SM:
//   * registers: 512 threads leave 128 registers per thread, so the 30 per-thread twiddles (120 registers in br_wide.cu) live in
//     TENSOR MEMORY: lane t of the TMEM holds the twiddles of thread index t = tid & 127 (the four warps with the same
//     warp % 4 -- the two polynomials of both samples -- share a lane window, and they need the same values), 32 columns per
//     stage, fetched with one tcgen05.ld ahead of the barrier that precedes the stage;
//   * shared memory: ONE 64 KiB stage for the Fourier GGSW of the step, shared by both samples and handed over like in the
//     throughput kernel (the last warp past its MAC of step i issues the bulk copy of the next needed step), 2 x (two transform
//     buffers 64 KiB + accumulator 16 KiB), the mod-switched masks of both samples: 232 368 of 232 448 bytes.
#include <cuda_runtime.h>
#include <stdint.h>
#include "br_tmem.cuh"
#include "br_wide.cuh"
#include "kernels.h"
#include "ptx_sync.cuh"

namespace fb {

namespace {
constexpr int kStageBytes2 = 4 * kHalfN * (int)sizeof(c2);      // 65536
constexpr int kBufBytes2 = 2 * kHalfN * (int)sizeof(c2);        // 32768: one transform buffer, both polynomials
constexpr int kSampleBytes = 2 * kBufBytes2 + 2 * kN * (int)sizeof(uint32_t);   // 81920
constexpr size_t kOffSample = (size_t)kStageBytes2;
constexpr size_t kOffAt2 = kOffSample + 2 * (size_t)kSampleBytes;               // [2][742] u16
constexpr size_t kOffBars2 = kOffAt2 + 2 * kLweN * sizeof(uint16_t);
constexpr size_t kWide2Smem = kOffBars2 + 8 + 4 + 4 + 4;                         // mbarrier, arrival counter, TMEM slot, first step
constexpr uint32_t kTwCols = 128;                                              // 4 groups of 32 columns
static_assert(kWide2Smem <= 232448, "exceeds the opt-in shared memory of an sm_100 CTA");
}  // namespace

__device__ __forceinline__ void unpack4(const uint32_t* k, c2* w) {
  w->x = __hiloint2double((int)k[1], (int)k[0]);
  w->y = __hiloint2double((int)k[3], (int)k[2]);
}

template <int kMacPrefetch>
__global__ void __launch_bounds__(512, 1)
blind_rotate_wide2_kernel(const c2* __restrict__ fbsk, const uint64_t* __restrict__ small, const uint64_t* __restrict__ luts,
                          const uint32_t* __restrict__ lut_idx, uint64_t* __restrict__ out, const int32_t* __restrict__ out_rows,
                          const c2* __restrict__ wtab, int count, int skew_cycles, int sample_offset) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint16_t* at_all = reinterpret_cast<uint16_t*>(smem + kOffAt2);   // bit 15: this sample needs the step, bit 14: some sample does
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kOffBars2);
  uint32_t* done_cnt = reinterpret_cast<uint32_t*>(full_bar + 1);
  uint32_t* tmem_slot = done_cnt + 1;
  int* first_step_p = reinterpret_cast<int*>(tmem_slot + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s = tid >> 8, P = (tid >> 7) & 1, t = tid & 127;
  const int sample = blockIdx.x * 2 + s;
  const bool active = sample < count;
  const uint32_t n_warps_active = (blockIdx.x * 2 + 1 < count) ? 16u : 8u;

  unsigned char* sbase = smem + kOffSample + (size_t)s * kSampleBytes;
  c2* bufA = reinterpret_cast<c2*>(sbase);
  c2* bufB = reinterpret_cast<c2*>(sbase + kBufBytes2);
  uint32_t* acc = reinterpret_cast<uint32_t*>(sbase + 2 * kBufBytes2);
  uint16_t* at = at_all + s * kLweN;

  if (warp == 0) tmem_alloc(tmem_slot, kTwCols);
  uint32_t b_tilde = 0;
  for (int i = tid & 255; i < kLweN; i += 256) {
    uint32_t a = 0;
    if (active) {
      const uint64_t x = small[(size_t)sample * kSmall + i];
      a = modswitch(x);
      a = (a & 4095u) | ((x != 0 && (a & 4095u) != 0) ? 0x8000u : 0u);
    }
    at[i] = (uint16_t)a;
  }
  if (active) b_tilde = modswitch(small[(size_t)sample * kSmall + kLweN]);
  if (tid == 0) {
    mbar_init(full_bar, 1);
    *done_cnt = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t ttw = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  // twiddles of thread index t into lane t of the tensor memory: groups of 32 columns w1f[8] | w1i[7] . | w2f[7] . | w2i[8]
  if (warp < 4) {
    constexpr int first[4] = {0, 8, 15, 22};
    constexpr int cnt[4] = {8, 7, 7, 8};
#pragma unroll
    for (int g = 0; g < 4; g++) {
#pragma unroll
      for (int h = 0; h < 2; h++) {
        uint32_t v[16];
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const int k = 4 * h + e;
          c2 w;
          w.x = 0.0;
          w.y = 0.0;
          if (k < cnt[g]) w = wtab[(first[g] + k) * 128 + t];
          v[4 * e] = (uint32_t)__double2loint(w.x);
          v[4 * e + 1] = (uint32_t)__double2hiint(w.x);
          v[4 * e + 2] = (uint32_t)__double2loint(w.y);
          v[4 * e + 3] = (uint32_t)__double2hiint(w.y);
        }
        tmem_st16(ttw + 32 * g + 16 * h, v);
      }
    }
    tmem_wait_st();
  }
  for (int i = tid; i < kLweN; i += 512) {
    const uint16_t f = (uint16_t)((at_all[i] | at_all[kLweN + i]) & 0x8000u);
    at_all[i] |= f >> 1;
    at_all[kLweN + i] |= f >> 1;
  }
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();

  auto next_needed = [&](int i) {
    int j = i + 1;
    while (j < kLweN && !(at_all[j] & 0x4000u)) j++;
    return j;
  };
  auto issue_ggsw = [&](int i) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive_expect_tx(full_bar, (uint32_t)kStageBytes2);
    const unsigned char* src = reinterpret_cast<const unsigned char*>(fbsk + (size_t)i * 4 * kHalfN);
#pragma unroll
    for (int c = 0; c < 4; c++) bulk_g2s(smem + c * (kStageBytes2 / 4), src + c * (kStageBytes2 / 4), kStageBytes2 / 4, full_bar);
  };
  // the warp is done with the staged GGSW of step i: the last one to say so issues the copy of the next needed step
  auto release_stage = [&](int i) {
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      if (atomicAdd(done_cnt, 1u) == n_warps_active - 1u) {
        *reinterpret_cast<volatile uint32_t*>(done_cnt) = 0u;
        const int j = next_needed(i);
        if (j < kLweN) issue_ggsw(j);
      }
    }
  };
  if (tid == 0) {
    const int j = next_needed(-1);
    *first_step_p = j;
    if (j < kLweN) issue_ggsw(j);
  }

  if (active) {
    // accumulator init: (0, lut * X^{-b}), top words; thread (P, t) owns coefficients t + 128m and t + 128m + 1024 of polynomial P
    uint32_t own[16];
    {
      const uint64_t* lut = luts + (size_t)lut_idx[sample] * kN;
      const uint32_t rot = (4096u - b_tilde) & 4095u;
#pragma unroll
      for (int m = 0; m < 8; m++) {
        const uint32_t j = (uint32_t)t + 128u * m;
        own[2 * m] = (P == 0) ? 0u : (uint32_t)(rot_read(lut, j, rot) >> 32);
        own[2 * m + 1] = (P == 0) ? 0u : (uint32_t)(rot_read(lut, j + 1024u, rot) >> 32);
        acc[P * kN + j] = own[2 * m];
        acc[P * kN + j + 1024u] = own[2 * m + 1];
      }
    }
    const int bar_half = 1 + 2 * s + P, bar_sample = 5 + s;
    auto half_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(bar_half) : "memory"); };
    auto sample_sync = [&]() { asm volatile("bar.sync %0, 256;" ::"r"(bar_sample) : "memory"); };
    sample_sync();

    uint32_t* accp = acc + P * kN;
    c2* bufA_p = bufA + P * kHalfN;
    c2* bufB_p = bufB + P * kHalfN;
    const c2* ggsw = reinterpret_cast<const c2*>(smem);
    uint32_t n_exec = 0;
    uint32_t k[32];
    // the two samples only meet at the hand-over of the GGSW stage (once per step): starting the second one a fraction of a
    // step late keeps its shared-memory-heavy stages away from the first one's
    if (s == 1 && sample_offset > 0) {
      const long long t0 = clock64();
      while (clock64() - t0 < (long long)sample_offset) {}
    }
    tmem_ld32_issue(ttw, k);   // w1f of the first executed step
#pragma unroll 1
    for (int i = next_needed(-1); i < kLweN; i = next_needed(i)) {
      const uint32_t av = at[i];
      const uint32_t par = n_exec & 1u;
      n_exec++;
      if (!(av & 0x8000u)) {   // this sample skips the step (mask element switched to 0) but takes part in the hand-over
        mbar_wait(full_bar, par);
        release_stage(i);
        continue;
      }
      const uint32_t a = av & 4095u;
      {
        c2 w[8];
        tmem_ld32_wait(k);
#pragma unroll
        for (int e = 0; e < 8; e++) unpack4(k + 4 * e, w + e);
        tmem_ld32_issue(ttw + 64, k);   // w2f
        wide::fwd_stage1(accp, own, a, t, w, bufA_p);
      }
      half_sync();
      {
        c2 w[7];
        tmem_ld32_wait(k);
#pragma unroll
        for (int e = 0; e < 7; e++) unpack4(k + 4 * e, w + e);
        tmem_ld32_issue(ttw + 32, k);   // w1i
        wide::fwd_stage2(bufA_p, bufB_p, t, w);
      }
      half_sync();
      wide::fwd_stage3(bufB_p, bufA_p, t);
      mbar_wait(full_bar, par);
      c2 gpre[4 * (kMacPrefetch > 0 ? kMacPrefetch : 1)];
      wide::mac_prefetch<kMacPrefetch>(ggsw, P, t, gpre);
      sample_sync();                                // both spectra complete
      {
        c2 w[7];
        tmem_ld32_wait(k);
#pragma unroll
        for (int e = 0; e < 7; e++) unpack4(k + 4 * e, w + e);
        tmem_ld32_issue(ttw + 96, k);   // w2i
        wide::mac_inv_stage1<kMacPrefetch>(bufA, bufA + kHalfN, ggsw, gpre, P, t, w, bufB_p);
      }
      sample_sync();                                // nobody of this sample reads bufA or the GGSW stage any more
      release_stage(i);
      if (P == 1 && skew_cycles > 0) {
        const long long t0 = clock64();
        while (clock64() - t0 < (long long)skew_cycles) {}
      }
      {
        c2 w[8];
        tmem_ld32_wait(k);
#pragma unroll
        for (int e = 0; e < 8; e++) unpack4(k + 4 * e, w + e);
        tmem_ld32_issue(ttw, k);        // w1f of the next step
        wide::inv_stage2(bufB_p, bufA_p, t, w);
      }
      half_sync();
      wide::inv_stage3(bufA_p, bufB_p, t);
      half_sync();
      wide::phaseC_accumulate(bufB_p, t, own, accp);
      half_sync();
    }
    tmem_ld32_wait(k);
    sample_sync();

    // K4: sample extract of the constant coefficient: mask_0 = a_0, mask_j = -a_{N-j}; body = b_0
    {
      const size_t row = out_rows ? (size_t)out_rows[sample] : (size_t)sample;
      uint64_t* o = out + row * kBig;
      for (int j = tid & 255; j < kN; j += 256) {
        const uint32_t v = (j == 0) ? acc[0] : 0u - acc[kN - j];
        o[j] = (uint64_t)v << 32;
      }
      if ((tid & 255) == 0) o[kN] = (uint64_t)acc[kN] << 32;
    }
  }
  tmem_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTwCols);
}

template <int NPRE>
static cudaError_t launch_wide2_n(const c2* fbsk, const uint64_t* small, const uint64_t* luts, const uint32_t* lut_idx, uint64_t* out,
                                  const int32_t* out_rows, const c2* wtab, int count, int skew, int offset, cudaStream_t st) {
  static PerDeviceOnce once;
  bool& configured = *once.slot();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(blind_rotate_wide2_kernel<NPRE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWide2Smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  blind_rotate_wide2_kernel<NPRE><<<(count + 1) / 2, 512, kWide2Smem, st>>>(fbsk, small, luts, lut_idx, out, out_rows, wtab, count, skew, offset);
  return cudaGetLastError();
}

cudaError_t launch_blind_rotate_wide2(const c2* fbsk, const uint64_t* small, const uint64_t* luts, const uint32_t* lut_idx,
                                      uint64_t* out, const int32_t* out_rows, const c2* wtab, int count, int skew, int npre, int offset, cudaStream_t st) {
  if (count <= 0) return cudaSuccess;
  switch (npre) {
    case 0: return launch_wide2_n<0>(fbsk, small, luts, lut_idx, out, out_rows, wtab, count, skew, offset, st);
    case 2: return launch_wide2_n<2>(fbsk, small, luts, lut_idx, out, out_rows, wtab, count, skew, offset, st);
    default: return launch_wide2_n<1>(fbsk, small, luts, lut_idx, out, out_rows, wtab, count, skew, offset, st);
  }
}

}  // namespace fb
