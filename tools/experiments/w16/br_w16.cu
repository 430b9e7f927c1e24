// br_w16.cu -- K2-K4, radix-16 latency variant: one PBS per CTA of 128 threads (64 per polynomial, 16 complex points per
// thread), for DAG levels narrower than two waves of SMs.  Same arithmetic as br_wide.cu (modulus switch, accumulator init,
// CMUX steps with the f64 negacyclic FFT external product on a 32-bit torus accumulator, sample extract) and the same Fourier
// bootstrapping key; the per-thread stages are in br_w16.cuh.
//
// Replaces, like br_wide.cu, the blind rotation under /root/reference/src/regex/execution.rs:76,93,110,143,173,190 for the
// narrow levels of has_match (engine.rs:22-35), whose cost is pure latency.
//
// Why a second latency kernel: br_wide.cu (1024 = 8 x 8 x 8 x 2) is bound by its shared-memory traffic -- three exchanges per
// transform, 4 170 wavefronts per CMUX step against a pipe that moves ~1 per cycle (profiles/r02_pair_kernel_probe.json: two
// samples on one SM take 1.9 x the time of one).  1024 = 16 x 16 x 4 needs two exchanges per transform (~3 100 wavefronts) and
// five barrier-separated stages instead of seven.  A thread then needs 48 twiddles; they live in TENSOR MEMORY (lane = thread
// index, 192 columns, fetched with tcgen05.ld), not in registers or shared memory.
//
// S samples per CTA (128 threads each).  A lone sample leaves one warp per scheduler, which cannot hide its own latencies
// (measured: 2.65 ms per wave, FP64 pipe 42 % busy, against 2.30 ms for br_wide.cu); with two or three samples per SM the
// schedulers have two or three independent warps each.  All stages work in place, so a sample needs 48.5 KiB of shared memory
// (one transform buffer of 2 x 1040 complex with rows padded to 65, the accumulator [2][2048] u32); the Fourier GGSW of the
// step is staged once per CTA (64 KiB, cp.async.bulk) and handed over like in the throughput kernel: the last warp past its
// MAC of step i issues the copy of the next needed step.
#include <cuda_runtime.h>
#include <stdint.h>
#include "br_tmem.cuh"
#include "br_w16.cuh"
#include "kernels.h"
#include "ptx_sync.cuh"

namespace fb {

namespace {
constexpr int kStageBytes16 = 4 * kHalfN * (int)sizeof(c2);            // 65536
constexpr int kBufBytes16 = 2 * w16::kBufC2 * (int)sizeof(c2);         // 33280: the transform buffer of a sample, both polynomials
constexpr int kSampleBytes16 = kBufBytes16 + 2 * kN * (int)sizeof(uint32_t);   // 49664
constexpr uint32_t kTw16Cols = 256;   // 192 used: T1 [0,64) | W_64^{bc} [64,128) | W_64^{-b'(cl + 4 e)} at 128 + 16 e + 4 (b' - 1)
constexpr size_t w16_smem_bytes(int S) { return (size_t)kStageBytes16 + (size_t)S * kSampleBytes16 + (size_t)S * kLweN * sizeof(uint16_t) + 32; }
static_assert(w16_smem_bytes(3) <= 232448, "exceeds the opt-in shared memory of an sm_100 CTA");
}  // namespace

// 16 complex twiddles out of 64 tensor-memory columns: two 32-column loads issued together, waited for at use
struct TwRegs {
  uint32_t lo[32], hi[32];
};
__device__ __forceinline__ void tw_issue(uint32_t taddr, TwRegs& r) {
  tmem_ld32_issue(taddr, r.lo);
  tmem_ld32_issue(taddr + 32, r.hi);
}
__device__ __forceinline__ void tw_wait(TwRegs& r, c2 (&w)[16]) {
  tmem_ld32_wait(r.lo);
  tmem_ld32_wait(r.hi);
#pragma unroll
  for (int e = 0; e < 16; e++) {
    const uint32_t* k = (e < 8) ? r.lo + 4 * e : r.hi + 4 * (e - 8);
    w[e].x = __hiloint2double((int)k[1], (int)k[0]);
    w[e].y = __hiloint2double((int)k[3], (int)k[2]);
  }
}

// Stages F2 and I2 share one copy of their code (w16::stage2)
__device__ __noinline__ void w16_stage2(c2* buf, int v, uint32_t ttw_f2, uint32_t sign) {
  TwRegs tr;
  c2 w[16];
  if (sign == 0u) {
    tw_issue(ttw_f2, tr);
    tw_wait(tr, w);
  } else {
#pragma unroll
    for (int c = 0; c < 16; c++) w[c] = w16::mk(1.0, 0.0);
  }
  w16::stage2(buf, v, w, sign);
}

template <int S>
__global__ void __launch_bounds__(w16::kThreads * S, 1)
blind_rotate_w16_kernel(const c2* __restrict__ fbsk, const uint64_t* __restrict__ small, const uint64_t* __restrict__ luts,
                        const uint32_t* __restrict__ lut_idx, uint64_t* __restrict__ out, const int32_t* __restrict__ out_rows,
                        const c2* __restrict__ tab, int count, int skew_cycles) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint16_t* at_all = reinterpret_cast<uint16_t*>(smem + kStageBytes16 + (size_t)S * kSampleBytes16);   // [S][742]; bit 15: this sample needs the step, bit 14: some sample does
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStageBytes16 + (size_t)S * kSampleBytes16 + (size_t)((S * kLweN * 2 + 7) / 8) * 8);
  uint32_t* done_cnt = reinterpret_cast<uint32_t*>(full_bar + 1);
  uint32_t* tmem_slot = done_cnt + 1;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s = tid >> 7, P = (tid >> 6) & 1, v = tid & 63;
  const int sample = blockIdx.x * S + s;
  const bool active = sample < count;
  const uint32_t n_warps_active = 4u * (uint32_t)min(S, count - (int)blockIdx.x * S);

  unsigned char* sbase = smem + kStageBytes16 + (size_t)s * kSampleBytes16;
  c2* buf = reinterpret_cast<c2*>(sbase);
  uint32_t* acc = reinterpret_cast<uint32_t*>(sbase + kBufBytes16);
  uint16_t* at = at_all + s * kLweN;

  if (warp == 0) tmem_alloc(tmem_slot, kTw16Cols);
  uint32_t b_tilde = 0;
  for (int i = tid & 127; i < kLweN; i += 128) {
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
  // warp w of a sample is (P, half) = (w >> 1 & 1, w & 1); its tensor-memory lane window is 32 (w % 4): the same for every
  // sample of the CTA, so the twiddles of thread index (P, v) are written once, by sample 0
  const uint32_t ttw = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  if (s == 0) {
#pragma unroll
    for (int g = 0; g < 12; g++) {
      uint32_t k[16];
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const c2 w = tab[(4 * g + e) * 64 + v];
        k[4 * e] = (uint32_t)__double2loint(w.x);
        k[4 * e + 1] = (uint32_t)__double2hiint(w.x);
        k[4 * e + 2] = (uint32_t)__double2loint(w.y);
        k[4 * e + 3] = (uint32_t)__double2hiint(w.y);
      }
      tmem_st16(ttw + 16 * g, k);
    }
    tmem_wait_st();
  }
  for (int i = tid; i < kLweN; i += w16::kThreads * S) {
    uint32_t f = 0;
#pragma unroll
    for (int ss = 0; ss < S; ss++) f |= at_all[ss * kLweN + i];
    if (f & 0x8000u) {
#pragma unroll
      for (int ss = 0; ss < S; ss++) at_all[ss * kLweN + i] |= 0x4000u;
    }
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
    mbar_arrive_expect_tx(full_bar, (uint32_t)kStageBytes16);
    const unsigned char* src = reinterpret_cast<const unsigned char*>(fbsk + (size_t)i * 4 * kHalfN);
#pragma unroll
    for (int c = 0; c < 4; c++) bulk_g2s(smem + c * (kStageBytes16 / 4), src + c * (kStageBytes16 / 4), kStageBytes16 / 4, full_bar);
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
    if (j < kLweN) issue_ggsw(j);
  }

  if (active) {
    // accumulator init: (0, lut * X^{-b}), top words; thread (P, v) owns coefficients v + 64 m and v + 64 m + 1024 of polynomial P
    uint32_t own[32];
    {
      const uint64_t* lut = luts + (size_t)lut_idx[sample] * kN;
      const uint32_t rot = (4096u - b_tilde) & 4095u;
#pragma unroll
      for (int m = 0; m < 16; m++) {
        const uint32_t j = (uint32_t)v + 64u * m;
        own[2 * m] = (P == 0) ? 0u : (uint32_t)(rot_read(lut, j, rot) >> 32);
        own[2 * m + 1] = (P == 0) ? 0u : (uint32_t)(rot_read(lut, j + 1024u, rot) >> 32);
        acc[P * kN + j] = own[2 * m];
        acc[P * kN + j + 1024u] = own[2 * m + 1];
      }
    }
    const int bar_half = 1 + 2 * s + P, bar_sample = 1 + 2 * S + s;
    auto half_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(bar_half) : "memory"); };
    auto sample_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(bar_sample) : "memory"); };
    sample_sync();

    uint32_t* accp = acc + P * kN;
    c2* buf_p = buf + P * w16::kBufC2;
    const c2* ggsw = reinterpret_cast<const c2*>(smem);
    uint32_t n_exec = 0;
    TwRegs tr;
    c2 t1[16];
    tw_issue(ttw, tr);
    tw_wait(tr, t1);
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
      w16::fwd1(accp, own, av & 4095u, v, t1, buf_p);
      half_sync();
      w16_stage2(buf_p, v, ttw + 64, 0u);
      tw_issue(ttw + 128, tr);
      mbar_wait(full_bar, par);
      sample_sync();                                // both spectra-to-be complete
      {
        c2 r[16];
        w16::mac_fwd(buf, buf + w16::kBufC2, ggsw, P, v, r);
        sample_sync();                              // nobody of this sample reads the buffers (or the GGSW stage) any more
        release_stage(i);
        c2 w[16];
        tw_wait(tr, w);
        w16::mac_store(r, v, w, buf_p);
      }
      if (P == 1 && skew_cycles > 0) {
        const long long t0 = clock64();
        while (clock64() - t0 < (long long)skew_cycles) {}
      }
      half_sync();
      w16_stage2(buf_p, v, ttw + 64, 0x80000000u);
      tw_issue(ttw, tr);
      half_sync();
      tw_wait(tr, t1);                              // T1 again: stage I3 now, stage F1 of the next step
      w16::inv3_accumulate(buf_p, v, t1, own, accp);
      half_sync();
    }
    sample_sync();

    // K4: sample extract of the constant coefficient: mask_0 = a_0, mask_j = -a_{N-j}; body = b_0
    {
      const size_t row = out_rows ? (size_t)out_rows[sample] : (size_t)sample;
      uint64_t* o = out + row * kBig;
      for (int j = tid & 127; j < kN; j += 128) {
        const uint32_t x = (j == 0) ? acc[0] : 0u - acc[kN - j];
        o[j] = (uint64_t)x << 32;
      }
      if ((tid & 127) == 0) o[kN] = (uint64_t)acc[kN] << 32;
    }
  }
  tmem_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTw16Cols);
}

size_t br_w16_table_bytes() { return (size_t)w16::kTwC2 * 64 * sizeof(c2); }
void br_w16_make_table(c2* host_tab) { w16::make_w16_table(host_tab); }

template <int S>
static cudaError_t launch_w16_s(const c2* fbsk, const uint64_t* small, const uint64_t* luts, const uint32_t* lut_idx, uint64_t* out,
                                const int32_t* out_rows, const c2* tab, int count, int skew, cudaStream_t st) {
  static PerDeviceOnce once;
  bool& configured = *once.slot();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(blind_rotate_w16_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w16_smem_bytes(S));
    if (e != cudaSuccess) return e;
    configured = true;
  }
  blind_rotate_w16_kernel<S><<<(count + S - 1) / S, w16::kThreads * S, w16_smem_bytes(S), st>>>(fbsk, small, luts, lut_idx, out, out_rows, tab, count, skew);
  return cudaGetLastError();
}

cudaError_t launch_blind_rotate_w16(const c2* fbsk, const uint64_t* small, const uint64_t* luts, const uint32_t* lut_idx,
                                    uint64_t* out, const int32_t* out_rows, const c2* tab, int count, int samples_per_cta, int skew, cudaStream_t st) {
  if (count <= 0) return cudaSuccess;
  switch (samples_per_cta) {
    case 1: return launch_w16_s<1>(fbsk, small, luts, lut_idx, out, out_rows, tab, count, skew, st);
    case 3: return launch_w16_s<3>(fbsk, small, luts, lut_idx, out, out_rows, tab, count, skew, st);
    default: return launch_w16_s<2>(fbsk, small, luts, lut_idx, out, out_rows, tab, count, skew, st);
  }
}

}  // namespace fb
