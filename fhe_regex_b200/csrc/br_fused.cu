// br_fused.cu -- throughput blind rotation, "fused" CMUX body (K2-K4 of SURVEY.md; same function as
// kernels.cu::blind_rotate_kernel<4>, same data layout, same Fourier key).
//
// Replaces the arithmetic under /root/reference/src/regex/execution.rs:76-190 (tfhe-rs blind_rotate_assign +
// sample extract) for batches that fill the GPU at 4 PBS per SM.
//
// Why a second body: blind_rotate_kernel<4> runs its 8 warps in lock-step through phases that are either FP64-bound
// (the 32-point transforms) or shared-memory / integer bound (decomposition, Fourier MAC, transposes), and the phases
// add up (profiles/r02_blind_rotate_base_phases.txt: FP64 pipe 64 % busy).  Here the per-thread program is re-ordered so
// that both kinds of work sit next to each other in program order:
//   * phase A + forward pass 1: the digits of slot r are produced in bit-reversed slot order and every butterfly is
//     issued as soon as its inputs exist (br_core.cuh::phaseA_f1, fft32_gen.h::fft32_f1_step<n>);
//   * forward pass 2 + Fourier MAC + inverse pass 1: the last three forward stages and the first three inverse stages
//     stay inside aligned blocks of 8 registers, so the MAC (shuffles + key reads from the staged GGSW) runs block by
//     block between them;
//   * inverse pass 2 + phase C: the untwist is folded into the last-stage butterflies as a pending rotation, its real
//     factor into the FMAs of a three-instruction torus rounding (br_core.cuh::torus32_round_scaled).
#include <cuda_runtime.h>
#include <stdint.h>
#include "br_core.cuh"
#include "br_tmem.cuh"
#include "kernels.h"
#include "ptx_sync.cuh"

namespace fb {

template <int B>
__device__ __forceinline__ void mid_block(double (&xr)[32], double (&xi)[32], const c2* __restrict__ b_own, const c2* __restrict__ b_in) {
  fft32_fwd_s3<B>(xr, xi);
  fft32_fwd_s45<2 * B>(xr, xi);
  fft32_fwd_s45<2 * B + 1>(xr, xi);
#pragma unroll
  for (int t = 0; t < 8; t++) {
    const int q = 8 * B + t;
    const int k2 = brev5(q);
    const double pr = __shfl_xor_sync(0xffffffffu, xr[q], 16);
    const double pi = __shfl_xor_sync(0xffffffffu, xi[q], 16);
    mac_point2(xr[q], xi[q], pr, pi, b_own[32 * k2], b_in[32 * k2]);
  }
  fft32_inv_s12<2 * B>(xr, xi);
  fft32_inv_s12<2 * B + 1>(xr, xi);
  fft32_inv_s3<B>(xr, xi);
}

// phase C for the two slots of last-stage butterfly A (slots A and A+16): increments into the TMEM words and the
// shared-memory copy.  lo / hi: the 16 TMEM words of register groups A/8 and A/8 + 2.
template <int A>
__device__ __forceinline__ void fin_pair(double (&xr)[32], double (&xi)[32], uint32_t (&lo)[16], uint32_t (&hi)[16], uint32_t* __restrict__ shp,
                                         int lane) {
  fft32_i2_fin<A>(xr, xi);
  constexpr int t = A & 7;
  uint32_t i0, i1;
  phaseC_slot<A>(xr, xi, i0, i1);
  lo[2 * t] += i0;
  lo[2 * t + 1] += i1;
  shp[32 * A + lane] = lo[2 * t];
  shp[32 * A + lane + 1024] = lo[2 * t + 1];
  phaseC_slot<A + 16>(xr, xi, i0, i1);
  hi[2 * t] += i0;
  hi[2 * t + 1] += i1;
  shp[32 * (A + 16) + lane] = hi[2 * t];
  shp[32 * (A + 16) + lane + 1024] = hi[2 * t + 1];
}
template <int H>
__device__ __forceinline__ void fin_half(double (&xr)[32], double (&xi)[32], uint32_t tacc, uint32_t* __restrict__ shp, int lane) {
  uint32_t lo[16], hi[16];
  tmem_ld16(tacc + 16 * H, lo);
  tmem_ld16(tacc + 16 * (H + 2), hi);
  fin_pair<8 * H + 0>(xr, xi, lo, hi, shp, lane);
  fin_pair<8 * H + 1>(xr, xi, lo, hi, shp, lane);
  fin_pair<8 * H + 2>(xr, xi, lo, hi, shp, lane);
  fin_pair<8 * H + 3>(xr, xi, lo, hi, shp, lane);
  fin_pair<8 * H + 4>(xr, xi, lo, hi, shp, lane);
  fin_pair<8 * H + 5>(xr, xi, lo, hi, shp, lane);
  fin_pair<8 * H + 6>(xr, xi, lo, hi, shp, lane);
  fin_pair<8 * H + 7>(xr, xi, lo, hi, shp, lane);
  tmem_st16(tacc + 16 * H, lo);
  tmem_st16(tacc + 16 * (H + 2), hi);
}

// Same launch geometry, shared-memory layout and hand-over protocol as kernels.cu::blind_rotate_kernel<S>.
// V: bit 0 = digits through the integer-to-double conversion unit
template <int S, int V>
__global__ void __launch_bounds__(64 * S, 1)
blind_rotate_fused_kernel(const c2* __restrict__ fbsk, const uint64_t* __restrict__ small, const uint64_t* __restrict__ luts,
                          const uint32_t* __restrict__ lut_idx, uint64_t* __restrict__ out, const int32_t* __restrict__ out_rows,
                          const c2* __restrict__ tabs_g, int count) {
  static_assert(2 * S <= 32, "64 TMEM columns per warp, 8 warps per lane quarter");
  extern __shared__ __align__(128) unsigned char smem[];
  c2* stage = reinterpret_cast<c2*>(smem);                                             // [2][2][1024]
  uint32_t* shadow_all = reinterpret_cast<uint32_t*>(smem + kGgswBytes);                // [S][2][2048], every polynomial at a multiple of 8 KiB
  double* plane_all = reinterpret_cast<double*>(smem + kGgswBytes + (size_t)S * 16384);  // [S][2][1024]
  c2* tab_f = reinterpret_cast<c2*>(smem + kGgswBytes + (size_t)S * 32768);              // [12][32]
  c2* tab_i = tab_f + kTabEntries * 32;                                                 // [12][32]
  uint16_t* at_all = reinterpret_cast<uint16_t*>(tab_i + kTabEntries * 32);             // [S][768]
  uint8_t* need = reinterpret_cast<uint8_t*>(at_all + S * 768);                         // [768]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(need + 768);                         // GGSW bytes have landed
  uint32_t* done_cnt = reinterpret_cast<uint32_t*>(full_bar + 1);                       // warps done with the stage
  uint32_t* tmem_slot = done_cnt + 1;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s = warp >> 1, w = warp & 1;
  const int sample = blockIdx.x * S + s;
  const bool active = sample < count;
  const uint32_t n_warps_active = 2u * (uint32_t)min(S, count - (int)blockIdx.x * S);

  for (int t = tid; t < 2 * kTabEntries * 32; t += 64 * S) tab_f[t] = tabs_g[t];
  if (tid == 0) {
    mbar_init(full_bar, 1);
    *done_cnt = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, kTmemCols);
  uint16_t* at = at_all + s * 768;
  {
    for (int t = w * 32 + lane; t < 768; t += 64) {
      uint32_t a = 0;
      if (active && t < kSmall) {
        const uint64_t x = small[(size_t)sample * kSmall + t];
        a = modswitch(x);
        if (t < kLweN) a = (a & 4095u) | ((x != 0 && (a & 4095u) != 0) ? 0x8000u : 0u);
      }
      at[t] = (uint16_t)a;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  for (int t = tid; t < 768; t += 64 * S) {
    uint32_t f = 0;
#pragma unroll
    for (int ss = 0; ss < S; ss++) f |= at_all[ss * 768 + t];
    need[t] = (t < kLweN && (f & 0x8000u)) ? 1 : 0;
  }
  __syncthreads();

  auto issue_ggsw = [&](int i) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive_expect_tx(full_bar, (uint32_t)kGgswBytes);
    const unsigned char* src = reinterpret_cast<const unsigned char*>(fbsk + (size_t)i * 4 * kHalfN);
#pragma unroll
    for (int c = 0; c < 4; c++) bulk_g2s(smem + c * (kGgswBytes / 4), src + c * (kGgswBytes / 4), kGgswBytes / 4, full_bar);
  };
  auto release_stage = [&](int i) {
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      if (atomicAdd(done_cnt, 1u) == n_warps_active - 1u) {
        *reinterpret_cast<volatile uint32_t*>(done_cnt) = 0u;
        int j = i + 1;
        while (j < kLweN && !need[j]) j++;
        if (j < kLweN) issue_ggsw(j);
      }
    }
  };
  if (tid == 0) {
    int j = 0;
    while (j < kLweN && !need[j]) j++;
    if (j < kLweN) issue_ggsw(j);
  }

  if (active) {
    const uint32_t shp_off = (uint32_t)kGgswBytes + (uint32_t)(s * 2 + w) * 8192u;   // byte offset of this polynomial's accumulator copy
    uint32_t* shp = shadow_all + (size_t)s * 2 * kN + (size_t)w * kN;
    double* plane = plane_all + (size_t)s * 2 * kHalfN;
    const int bar_id = 1 + s;
    const uint32_t tacc = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);

    {
      const uint64_t* lut = luts + (size_t)lut_idx[sample] * kN;
      const uint32_t rot = (4096u - (uint32_t)at[kLweN]) & 4095u;
#pragma unroll 1
      for (int g = 0; g < 4; g++) {
        uint32_t v[16];
#pragma unroll
        for (int t = 0; t < 8; t++) {
          const uint32_t j = 32u * (8 * g + t) + lane;
          v[2 * t] = (w == 0) ? 0u : (uint32_t)(rot_read(lut, j, rot) >> 32);
          v[2 * t + 1] = (w == 0) ? 0u : (uint32_t)(rot_read(lut, j + 1024u, rot) >> 32);
          shp[j] = v[2 * t];
          shp[j + 1024u] = v[2 * t + 1];
        }
        tmem_st16(tacc + 16 * g, v);
      }
      tmem_wait_st();
    }
    __syncwarp();

    double xr[32], xi[32];
    const int pp = lane >> 4, k1 = 16 * w + (lane & 15);
    const c2* b_own = stage + ((size_t)(pp * 2 + pp) * kHalfN + k1);
    const c2* b_in = stage + ((size_t)((1 - pp) * 2 + pp) * kHalfN + k1);
    uint32_t n_exec = 0;
    for (int i = 0; i < kLweN; i++) {
      if (!__any_sync(0xffffffffu, need[i] != 0)) continue;
      const uint32_t a = at[i];
      const uint32_t par = n_exec & 1u;
      n_exec++;
      if (!__any_sync(0xffffffffu, (a & 0x8000u) != 0)) {
        mbar_wait(full_bar, par);
        release_stage(i);
        continue;
      }
      // phase A + forward pass 1, interleaved
      phaseA_f1<(V & 1) != 0>(xr, xi, smem, shp_off, a & 4095u, lane);
      fwd_twiddle_inplace(xr, xi, tab_f, lane);
      col_store_brev(xr, plane + w * kHalfN, lane);
      bar_sync(bar_id, 64);
      row_load(xr, plane + pp * kHalfN, k1);
      bar_sync(bar_id, 64);
      col_store_brev(xi, plane + w * kHalfN, lane);
      bar_sync(bar_id, 64);
      row_load(xi, plane + pp * kHalfN, k1);
      bar_sync(bar_id, 64);
      // forward pass 2, Fourier MAC, inverse pass 1: block by block
      fft32_fwd_s12(xr, xi);
      mbar_wait(full_bar, par);
      mid_block<0>(xr, xi, b_own, b_in);
      mid_block<1>(xr, xi, b_own, b_in);
      mid_block<2>(xr, xi, b_own, b_in);
      mid_block<3>(xr, xi, b_own, b_in);
      release_stage(i);
      fft32_inv_s45(xr, xi);
      inv_twiddle_inplace(xr, xi, tab_i, k1);
      row_store(xr, plane + pp * kHalfN, k1);
      bar_sync(bar_id, 64);
      col_load_brev(xr, plane + w * kHalfN, lane);
      bar_sync(bar_id, 64);
      row_store(xi, plane + pp * kHalfN, k1);
      bar_sync(bar_id, 64);
      col_load_brev(xi, plane + w * kHalfN, lane);
      bar_sync(bar_id, 64);
      // inverse pass 2 + phase C, interleaved
      fft32_i2_head(xr, xi);
      fin_half<0>(xr, xi, tacc, shp, lane);
      fin_half<1>(xr, xi, tacc, shp, lane);
      tmem_wait_st();
      __syncwarp();
    }

    {
      const size_t row = out_rows ? (size_t)out_rows[sample] : (size_t)sample;
      uint64_t* o = out + row * kBig;
      if (w == 0) {
        for (int j = lane; j < kN; j += 32) {
          const uint32_t v = (j == 0) ? shp[0] : 0u - shp[kN - j];
          o[j] = (uint64_t)v << 32;
        }
      } else if (lane == 0) {
        o[kN] = (uint64_t)shp[0] << 32;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

template <int S, int V>
static cudaError_t launch_fused_s(const c2* fbsk, const uint64_t* small, const uint64_t* luts, const uint32_t* lut_idx, uint64_t* out,
                                  const int32_t* out_rows, const c2* tabs, int count, cudaStream_t st) {
  const size_t smem = br_smem_bytes(S);
  static PerDeviceOnce once;
  bool& configured = *once.slot();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(blind_rotate_fused_kernel<S, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  blind_rotate_fused_kernel<S, V><<<(count + S - 1) / S, 64 * S, smem, st>>>(fbsk, small, luts, lut_idx, out, out_rows, tabs, count);
  return cudaGetLastError();
}

cudaError_t launch_blind_rotate_fused(const c2* fbsk, const uint64_t* small, const uint64_t* luts, const uint32_t* lut_idx, uint64_t* out,
                                      const int32_t* out_rows, const c2* tabs, int count, int variant, cudaStream_t st) {
  if (count <= 0) return cudaSuccess;
  if (variant & 1) return launch_fused_s<4, 1>(fbsk, small, luts, lut_idx, out, out_rows, tabs, count, st);
  return launch_fused_s<4, 0>(fbsk, small, luts, lut_idx, out, out_rows, tabs, count, st);
}

}  // namespace fb
