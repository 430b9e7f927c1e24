// kernels.cu -- sm_100a kernels of the batched programmable bootstrap (KS -> BR -> SE) and glue.
//
// Replaces the tfhe-rs 0.2.0 arithmetic under fhe-regex's six smart_* call sites
// (/root/reference/src/regex/execution.rs:76,93,110,143,173,190) -- SURVEY.md section 2, kernels K1-K5, K7.
//   K1 (ks_kernels.cu)           exact mod-2^64 LWE x KSK gadget contraction on the int8 tensor cores
//   K2-K4 blind_rotate_kernel    modulus switch + LUT accumulator init, 742 CMUX steps with an f64
//                                 negacyclic FFT external product, sample extract -- one launch
//   K5 lincomb_kernel            sum_i c_i * ct_i + trivial(const)
//   K7 bsk_convert_kernel        standard-domain bootstrapping key -> Fourier domain (key load)
#include <cuda_runtime.h>
#include <stdint.h>
#include "br_core.cuh"
#include "kernels.h"
#include "ptx_sync.cuh"
#include "br_tmem.cuh"

namespace fb {

// ------------------------------------------------------------------------------------------------
// K7: key conversion.  One CTA (2 warps) per (i, row): the two polynomials of one GLWE row.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64, 1)
bsk_convert_kernel(const uint64_t* __restrict__ bsk_std, c2* __restrict__ fbsk, const c2* __restrict__ tabs_g) {
  extern __shared__ __align__(128) unsigned char smem[];
  double* plane = reinterpret_cast<double*>(smem);            // [2][kPlaneDoubles]
  c2* tab_f = reinterpret_cast<c2*>(plane + 2 * kPlaneDoubles);   // [12][32]
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  for (int t = tid; t < kTabEntries * 32; t += 64) tab_f[t] = tabs_g[t];
  __syncthreads();
  const size_t row = blockIdx.x;  // (i * 2 + r)
  double xr[32], xi[32];
  load_torus_poly(xr, xi, bsk_std + (row * 2 + w) * kN, lane);
  fft32_fwd_twist(xr, xi);
  fwd_twiddle_inplace(xr, xi, tab_f, lane);
  const int pp = lane >> 4, k1 = 16 * w + (lane & 15);
  col_store_brev(xr, plane + w * kPlaneDoubles, lane);
  __syncthreads();
  row_load(xr, plane + pp * kPlaneDoubles, k1);
  __syncthreads();
  col_store_brev(xi, plane + w * kPlaneDoubles, lane);
  __syncthreads();
  row_load(xi, plane + pp * kPlaneDoubles, k1);
  fft32_fwd(xr, xi);
  c2* dst = fbsk + (row * 2 + pp) * kHalfN + k1;
#pragma unroll
  for (int q = 0; q < 32; q++) {
    c2 v;
    v.x = xr[q];
    v.y = xi[q];
    dst[32 * brev5(q)] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// K2-K4: blind rotation.  One CTA per SM: S samples, 2 warps each (warp w owns polynomial w in the
// coefficient-domain phases and rows 16w..16w+15 of both polynomials in the frequency-domain phase).
// The Fourier GGSW of CMUX step i (64 KiB, contiguous in HBM/L2) is streamed into one shared-memory
// stage with cp.async.bulk (TMA bulk copy) and consumed by all samples of the CTA; the warp that is
// last to finish its Fourier MAC of step i (shared-memory arrival counter) issues the copy of step
// i+1, which lands while the samples run their inverse transforms and the next forward transforms;
// an mbarrier with a transaction count tells the consumers when the bytes are there.
// The accumulator is kept on 32 torus bits (br_core.cuh): every thread owns its 64 coefficients in tensor
// memory (TMEM, tcgen05.ld/st) and mirrors them into shared memory for the rotation reads of the
// decomposition.  No FP64<->integer conversion goes through the conversion unit (magic-number rounding).
//
// Shared memory: 64 KiB GGSW stage + S * (16 KiB u32 accumulator + 16 KiB transpose plane)
//                + 12 KiB twiddles + S * 1.5 KiB mod-switched mask + 768 B step flags + mbarrier.
// ------------------------------------------------------------------------------------------------


template <int S>
__global__ void __launch_bounds__(64 * S, 1)
blind_rotate_kernel(const c2* __restrict__ fbsk, const uint64_t* __restrict__ small, const uint64_t* __restrict__ luts,
                    const uint32_t* __restrict__ lut_idx, uint64_t* __restrict__ out, const int32_t* __restrict__ out_rows,
                    const c2* __restrict__ tabs_g, int count) {
  static_assert(2 * S <= 32, "64 TMEM columns per warp, 8 warps per lane quarter");
  extern __shared__ __align__(128) unsigned char smem[];
  c2* stage = reinterpret_cast<c2*>(smem);                                             // [2][2][1024]
  uint32_t* shadow_all = reinterpret_cast<uint32_t*>(smem + kGgswBytes);                // [S][2][2048]
  double* plane_all = reinterpret_cast<double*>(smem + kGgswBytes + (size_t)S * 16384);  // [S][2][kPlaneDoubles]
  c2* tab_f = reinterpret_cast<c2*>(smem + kGgswBytes + (size_t)S * (16384 + 2 * kPlaneDoubles * sizeof(double)));   // [12][32]
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
    mbar_init(full_bar, 1);                // one arrive.expect_tx per staged GGSW (+ its bytes)
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
    need[t] = (t < kLweN && (f & 0x8000u)) ? 1 : 0;  // some sample of this CTA rotates at step t
  }
  __syncthreads();

  // stage hand-over: issue the bulk copy of the GGSW of step i (called by exactly one thread)
  auto issue_ggsw = [&](int i) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive_expect_tx(full_bar, (uint32_t)kGgswBytes);
    const unsigned char* src = reinterpret_cast<const unsigned char*>(fbsk + (size_t)i * 4 * kHalfN);
#pragma unroll
    for (int c = 0; c < 4; c++) bulk_g2s(smem + c * (kGgswBytes / 4), src + c * (kGgswBytes / 4), kGgswBytes / 4, full_bar);
  };
  // a warp is done reading the stage for step i; the last one to say so starts the copy of the next needed step
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
    uint32_t* shp = shadow_all + (size_t)s * 2 * kN + (size_t)w * kN;  // shared copy of polynomial w of this sample
    double* plane = plane_all + (size_t)s * 2 * kPlaneDoubles;
    const int bar_id = 1 + s;
    // this warp's 64 TMEM columns: coefficient 32r+lane in column 2r, 32r+lane+1024 in column 2r+1
    const uint32_t tacc = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);

    // accumulator init: (0, lut * X^{-b}), top words
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
    // the samples of the CTA start their rotation together (their accumulator reads above have different latencies, and
    // samples that start apart stay apart: br_fused.cu, "br_sync")
    if (S > 1) bar_sync(15, 32 * (int)n_warps_active);

    double xr[32], xi[32];
    const int pp = lane >> 4, k1 = 16 * w + (lane & 15);
    const c2* b_own = stage + ((size_t)(pp * 2 + pp) * kHalfN + k1);        // GGSW[row pp][column pp]
    const c2* b_in = stage + ((size_t)((1 - pp) * 2 + pp) * kHalfN + k1);   // GGSW[row 1-pp][column pp]
    uint32_t n_exec = 0;
    for (int i = 0; i < kLweN; i++) {
      // CTA-uniform: zero mask element (trivial inputs) or X^0 for every sample.  The votes only tell the
      // compiler what is true anyway: these branches are warp-uniform.
      if (!__any_sync(0xffffffffu, need[i] != 0)) continue;
      const uint32_t a = at[i];
      const uint32_t par = n_exec & 1u;
      n_exec++;
      if (!__any_sync(0xffffffffu, (a & 0x8000u) != 0)) {   // this sample skips the step, but still takes part in the hand-over
        mbar_wait(full_bar, par);
        release_stage(i);
        continue;
      }
      // phase A: this warp's polynomial -> digits -> folded/twisted -> pass 1 -> twiddle -> transpose (re, im)
      phaseA_load32(xr, xi, shp, a & 4095u, lane);
      fft32_fwd_twist(xr, xi);
      fwd_twiddle_inplace(xr, xi, tab_f, lane);
      col_store_brev(xr, plane + w * kPlaneDoubles, lane);   // the plane is free: see the barrier after the last loads
      bar_sync(bar_id, 64);
      row_load(xr, plane + pp * kPlaneDoubles, k1);
      bar_sync(bar_id, 64);
      col_store_brev(xi, plane + w * kPlaneDoubles, lane);
      bar_sync(bar_id, 64);
      row_load(xi, plane + pp * kPlaneDoubles, k1);
      bar_sync(bar_id, 64);                           // both warps have read: the next transpose may store at once
      // phase B: pass 2 -> Fourier MAC with the staged GGSW_i -> inverse pass 1 -> twiddle -> transpose
      fft32_fwd(xr, xi);
      mbar_wait(full_bar, par);
#pragma unroll
      for (int q = 0; q < 32; q++) {
        const int k2 = brev5(q);
        const double pr = __shfl_xor_sync(0xffffffffu, xr[q], 16);
        const double pi = __shfl_xor_sync(0xffffffffu, xi[q], 16);
        mac_point2(xr[q], xi[q], pr, pi, b_own[32 * k2], b_in[32 * k2]);
      }
      release_stage(i);                           // this warp no longer reads the stage
      fft32_inv(xr, xi);
      inv_twiddle_inplace(xr, xi, tab_i, k1);
      row_store(xr, plane + pp * kPlaneDoubles, k1);
      bar_sync(bar_id, 64);
      col_load_brev(xr, plane + w * kPlaneDoubles, lane);
      bar_sync(bar_id, 64);
      row_store(xi, plane + pp * kPlaneDoubles, k1);
      bar_sync(bar_id, 64);
      col_load_brev(xi, plane + w * kPlaneDoubles, lane);
      bar_sync(bar_id, 64);
      // phase C: inverse pass 2 -> untwist, round to the 32-bit torus, accumulate (TMEM copy + shared copy)
      fft32_inv(xr, xi);
#pragma unroll
      for (int g = 0; g < 4; g++) {
        uint32_t v[16];
        tmem_ld16(tacc + 16 * g, v);
#pragma unroll
        for (int t = 0; t < 8; t++) {
          const int r = 8 * g + t;
          uint32_t inc0, inc1;
          phaseC_increments32(xr, xi, r, inc0, inc1);
          v[2 * t] += inc0;
          v[2 * t + 1] += inc1;
          shp[32 * r + lane] = v[2 * t];
          shp[32 * r + lane + 1024] = v[2 * t + 1];
        }
        tmem_st16(tacc + 16 * g, v);
      }
      tmem_wait_st();
      __syncwarp();
    }

    // K4: sample extract of the constant coefficient: mask_0 = a_0, mask_j = -a_{N-j}; body = b_0
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

// ------------------------------------------------------------------------------------------------
// K5: linear glue.  out[row_o] = sum_t coef[t] * arena[row_t] + (0,...,0, body_const[o])
// terms of output o are term_off[o] .. term_off[o+1]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
lincomb_kernel(uint64_t* __restrict__ arena, const int32_t* __restrict__ out_rows, const int32_t* __restrict__ term_off,
               const int32_t* __restrict__ term_rows, const int64_t* __restrict__ term_coef,
               const uint64_t* __restrict__ body_const, int n_out) {
  const int o = blockIdx.x;
  if (o >= n_out) return;
  const int t0 = term_off[o], t1 = term_off[o + 1];
  uint64_t* dst = arena + (size_t)out_rows[o] * kBig;
  for (int j = threadIdx.x; j < kBig; j += 256) {
    uint64_t v = (j == kN) ? body_const[o] : 0ull;
    for (int t = t0; t < t1; t++) v += (uint64_t)term_coef[t] * arena[(size_t)term_rows[t] * kBig + j];
    dst[j] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// FP64 FMA pipe probe: the roofline denominator of the blind rotation (MEASURED_PEAKS.json carries
// HBM and bf16 figures only).  8 independent DFMA chains per thread, 2 flop per DFMA.
// ------------------------------------------------------------------------------------------------
constexpr int PEAK_ITERS = 4096;
__global__ void __launch_bounds__(256)
fp64_peak_kernel(double* __restrict__ sink, double seed) {
  double a[8];
#pragma unroll
  for (int k = 0; k < 8; k++) a[k] = seed + (double)(threadIdx.x + k);
  const double m = 1.0000001, c = 1e-9;
  for (int it = 0; it < PEAK_ITERS; it++) {
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = __fma_rn(a[k], m, c);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += a[k];
  if (s == 12345.678) sink[0] = s;  // never true; keeps the chains alive
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
cudaError_t launch_fp64_peak(double* sink, int ctas, cudaStream_t st) {
  fp64_peak_kernel<<<ctas, 256, 0, st>>>(sink, 0.5);
  return cudaGetLastError();
}
double fp64_peak_flops_per_launch(int ctas) { return (double)ctas * 256.0 * 8.0 * 2.0 * (double)PEAK_ITERS; }
int br_samples_per_cta() { return 4; }

size_t br_smem_bytes(int S) {
  return (size_t)kGgswBytes + (size_t)S * (16384 + 2 * kPlaneDoubles * sizeof(double)) + 2 * kTabEntries * 32 * sizeof(c2) + (size_t)S * 768 * sizeof(uint16_t) + 768 + 16 + 32;
}

cudaError_t launch_bsk_convert(const uint64_t* bsk_std, c2* fbsk, const c2* tabs, cudaStream_t st) {
  const size_t smem = 2 * kPlaneDoubles * sizeof(double) + kTabEntries * 32 * sizeof(c2);
  cudaError_t e = cudaFuncSetAttribute(bsk_convert_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  bsk_convert_kernel<<<kLweN * 2, 64, smem, st>>>(bsk_std, fbsk, tabs);
  return cudaGetLastError();
}

template <int S>
static cudaError_t launch_br_s(const c2* fbsk, const uint64_t* small, const uint64_t* luts, const uint32_t* lut_idx,
                               uint64_t* out, const int32_t* out_rows, const c2* tabs, int count, cudaStream_t st) {
  const size_t smem = br_smem_bytes(S);
  static PerDeviceOnce once;
  bool& configured = *once.slot();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(blind_rotate_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  blind_rotate_kernel<S><<<(count + S - 1) / S, 64 * S, smem, st>>>(fbsk, small, luts, lut_idx, out, out_rows, tabs, count);
  return cudaGetLastError();
}

cudaError_t launch_blind_rotate(const c2* fbsk, const uint64_t* small, const uint64_t* luts, const uint32_t* lut_idx,
                                uint64_t* out, const int32_t* out_rows, const c2* tabs, int count, cudaStream_t st) {
  if (count <= 0) return cudaSuccess;
  // Throughput wants 4 samples per SM; a batch that does not fill the GPU at that width (the narrow DAG
  // levels of a regex match) finishes sooner with fewer samples contending for each SM.
  static int sms_of[64] = {};   // per device: a process may hold contexts on several GPUs
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  int& sms = sms_of[dev & 63];
  if (sms == 0 && (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)) sms = 148;
  // measured (tools/wave_times.py): one sample per SM on all 148 SMs takes 6.6 ms, two per SM on 74-148 SMs 5.9 ms,
  // one per SM on <= 74 SMs 4.9-5.7 ms -- a lone sample per SM is only worth it while half the SMs stay idle
  if (2 * count <= sms) return launch_br_s<1>(fbsk, small, luts, lut_idx, out, out_rows, tabs, count, st);
  if (count <= 2 * sms) return launch_br_s<2>(fbsk, small, luts, lut_idx, out, out_rows, tabs, count, st);
  if (count <= 3 * sms) return launch_br_s<3>(fbsk, small, luts, lut_idx, out, out_rows, tabs, count, st);
  return launch_br_s<4>(fbsk, small, luts, lut_idx, out, out_rows, tabs, count, st);
}

cudaError_t launch_lincomb(uint64_t* arena, const int32_t* out_rows, const int32_t* term_off, const int32_t* term_rows,
                           const int64_t* term_coef, const uint64_t* body_const, int n_out, cudaStream_t st) {
  if (n_out <= 0) return cudaSuccess;
  lincomb_kernel<<<n_out, 256, 0, st>>>(arena, out_rows, term_off, term_rows, term_coef, body_const, n_out);
  return cudaGetLastError();
}

}  // namespace fb
