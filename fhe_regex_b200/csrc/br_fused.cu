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
// Shared-memory layouts of the transposes (template bits of V; all bit-identical):
//   one plane for both components, one after the other (V & 72 == 0, 4 PBS per CTA);
//   planes inside the accumulator copies (S > 4: 6 PBS per CTA; br_core.cuh "planes aliased");
//   a plane per component, the real one inside the accumulator copy (V & 8, the default: one barrier per transpose);
//   planes inside the accumulator copies + full inter-pass twiddle tables built once per launch (V & 64).
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

// The same block with the key served from tensor memory: the thread's 64 key values of the step sit in 256 columns of its own
// TMEM lane in MAC order (slot q: own at columns 8q..8q+3, in at 8q+4..8q+7; layout fbsk_lm_kernel), four slots per
// tcgen05.ld -- no shared-memory wavefronts for the key (tools/microbench/tmem_ld_probe.cu: > 700 B/cycle/SM).
template <int H>
__device__ __forceinline__ void mac_half_tmem(double (&xr)[32], double (&xi)[32], const uint32_t (&kv)[32]) {
#pragma unroll
  for (int t = 0; t < 4; t++) {
    const int q = 4 * H + t;
    const double pr = __shfl_xor_sync(0xffffffffu, xr[q], 16);
    const double pi = __shfl_xor_sync(0xffffffffu, xi[q], 16);
    c2 g_own, g_in;
    g_own.x = __hiloint2double((int)kv[8 * t + 1], (int)kv[8 * t]);
    g_own.y = __hiloint2double((int)kv[8 * t + 3], (int)kv[8 * t + 2]);
    g_in.x = __hiloint2double((int)kv[8 * t + 5], (int)kv[8 * t + 4]);
    g_in.y = __hiloint2double((int)kv[8 * t + 7], (int)kv[8 * t + 6]);
    mac_point2(xr[q], xi[q], pr, pi, g_own, g_in);
  }
}
// block B of the middle section, software-pipelined on the tensor-memory reads: the key values of a half (4 slots, 32
// columns) are requested one half ahead, so that a tcgen05.wait::ld finds them there.  ka / kb alternate; on entry the load
// of half 2B into ka is in flight, on exit the load of half 2B + 2 (if any).
template <int B>
__device__ __forceinline__ void mid_block_tmem(double (&xr)[32], double (&xi)[32], uint32_t tkey, uint32_t (&ka)[32], uint32_t (&kb)[32]) {
  fft32_fwd_s3<B>(xr, xi);
  fft32_fwd_s45<2 * B>(xr, xi);
  tmem_ld32_wait(ka);
  tmem_ld32_issue(tkey + 32 * (2 * B + 1), kb);
  mac_half_tmem<2 * B>(xr, xi, ka);
  fft32_fwd_s45<2 * B + 1>(xr, xi);
  tmem_ld32_wait(kb);
  if (B < 3) tmem_ld32_issue(tkey + 32 * (2 * B + 2), ka);
  mac_half_tmem<2 * B + 1>(xr, xi, kb);
  fft32_inv_s12<2 * B>(xr, xi);
  fft32_inv_s12<2 * B + 1>(xr, xi);
  fft32_inv_s3<B>(xr, xi);
}

// the Fourier key re-ordered for the tensor-memory path: per step 64 chunks of 64 rows of one complex value;
// chunk j = 2q + o (MAC slot q, o = 0: own column, 1: the other row's), row r = 32 w + L (warp parity w, lane L)
__global__ void __launch_bounds__(256)
fbsk_lm_kernel(const c2* __restrict__ fbsk, c2* __restrict__ fbsk_lm) {
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;      // (i, j, r)
  if (idx >= (size_t)kLweN * 4096) return;
  const int r = (int)(idx & 63), j = (int)((idx >> 6) & 63), i = (int)(idx >> 12);
  const int w = r >> 5, L = r & 31, pp = L >> 4, k1 = 16 * w + (L & 15);
  const int q = j >> 1, o = j & 1;
  const int k = k1 + 32 * brev5(q);
  const int pin = o ? 1 - pp : pp;
  fbsk_lm[idx] = fbsk[fbsk_index(i, pin, pp, k)];
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
// V: bit 0 = digits through the integer-to-double conversion unit; bit 1 = Fourier key of the step staged in tensor memory
// (fbsk then points to the re-ordered key of fbsk_lm_kernel)
template <int S, int V>
__global__ void __launch_bounds__(64 * S, 1)
blind_rotate_fused_kernel(const c2* __restrict__ fbsk, const uint64_t* __restrict__ small, const uint64_t* __restrict__ luts,
                          const uint32_t* __restrict__ lut_idx, uint64_t* __restrict__ out, const int32_t* __restrict__ out_rows,
                          const c2* __restrict__ tabs_g, int count, int stagger) {
  static_assert(2 * S <= 32, "64 TMEM columns per warp, 8 warps per lane quarter");
  // DUAL (V & 8, 4 PBS per CTA): a plane per component -- the real plane inside the accumulator copy, the imaginary plane in a
  // buffer of its own; one barrier per transpose.  AL: the (real) transpose planes live inside the accumulator copies
  // (br_core.cuh, "planes aliased"): more than 4 PBS per CTA, or DUAL.
  constexpr bool DUAL = (V & 8) != 0;
  constexpr bool FULLTAB = (V & 64) != 0;   // full inter-pass twiddle tables in shared memory (built once per launch)
  constexpr bool AL = S > 4 || DUAL || FULLTAB;
  static_assert(!(FULLTAB && DUAL), "no shared memory for both");
  static_assert(!(S > 4 && (V & 2)), "the tensor-memory key uses the columns of the third group of accumulators");
  static_assert(!DUAL || S <= 4, "no shared memory for imaginary planes beyond 4 PBS per CTA");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // the block starts at the next 8 KiB boundary of the shared-memory window (the accumulator copies must sit at absolute
  // multiples of 8 KiB: br_core.cuh::phaseA_slot); the launch reserves 8 KiB for this
  const uint32_t pad = (8192u - (smem_u32(smem_raw) & 8191u)) & 8191u;
  unsigned char* smem = smem_raw + pad;
  c2* stage = reinterpret_cast<c2*>(smem);                                             // [2][2][1024]
  uint32_t* shadow_all = reinterpret_cast<uint32_t*>(smem + kGgswBytes);                // [S][2][2048], every polynomial at a multiple of 8 KiB
  double* plane_all = reinterpret_cast<double*>(smem + kGgswBytes + (size_t)S * 16384);  // [S][2][kPlaneDoubles]; AL: [S][2] overflow blocks
  constexpr size_t kPlaneBytes = AL ? (size_t)kPlaneOvfBytes : kPlaneDoubles * sizeof(double);
  double* implane_all = reinterpret_cast<double*>(smem + kGgswBytes + (size_t)S * (16384 + 2 * kPlaneBytes));   // DUAL: [S][2][kPlaneDoubles]
  constexpr size_t kImBytes = DUAL ? kPlaneDoubles * sizeof(double) : 0;
  c2* tab_f = reinterpret_cast<c2*>(smem + kGgswBytes + (size_t)S * (16384 + 2 * kPlaneBytes + 2 * kImBytes));   // [12][32]
  constexpr int kTabRows = FULLTAB ? 32 : kTabEntries;
  c2* tab_i = tab_f + kTabRows * 32;                                                    // [12][32]; FULLTAB: both [32][32]
  // the small arrays go into the alignment gap in front of the block when it is large enough (DUAL needs the room: the block
  // plus a worst-case gap would not fit the 227 KiB of a CTA otherwise)
  constexpr uint32_t kTailBytes = (uint32_t)S * 768u * 2u + 768u + 64u;
  unsigned char* tail = (DUAL && pad >= kTailBytes) ? smem_raw : reinterpret_cast<unsigned char*>(tab_i + kTabRows * 32);
  uint16_t* at_all = reinterpret_cast<uint16_t*>(tail);                                 // [S][768]
  uint8_t* need = reinterpret_cast<uint8_t*>(at_all + S * 768);                         // [768]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(need + 768);                         // GGSW bytes have landed
  uint64_t* tkey_bar = full_bar + 1;                                                    // (V & 2) GGSW copied into tensor memory
  uint64_t* macdone_bar = tkey_bar + 1;                                                 // (V & 2) every warp is past its MAC of the step
  uint32_t* done_cnt = reinterpret_cast<uint32_t*>(macdone_bar + 1);                    // warps done with the stage
  uint32_t* tmem_slot = done_cnt + 1;
  constexpr bool kTKey = (V & 2) != 0;
  constexpr bool kRedundantBarriers = (V & 4) != 0;   // the two barriers per step that round 2 found unnecessary (A/B only)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s = warp >> 1, w = warp & 1;
  const int sample = blockIdx.x * S + s;
  const bool active = sample < count;
  const uint32_t n_warps_active = 2u * (uint32_t)min(S, count - (int)blockIdx.x * S);

  if (FULLTAB) {
    for (int t = tid; t < 1024; t += 64 * S) {
      tab_f[t] = fb_full_twiddle(tabs_g, t >> 5, t & 31);
      tab_i[t] = fb_full_twiddle(tabs_g + kTabEntries * 32, t >> 5, t & 31);
    }
  } else {
    for (int t = tid; t < 2 * kTabEntries * 32; t += 64 * S) tab_f[t] = tabs_g[t];
  }
  if (tid == 0) {
    mbar_init(full_bar, 1);
    mbar_init(tkey_bar, n_warps_active > 0 ? n_warps_active : 1);   // one tcgen05.commit per active warp and step
    mbar_init(macdone_bar, n_warps_active > 0 ? n_warps_active : 1);
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
  // ---- tensor-memory key (V & 2) ----------------------------------------------------------------------------------
  // e = number of steps this CTA has executed so far.  The key of executed step e travels: HBM -> (bulk copy #e) -> the
  // shared-memory stage -> (tcgen05.cp, 64 chunks spread over the active warps' lane 0) -> TMEM columns [128, 384) of every lane.
  //   MAC of step e:      wait tkey_bar phase e; thread 0 then issues bulk copy #e+1 (the stage is free: the copies into TMEM
  //                       that read it have completed); after its MAC a warp bumps done_cnt.
  //   later in step e:    every warp waits until all warps are past their MAC (done_cnt) and bulk copy #e+1 has landed, its
  //                       lane 0 issues its share of the tcgen05.cp of key e+1 and commits to tkey_bar.
  const uint32_t tkey = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 128u;
  auto next_needed = [&](int i) {
    int j = i + 1;
    while (j < kLweN && !need[j]) j++;
    return j;
  };
  const uint32_t uwarp = (uint32_t)__shfl_sync(0xffffffffu, warp, 0);   // provably warp-uniform: the copy operands stay in uniform registers
  auto stage_key_to_tmem = [&](uint32_t e) {   // copies of executed step e's key, issued at the top of that step; warp-uniform
    if (e > 0) mbar_wait(macdone_bar, (e - 1u) & 1u);   // every warp is past its MAC of step e - 1 (hardware sleep, no spinning:
                                                        // a spinning warp would take issue slots from the warp it waits for)
    mbar_wait(full_bar, e & 1u);
    tmem_fence_after();
    if (lane == 0) {
      const uint64_t d0 = tmem_cp_desc(smem_u32(smem));
      if (n_warps_active == 8u) {
#pragma unroll
        for (uint32_t u = 0; u < 8u; u++) {
          const uint32_t j = uwarp + 8u * u;
          tmem_cp_64x128b_02_13(tmem_base + 128u + 4u * j, d0 + 64u * j);
        }
      } else {
        for (uint32_t j = uwarp; j < 64u; j += n_warps_active) tmem_cp_64x128b_02_13(tmem_base + 128u + 4u * j, d0 + 64u * j);
      }
      tmem_commit(tkey_bar);
    }
    __syncwarp();
  };
  auto mac_done = [&]() {
    tmem_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(macdone_bar);
  };
  const int first_step = next_needed(-1);
  if (tid == 0 && first_step < kLweN) issue_ggsw(first_step);

  if (active) {
    const uint32_t shp_off = (uint32_t)kGgswBytes + (uint32_t)(s * 2 + w) * 8192u;   // byte offset of this polynomial's accumulator copy
    uint32_t* shp = shadow_all + (size_t)s * 2 * kN + (size_t)w * kN;
    double* plane = plane_all + (size_t)s * 2 * kPlaneDoubles;   // (!AL)
    // AL: column side = this warp's own polynomial (main + overflow, lane folded in), row side = row k1 of polynomial pp
    unsigned char* ovf_s = reinterpret_cast<unsigned char*>(plane_all) + (size_t)s * 2 * kPlaneOvfBytes + kPlaneOvfLead;
    double* col_main = reinterpret_cast<double*>(smem + shp_off) + (tid & 31);
    double* col_ovf = reinterpret_cast<double*>(ovf_s + (size_t)w * kPlaneOvfBytes) + (tid & 31);
    double* row_al = nullptr;
    double* im_col = implane_all + (size_t)(s * 2 + w) * kPlaneDoubles + (tid & 31);   // DUAL
    double* im_row = nullptr;
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
    // Every sample of the CTA starts its rotation together (stagger bit 25, "br_sync" 1): the warps that own a body polynomial
    // have just read their accumulator from global memory, each with its own latency -- with 36 accumulators in random order
    // the samples of a CTA left the initialisation up to a few thousand cycles apart -- and they meet again every
    // `resync` CMUX steps (stagger bits 26-31, "br_resync"): nothing else brings samples that drift apart back into step, and
    // lock-step is worth several per cent (profiles/r02_experiments.md section 11); a barrier at EVERY step costs more than it
    // gains (the slowest warp of each step sets the pace).
    const bool sync_start = ((stagger >> 25) & 1) != 0;
    const uint32_t resync = ((uint32_t)stagger >> 26) & 63u;
    uint32_t resync_cnt = 0;
    if (sync_start) bar_sync(15, 32 * (int)n_warps_active);

    // optional start skew between the samples of a CTA (cycles per sample index): a few hundred cycles keep every warp
    // inside the same instruction-cache window while one sample's shared-memory stores fall into another's arithmetic
    // stagger bit 24 ("br_stagger_groups"): the skew goes to the odd samples only.  Warps 2s, 2s+1 of sample s sit on schedulers
    // 0, 1 (s even) or 2, 3 (s odd), so every scheduler still runs ONE instruction stream (its warps in lock-step) and the SM
    // two: the even samples' shared-memory bursts (stores, shuffles: pipes shared by the whole SM) fall into the odd samples'
    // FP64 phases (pipes owned by a scheduler).
    {
      const int cyc = stagger & 0xffffff;
      const int mult = (stagger >> 24) ? (s & 1) : s;
      if (cyc > 0 && mult > 0) {
        const long long t0 = clock64();
        while (clock64() - t0 < (long long)cyc * mult) {}
      }
    }
    double xr[32], xi[32];
    const int pp = lane >> 4, k1 = 16 * w + (lane & 15);
    if (AL)
      row_al = plane_al_row(reinterpret_cast<double*>(smem + kGgswBytes + (size_t)(s * 2 + pp) * 8192u),
                            reinterpret_cast<double*>(ovf_s + (size_t)pp * kPlaneOvfBytes), k1);
    if (DUAL) im_row = implane_all + (size_t)(s * 2 + pp) * kPlaneDoubles + (size_t)k1 * kPlaneRow;
    const c2* b_own = stage + ((size_t)(pp * 2 + pp) * kHalfN + k1);
    const c2* b_in = stage + ((size_t)((1 - pp) * 2 + pp) * kHalfN + k1);
    uint32_t n_exec = 0;
    for (int i = 0; i < kLweN; i++) {
      if (!__any_sync(0xffffffffu, need[i] != 0)) continue;
      const uint32_t a = at[i];
      const uint32_t par = n_exec & 1u;
      if (resync != 0 && ++resync_cnt == resync) {
        resync_cnt = 0;
        bar_sync(15, 32 * (int)n_warps_active);
      }
      if (kTKey) stage_key_to_tmem(n_exec);   // this step's key: shared-memory stage -> tensor memory (needed ~7k cycles from here)
      n_exec++;
      if (!__any_sync(0xffffffffu, (a & 0x8000u) != 0)) {   // this sample skips the step but takes part in the hand-over
        if (kTKey) {
          const int nj = next_needed(i);
          mbar_wait(tkey_bar, par);
          if (tid == 0 && nj < kLweN) issue_ggsw(nj);
          mac_done();
        } else {
          mbar_wait(full_bar, par);
          release_stage(i);
        }
        continue;
      }
      // phase A + forward pass 1, interleaved
      phaseA_f1<(V & 1) != 0>(xr, xi, smem, shp_off, a & 4095u, lane);
      // Transposes (barriers of two warps).  Column side: warp w touches the plane(s) of polynomial w only; row side: thread
      // (pp, k1) touches row k1 of the plane(s) of polynomial pp only, in the forward and in the inverse direction.  So no
      // barrier is needed behind the last row load (the next access to the planes is this thread's own row store) nor behind
      // the last column load of the step (the next is this warp's own column store -- or, with the plane inside the
      // accumulator copy, its own phase C; the other warp's next access sits behind the next step's first barrier).
      if (DUAL) {
        fwd_twiddle_col_store(xr, xi, tab_f, lane, col_main, col_ovf, im_col);
        bar_sync(bar_id, 64);
        row_load(xr, row_al, 0);
        row_load(xi, im_row, 0);
      } else {
        if (FULLTAB) fwd_twiddle_full(xr, xi, tab_f + lane);
        else fwd_twiddle_inplace(xr, xi, tab_f, lane);
        if (AL) {
          // the plane of polynomial w overwrites the accumulator copy of polynomial w, which only this warp read (phase A above)
          col_store_brev_al(xr, col_main, col_ovf);
          bar_sync(bar_id, 64);
          row_load(xr, row_al, 0);
          bar_sync(bar_id, 64);
          col_store_brev_al(xi, col_main, col_ovf);
          bar_sync(bar_id, 64);
          row_load(xi, row_al, 0);
        } else {
          col_store_brev(xr, plane + w * kPlaneDoubles, lane);
          bar_sync(bar_id, 64);
          row_load(xr, plane + pp * kPlaneDoubles, k1);
          bar_sync(bar_id, 64);
          col_store_brev(xi, plane + w * kPlaneDoubles, lane);
          bar_sync(bar_id, 64);
          row_load(xi, plane + pp * kPlaneDoubles, k1);
        }
        if (kRedundantBarriers) bar_sync(bar_id, 64);
      }
      // forward pass 2, Fourier MAC, inverse pass 1: block by block
      fft32_fwd_s12(xr, xi);
      int nj = kLweN;
      if (kTKey) {
        nj = next_needed(i);
        mbar_wait(tkey_bar, par);
        tmem_fence_after();
        if (tid == 0 && nj < kLweN) issue_ggsw(nj);
        uint32_t ka[32], kb[32];
        tmem_ld32_issue(tkey, ka);
        mid_block_tmem<0>(xr, xi, tkey, ka, kb);
        mid_block_tmem<1>(xr, xi, tkey, ka, kb);
        mid_block_tmem<2>(xr, xi, tkey, ka, kb);
        mid_block_tmem<3>(xr, xi, tkey, ka, kb);
        mac_done();
      } else {
        mbar_wait(full_bar, par);
        mid_block<0>(xr, xi, b_own, b_in);
        mid_block<1>(xr, xi, b_own, b_in);
        mid_block<2>(xr, xi, b_own, b_in);
        mid_block<3>(xr, xi, b_own, b_in);
        release_stage(i);
      }
      fft32_inv_s45(xr, xi);
      if (DUAL) {
        inv_twiddle_row_store(xr, xi, tab_i, k1, row_al, im_row);
        bar_sync(bar_id, 64);
        col_load_brev_al(xr, col_main, col_ovf);
        col_load_brev(xi, im_col, 0);
      } else {
        if (FULLTAB) inv_twiddle_full(xr, xi, tab_i + k1);
        else inv_twiddle_inplace(xr, xi, tab_i, k1);
        if (AL) {
          row_store(xr, row_al, 0);
          bar_sync(bar_id, 64);
          col_load_brev_al(xr, col_main, col_ovf);
          bar_sync(bar_id, 64);
          row_store(xi, row_al, 0);
          bar_sync(bar_id, 64);
          col_load_brev_al(xi, col_main, col_ovf);
        } else {
          row_store(xr, plane + pp * kPlaneDoubles, k1);
          bar_sync(bar_id, 64);
          col_load_brev(xr, plane + w * kPlaneDoubles, lane);
          bar_sync(bar_id, 64);
          row_store(xi, plane + pp * kPlaneDoubles, k1);
          bar_sync(bar_id, 64);
          col_load_brev(xi, plane + w * kPlaneDoubles, lane);
        }
        if (kRedundantBarriers) bar_sync(bar_id, 64);
      }
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

// shared memory of blind_rotate_fused_kernel<S, V> (without the 8 KiB of alignment slack)
static size_t fused_smem_bytes(int S, bool dual, bool fulltab = false) {
  if (fulltab)
    return (size_t)kGgswBytes + (size_t)S * (16384 + 2 * kPlaneOvfBytes) + 2 * 32 * 32 * sizeof(c2) + (size_t)S * 768 * sizeof(uint16_t) + 768 + 16 + 32;
  if (dual)   // the gap in front of the block takes the small arrays when it can: worst case = a gap one byte too small for them
    // (the caller adds 8192: block + small arrays + a gap of up to their size, 232 064 bytes at S = 4)
    return (size_t)kGgswBytes + (size_t)S * (16384 + 2 * kPlaneOvfBytes + 2 * kPlaneDoubles * sizeof(double)) + 2 * kTabEntries * 32 * sizeof(c2) +
           2 * ((size_t)S * 768 * sizeof(uint16_t) + 768 + 64) - 8192;
  if (S <= 4) return br_smem_bytes(S);
  return (size_t)kGgswBytes + (size_t)S * (16384 + 2 * kPlaneOvfBytes) + 2 * kTabEntries * 32 * sizeof(c2) + (size_t)S * 768 * sizeof(uint16_t) + 768 + 16 + 32;
}

template <int S, int V>
static cudaError_t launch_fused_s(const c2* fbsk, const uint64_t* small, const uint64_t* luts, const uint32_t* lut_idx, uint64_t* out,
                                  const int32_t* out_rows, const c2* tabs, int count, int stagger, cudaStream_t st) {
  const size_t smem = fused_smem_bytes(S, (V & 8) != 0, (V & 64) != 0) + 8192;   // + alignment of the block to an absolute 8 KiB boundary
  static PerDeviceOnce once;
  bool& configured = *once.slot();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(blind_rotate_fused_kernel<S, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  blind_rotate_fused_kernel<S, V><<<(count + S - 1) / S, 64 * S, smem, st>>>(fbsk, small, luts, lut_idx, out, out_rows, tabs, count, stagger);
  return cudaGetLastError();
}

// variant: bit 0 I2F digits, bit 1 tensor-memory key (4 PBS per CTA only), bit 2 keep the two redundant barriers (A/B),
// bit 3 a transpose plane per component (4 PBS per CTA only);
// samples: PBS per CTA, 4 or 6 (6: transpose planes inside the accumulator copies, 12 warps of 168 registers)
cudaError_t launch_blind_rotate_fused(const c2* fbsk, const c2* fbsk_lm, const uint64_t* small, const uint64_t* luts, const uint32_t* lut_idx,
                                      uint64_t* out, const int32_t* out_rows, const c2* tabs, int count, int variant, int stagger, int samples,
                                      cudaStream_t st) {
  if (count <= 0) return cudaSuccess;
  if (samples == 6) {
    switch (variant & 5) {
      case 0: return launch_fused_s<6, 0>(fbsk, small, luts, lut_idx, out, out_rows, tabs, count, stagger, st);
      case 1: return launch_fused_s<6, 1>(fbsk, small, luts, lut_idx, out, out_rows, tabs, count, stagger, st);
      case 4: return launch_fused_s<6, 4>(fbsk, small, luts, lut_idx, out, out_rows, tabs, count, stagger, st);
      default: return launch_fused_s<6, 5>(fbsk, small, luts, lut_idx, out, out_rows, tabs, count, stagger, st);
    }
  }
  if ((variant & 10) == 10) {   // plane per component + tensor-memory key
    if (variant & 1) return launch_fused_s<4, 11>(fbsk_lm, small, luts, lut_idx, out, out_rows, tabs, count, stagger, st);
    return launch_fused_s<4, 10>(fbsk_lm, small, luts, lut_idx, out, out_rows, tabs, count, stagger, st);
  }
  if (variant & 8) {
    if (variant & 1) return launch_fused_s<4, 9>(fbsk, small, luts, lut_idx, out, out_rows, tabs, count, stagger, st);
    return launch_fused_s<4, 8>(fbsk, small, luts, lut_idx, out, out_rows, tabs, count, stagger, st);
  }
  if (variant & 64) {
    if (variant & 1) return launch_fused_s<4, 65>(fbsk, small, luts, lut_idx, out, out_rows, tabs, count, stagger, st);
    return launch_fused_s<4, 64>(fbsk, small, luts, lut_idx, out, out_rows, tabs, count, stagger, st);
  }
  switch (variant & 7) {
    case 0: return launch_fused_s<4, 0>(fbsk, small, luts, lut_idx, out, out_rows, tabs, count, stagger, st);
    case 1: return launch_fused_s<4, 1>(fbsk, small, luts, lut_idx, out, out_rows, tabs, count, stagger, st);
    case 2: return launch_fused_s<4, 2>(fbsk_lm, small, luts, lut_idx, out, out_rows, tabs, count, stagger, st);
    case 3: return launch_fused_s<4, 3>(fbsk_lm, small, luts, lut_idx, out, out_rows, tabs, count, stagger, st);
    case 4: return launch_fused_s<4, 4>(fbsk, small, luts, lut_idx, out, out_rows, tabs, count, stagger, st);
    case 5: return launch_fused_s<4, 5>(fbsk, small, luts, lut_idx, out, out_rows, tabs, count, stagger, st);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_fbsk_lane_major(const c2* fbsk, c2* fbsk_lm, cudaStream_t st) {
  const size_t n = (size_t)kLweN * 4096;
  fbsk_lm_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(fbsk, fbsk_lm);
  return cudaGetLastError();
}

}  // namespace fb
