// br_duo.cu -- K2-K4, cluster variant: one PBS per PAIR of SMs (thread-block cluster of 2 CTAs x 256 threads), for
// DAG levels of at most half as many PBS as the GPU has SMs.  Same arithmetic as the other two blind rotations
// (kernels.cu, br_wide.cu) and the same Fourier bootstrapping key; the per-thread stages are in br_duo.cuh.
//
// Replaces, like kernels.cu, the blind rotation under /root/reference/src/regex/execution.rs:76,93,110,143,173,190.
// Exists because the last levels of has_match's fold (engine.rs:30-33) hold a handful of PBS whose cost is pure
// latency: CTA r of the pair owns polynomial r (decomposition + forward transform of accumulator polynomial r,
// output polynomial r of the external product), so each SM moves half the shared-memory traffic and issues half
// the FP64 work of a CMUX step; the spectra cross once per step through distributed shared memory.
//
// Shared memory per CTA (139 KiB): two stages of the two GGSW polynomials this CTA multiplies by (2 x 2 x 16 KiB,
// cp.async.bulk, issued a step ahead) + two transform buffers (2 x 16 KiB) + two spectrum buffers the peer reads
// (2 x 16 KiB, alternating: the peer may still read step n's while step n+1's is written) + the accumulator copy
// the rotation reads (8 KiB) + step list.
#include <cuda_runtime.h>
#include <stdint.h>
#include "br_duo.cuh"
#include "kernels.h"
#include "ptx_sync.cuh"

namespace fb {

namespace {
constexpr int kPolyBytes = kHalfN * (int)sizeof(c2);          // 16384
constexpr size_t kOffG = 0;                                   // [2 stages][own, peer][1024] c2
constexpr size_t kOffBufA = kOffG + 4 * (size_t)kPolyBytes;
constexpr size_t kOffBufB = kOffBufA + kPolyBytes;
constexpr size_t kOffSpec = kOffBufB + kPolyBytes;            // [2][1024] c2
constexpr size_t kOffAcc = kOffSpec + 2 * (size_t)kPolyBytes;
constexpr size_t kOffAt = kOffAcc + kN * sizeof(uint32_t);
constexpr size_t kOffSteps = kOffAt + 768 * sizeof(uint16_t);
constexpr size_t kOffBars = kOffSteps + 768 * sizeof(uint16_t);
constexpr size_t kDuoSmem = kOffBars + 2 * sizeof(uint64_t) + 16;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_peer(uint32_t local_smem_addr, uint32_t peer) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(peer));
  return r;
}
__device__ __forceinline__ c2 ld_cluster_c2(uint32_t addr) {
  c2 v;
  asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
  return v;
}
}  // namespace

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(duo::kThreads, 1)
blind_rotate_duo_kernel(const c2* __restrict__ fbsk, const uint64_t* __restrict__ small, const uint64_t* __restrict__ luts,
                        const uint32_t* __restrict__ lut_idx, uint64_t* __restrict__ out, const int32_t* __restrict__ out_rows,
                        const c2* __restrict__ dtab, int count) {
  extern __shared__ __align__(128) unsigned char smem[];
  c2* bufA = reinterpret_cast<c2*>(smem + kOffBufA);
  c2* bufB = reinterpret_cast<c2*>(smem + kOffBufB);
  c2* spec = reinterpret_cast<c2*>(smem + kOffSpec);
  uint32_t* acc = reinterpret_cast<uint32_t*>(smem + kOffAcc);
  uint16_t* at = reinterpret_cast<uint16_t*>(smem + kOffAt);      // mod-switched ciphertext, bit 15: step needed
  uint16_t* steps = reinterpret_cast<uint16_t*>(smem + kOffSteps);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kOffBars);
  int* n_steps_p = reinterpret_cast<int*>(full_bar + 2);

  const int b = threadIdx.x;
  const int me = (int)cluster_ctarank();           // polynomial this CTA owns
  const int sample = blockIdx.x >> 1;              // grid = 2 * count: no CTA leaves early (cluster barriers)

  duo::Tw tw;
  duo::load_tw(tw, dtab, b);
  for (int i = b; i < 768; i += duo::kThreads) {
    uint32_t a = 0;
    if (i < kSmall) {
      const uint64_t x = small[(size_t)sample * kSmall + i];
      a = modswitch(x);
      if (i < kLweN) a = (a & 4095u) | ((x != 0 && (a & 4095u) != 0) ? 0x8000u : 0u);
    }
    at[i] = (uint16_t)a;
  }
  __syncthreads();

  // this CTA multiplies by GGSW[row me][column me] (own spectrum) and GGSW[row 1-me][column me] (peer's spectrum)
  auto issue_ggsw = [&](int i, int st) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive_expect_tx(full_bar + st, 2u * (uint32_t)kPolyBytes);
    const c2* g = fbsk + (size_t)i * 4 * kHalfN;
    unsigned char* dst = smem + kOffG + (size_t)st * 2 * kPolyBytes;
    bulk_g2s(dst, g + (size_t)(me * 2 + me) * kHalfN, kPolyBytes, full_bar + st);
    bulk_g2s(dst + kPolyBytes, g + (size_t)((1 - me) * 2 + me) * kHalfN, kPolyBytes, full_bar + st);
  };

  if (b == 0) {
    int n = 0;
    for (int i = 0; i < kLweN; i++)
      if (at[i] & 0x8000u) steps[n++] = (uint16_t)i;
    *n_steps_p = n;
    mbar_init(full_bar, 1);
    mbar_init(full_bar + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (n > 0) issue_ggsw(steps[0], 0);
  }
  // accumulator init: (0, lut * X^{-b}), top words; thread b owns coefficients b + 256r and b + 256r + 1024
  uint32_t own[8];
  {
    const uint64_t* lut = luts + (size_t)lut_idx[sample] * kN;
    const uint32_t rot = (4096u - (uint32_t)at[kLweN]) & 4095u;
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const uint32_t j = (uint32_t)b + 256u * r;
      own[2 * r] = (me == 0) ? 0u : (uint32_t)(rot_read(lut, j, rot) >> 32);
      own[2 * r + 1] = (me == 0) ? 0u : (uint32_t)(rot_read(lut, j + 1024u, rot) >> 32);
      acc[j] = own[2 * r];
      acc[j + 1024u] = own[2 * r + 1];
    }
  }
  __syncthreads();
  const int n_steps = *n_steps_p;
  const uint32_t peer_spec = map_to_peer(smem_u32(spec), (uint32_t)(1 - me));
  cluster_sync_all();                               // the peer's shared memory exists before anyone reads it

#pragma unroll 1
  for (int n = 0; n < n_steps; n++) {
    const int i = steps[n];
    const uint32_t a = (uint32_t)at[i] & 4095u;
    const int st = n & 1;
    // the other GGSW stage was last read by the MAC of step n-1, several barriers ago
    if (b == 0 && n + 1 < n_steps) issue_ggsw(steps[n + 1], st ^ 1);
    duo::fwd_stage0(acc, own, a, b, tw, bufA);
    __syncthreads();
    duo::mid_stage<4, false>(bufA, bufB, b, tw.f1);
    __syncthreads();
    duo::mid_stage<16, false>(bufB, bufA, b, tw.f2);
    __syncthreads();
    duo::mid_stage<64, false>(bufA, bufB, b, tw.f3);
    __syncthreads();
    c2 X[4], Xp[4];
    duo::fwd_stage4(bufB, b, X);
    c2* myspec = spec + st * kHalfN;
#pragma unroll
    for (int k = 0; k < 4; k++) myspec[b + 256 * k] = X[k];
    cluster_sync_all();                             // both spectra written and visible across the pair
#pragma unroll
    for (int k = 0; k < 4; k++) Xp[k] = ld_cluster_c2(peer_spec + (uint32_t)((st * kHalfN + b + 256 * k) * sizeof(c2)));
    mbar_wait(full_bar + st, (uint32_t)(n >> 1) & 1u);
    const c2* g = reinterpret_cast<const c2*>(smem + kOffG + (size_t)st * 2 * kPolyBytes);
    duo::mac_inv_stage0(X, Xp, g, g + kHalfN, b, tw, bufA);
    __syncthreads();
    duo::mid_stage<4, true>(bufA, bufB, b, tw.i1);
    __syncthreads();
    duo::mid_stage<16, true>(bufB, bufA, b, tw.i2);
    __syncthreads();
    duo::inv_stage3(bufA, bufB, b, tw);
    __syncthreads();
    duo::inv_stage4_accumulate(bufB, b, own);
#pragma unroll
    for (int r = 0; r < 4; r++) {
      acc[b + 256 * r] = own[2 * r];
      acc[b + 256 * r + 1024] = own[2 * r + 1];
    }
    __syncthreads();
  }

  // K4: sample extract of the constant coefficient: mask_0 = a_0, mask_j = -a_{N-j} (polynomial 0); body = b_0 (polynomial 1)
  {
    const size_t row = out_rows ? (size_t)out_rows[sample] : (size_t)sample;
    uint64_t* o = out + row * kBig;
    if (me == 0) {
      for (int j = b; j < kN; j += duo::kThreads) {
        const uint32_t v = (j == 0) ? acc[0] : 0u - acc[kN - j];
        o[j] = (uint64_t)v << 32;
      }
    } else if (b == 0) {
      o[kN] = (uint64_t)acc[0] << 32;
    }
  }
  cluster_sync_all();                               // nobody exits while the peer may still read its spectrum
}

size_t br_duo_table_bytes() { return (size_t)duo::kTabC2 * sizeof(c2); }
void br_duo_make_table(c2* host_tab) { duo::make_duo_table(host_tab); }

// how many pairs the device runs at once (GPC boundaries can cost a pair or two below SMs / 2)
int br_duo_max_clusters() {
  if (cudaFuncSetAttribute(blind_rotate_duo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDuoSmem) != cudaSuccess) return 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * 148);
  cfg.blockDim = dim3(duo::kThreads);
  cfg.dynamicSmemBytes = kDuoSmem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, blind_rotate_duo_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

cudaError_t launch_blind_rotate_duo(const c2* fbsk, const uint64_t* small, const uint64_t* luts, const uint32_t* lut_idx,
                                    uint64_t* out, const int32_t* out_rows, const c2* dtab, int count, cudaStream_t st) {
  if (count <= 0) return cudaSuccess;
  static PerDeviceOnce once;
  bool& configured = *once.slot();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(blind_rotate_duo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDuoSmem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  blind_rotate_duo_kernel<<<2 * count, duo::kThreads, kDuoSmem, st>>>(fbsk, small, luts, lut_idx, out, out_rows, dtab, count);
  return cudaGetLastError();
}

}  // namespace fb
