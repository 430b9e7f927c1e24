// wide_lab.cu -- timing sandbox for the latency blind rotation (br_wide.cu): the same stages on synthetic data (random Fourier
// key, random accumulator), with clock stamps at every barrier and switches that remove one cost at a time.  Not a
// correctness test (tests/ cover the product kernel); it answers "where do the ~870 cycles of a stage go".
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../fhe_regex_b200/csrc -o wide_lab wide_lab.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>
#include "br_wide.cuh"
#include "ptx_sync.cuh"

using namespace fb;

constexpr int kStageBytes = 4 * kHalfN * (int)sizeof(c2);
constexpr int kBufBytes = 2 * kHalfN * (int)sizeof(c2);
constexpr size_t kOffBufA = 2 * (size_t)kStageBytes;
constexpr size_t kOffBufB = kOffBufA + kBufBytes;
constexpr size_t kOffAcc = kOffBufB + kBufBytes;
constexpr size_t kOffAt = kOffAcc + 2 * kN * sizeof(uint32_t);
constexpr size_t kOffBars = kOffAt + 768 * sizeof(uint16_t);
constexpr size_t kLabSmem = kOffBars + 2 * sizeof(uint64_t) + 16;

__device__ __forceinline__ void half_sync(int P) { asm volatile("bar.sync %0, 128;" ::"r"(1 + P) : "memory"); }

// MODE bit 0: half 1 idles through the per-half stages (what does a lone half cost?)
//      bit 1: no stores in the transform stages (stores replaced by a cheap dependency sink)
//      bit 2: no FP64 butterflies (loads and stores only)
//      bit 3: the GGSW values of the MAC come from registers (no shared-memory key reads: upper bound of what serving the key
//             from tensor memory could gain)
template <int MODE>
__global__ void __launch_bounds__(256, 1)
lab_kernel(const c2* __restrict__ fbsk, const uint16_t* __restrict__ at_g, const c2* __restrict__ wtab, uint32_t* __restrict__ out,
           long long* __restrict__ stamps, int n_steps, int skew_cycles) {
  extern __shared__ __align__(128) unsigned char smem[];
  c2* bufA = reinterpret_cast<c2*>(smem + kOffBufA);
  c2* bufB = reinterpret_cast<c2*>(smem + kOffBufB);
  uint32_t* acc = reinterpret_cast<uint32_t*>(smem + kOffAcc);
  uint16_t* at = reinterpret_cast<uint16_t*>(smem + kOffAt);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kOffBars);
  const int tid = threadIdx.x, P = tid >> 7, t = tid & 127;
  wide::Tw tw;
  wide::load_tw(tw, wtab, t);
  for (int i = tid; i < 768; i += 256) at[i] = at_g[blockIdx.x * 768 + i];
  uint32_t own[16];
#pragma unroll
  for (int m = 0; m < 8; m++) {
    const uint32_t j = (uint32_t)t + 128u * m;
    own[2 * m] = (j * 2654435761u) ^ (blockIdx.x * 97u);
    own[2 * m + 1] = (j * 40503u) + P;
    acc[P * kN + j] = own[2 * m];
    acc[P * kN + j + 1024u] = own[2 * m + 1];
  }
  auto issue_ggsw = [&](int i, int b) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive_expect_tx(full_bar + b, (uint32_t)kStageBytes);
    const unsigned char* src = reinterpret_cast<const unsigned char*>(fbsk + (size_t)i * 4 * kHalfN);
    unsigned char* dst = smem + (size_t)b * kStageBytes;
#pragma unroll
    for (int c = 0; c < 4; c++) bulk_g2s(dst + c * (kStageBytes / 4), src + c * (kStageBytes / 4), kStageBytes / 4, full_bar + b);
  };
  if (tid == 0) {
    mbar_init(full_bar, 1);
    mbar_init(full_bar + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    issue_ggsw(0, 0);
  }
  __syncthreads();
  uint32_t* accp = acc + P * kN;
  c2* bufA_p = bufA + P * kHalfN;
  c2* bufB_p = bufB + P * kHalfN;
  const bool idle = (MODE & 1) && P == 1;
  const bool stamp = stamps != nullptr && blockIdx.x == 0 && (tid & 31) == 0;
  long long* my = stamps + (tid >> 5) * 4 * 10;
#define STAMP(k) if (stamp && n >= 300 && n < 304) my[(n - 300) * 10 + (k)] = clock64();
#pragma unroll 1
  for (int n = 0; n < n_steps; n++) {
    const int i = n % kLweN;
    const uint32_t a = (uint32_t)at[i] & 4095u;
    if (tid == 0 && n + 1 < n_steps) issue_ggsw((n + 1) % kLweN, (n + 1) & 1);
    STAMP(0)
    if (!idle) wide::fwd_stage1(accp, own, a, t, tw, bufA_p);
    half_sync(P);
    STAMP(1)
    if (!idle) wide::fwd_stage2(bufA_p, bufB_p, t, tw);
    half_sync(P);
    STAMP(2)
    if (!idle) wide::fwd_stage3(bufB_p, bufA_p, t);
    mbar_wait(full_bar + (n & 1), (uint32_t)(n >> 1) & 1u);
    const c2* ggsw = reinterpret_cast<const c2*>(smem + (size_t)(n & 1) * kStageBytes);
    c2 gpre[16];
    if (MODE & 8) {
#pragma unroll
      for (int g = 0; g < 16; g++) { gpre[g].x = __hiloint2double(0x3fe00000 + g, (int)(a + n)); gpre[g].y = __hiloint2double(0x3fd00000 + g, (int)a); }
    } else {
      c2 g3[12];
      wide::mac_prefetch<3>(ggsw, P, t, g3);
#pragma unroll
      for (int g = 0; g < 12; g++) gpre[g] = g3[g];
    }
    __syncthreads();
    STAMP(3)
    if (MODE & 8) wide::mac_inv_stage1<4>(bufA, bufA + kHalfN, ggsw, gpre, P, t, tw, bufB_p);
    else {
      c2 g3[12];
#pragma unroll
      for (int g = 0; g < 12; g++) g3[g] = gpre[g];
      wide::mac_inv_stage1<3>(bufA, bufA + kHalfN, ggsw, g3, P, t, tw, bufB_p);
    }
    __syncthreads();
    STAMP(4)
    if (P == 1 && skew_cycles > 0) {
      const long long t0 = clock64();
      while (clock64() - t0 < (long long)skew_cycles) {}
    }
    if (!idle) wide::inv_stage2(bufB_p, bufA_p, t, tw);
    half_sync(P);
    STAMP(5)
    if (!idle) wide::inv_stage3(bufA_p, bufB_p, t);
    half_sync(P);
    STAMP(6)
    if (!idle) wide::phaseC_accumulate(bufB_p, t, own, accp);
    half_sync(P);
    STAMP(7)
  }
  __syncthreads();
  uint32_t x = 0;
#pragma unroll
  for (int m = 0; m < 16; m++) x ^= own[m];
  out[blockIdx.x * 256 + tid] = x ^ acc[tid];
}

template <int MODE>
static void run(const char* name, const c2* fbsk, const uint16_t* at, const c2* wtab, uint32_t* out, long long* stamps, int skew, bool print_stamps) {
  cudaFuncSetAttribute(lab_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLabSmem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int n_steps = 742;
  lab_kernel<MODE><<<148, 256, kLabSmem>>>(fbsk, at, wtab, out, nullptr, n_steps, skew);
  cudaEventRecord(e0);
  lab_kernel<MODE><<<148, 256, kLabSmem>>>(fbsk, at, wtab, out, stamps, n_steps, skew);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("%-44s skew %4d  %.3f ms  %.0f cycles/step (at 1965 MHz)  %s\n", name, skew, ms, ms * 1.965e6 / n_steps, err == cudaSuccess ? "" : cudaGetErrorString(err));
  if (print_stamps) {
    long long h[8 * 4 * 10];
    cudaMemcpy(h, stamps, sizeof h, cudaMemcpyDeviceToHost);
    for (int w = 0; w < 8; w += 4) {
      for (int s = 1; s < 3; s++) {
        printf("   warp %d step %d: stage durations", w, 300 + s);
        for (int k = 0; k < 7; k++) printf(" %5lld", h[w * 40 + s * 10 + k + 1] - h[w * 40 + s * 10 + k]);
        printf("  | step %lld\n", h[w * 40 + (s + 1) * 10] - h[w * 40 + s * 10]);
      }
    }
  }
}

int main() {
  std::vector<c2> h_key((size_t)kLweN * 4 * kHalfN);
  uint64_t s = 88172645463325252ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
  for (auto& v : h_key) { v.x = (double)(int64_t)rnd() * 0x1p-64; v.y = (double)(int64_t)rnd() * 0x1p-64; }
  std::vector<uint16_t> h_at(148 * 768);
  for (auto& v : h_at) v = (uint16_t)((rnd() & 4095u) | 0x8000u);
  std::vector<c2> h_tab(wide::kTabC2);
  wide::make_wide_table(h_tab.data());
  c2 *fbsk, *wtab; uint16_t* at; uint32_t* out; long long* stamps;
  cudaMalloc(&fbsk, h_key.size() * sizeof(c2));
  cudaMalloc(&wtab, h_tab.size() * sizeof(c2));
  cudaMalloc(&at, h_at.size() * 2);
  cudaMalloc(&out, 148 * 256 * 4);
  cudaMalloc(&stamps, 8 * 4 * 10 * 8);
  cudaMemcpy(fbsk, h_key.data(), h_key.size() * sizeof(c2), cudaMemcpyHostToDevice);
  cudaMemcpy(wtab, h_tab.data(), h_tab.size() * sizeof(c2), cudaMemcpyHostToDevice);
  cudaMemcpy(at, h_at.data(), h_at.size() * 2, cudaMemcpyHostToDevice);
  run<0>("product stages", fbsk, at, wtab, out, stamps, 200, true);
  run<0>("product stages", fbsk, at, wtab, out, stamps, 0, true);
  run<0>("product stages", fbsk, at, wtab, out, stamps, 450, false);
  run<0>("product stages", fbsk, at, wtab, out, stamps, 700, false);
  run<1>("half 1 idles through the per-half stages", fbsk, at, wtab, out, stamps, 0, true);
  run<8>("GGSW values from registers (no key reads)", fbsk, at, wtab, out, stamps, 200, true);
  run<8>("GGSW values from registers (no key reads)", fbsk, at, wtab, out, stamps, 0, false);
  return 0;
}
