// coissue_probe.cu -- which instruction classes issue in the shadow of FP64 instructions?
// A DFMA/DADD occupies its scheduler's FP64 pipe for 2 cycles per warp instruction.  Per iteration and warp: 16 DFMA (constant-bank
// multiplier, 16 independent chains) interleaved with 32 instructions of one other class (8 independent chains, register operands).
// 148 CTAs x 256 threads = 2 warps per scheduler.  FP64 alone: 64 cycles per iteration; 32 single-issue instructions alone: 64.
//   both together: 64 = they co-issue, 128 = they serialise, 96 = bound by one issue slot per cycle.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o coissue_probe coissue_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__constant__ double kc[8] = {1.0000001, 0.9999999, 1.0000002, 0.9999998, 1.0000003, 0.9999997, 1.0000004, 0.9999996};
enum { NONE, IMAD, LOP3, IADD, SHF, FFMA, MOV64, PRMT, ISETP_SEL };
template <int MODE, int ND, int DCLASS>
__global__ void __launch_bounds__(256, 1) probe(double* sink, long long* cyc, int iters, double seed, uint32_t m0, uint32_t m1) {
  double a[16];
#pragma unroll
  for (int k = 0; k < 16; k++) a[k] = seed + threadIdx.x + k;
  uint32_t u[8], v[8];
  float f[8];
#pragma unroll
  for (int k = 0; k < 8; k++) { u[k] = threadIdx.x * 2654435761u + k; v[k] = m0 + k * m1; f[k] = 1.0f + k; }
  const float fm = 1.0f + (float)m0 * 1e-9f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < 16; k++) {
      if (k < ND) {
        if (DCLASS == 0) a[k] = __fma_rn(a[k], kc[k & 7], a[(k + 1) & 15]);
        else a[k] = __dadd_rn(a[k], a[(k + 1) & 15]);
      }
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const int c = (2 * k + j) & 7;
        if (MODE == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[c]) : "r"(v[c]), "r"(v[(c + 1) & 7]));
        if (MODE == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[c]) : "r"(v[c]), "r"(v[(c + 1) & 7]));
        if (MODE == IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[c]) : "r"(v[c]));
        if (MODE == SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(u[c]) : "r"(v[c]), "r"(v[(c + 1) & 7]));
        if (MODE == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(fm), "f"(f[(c + 1) & 7]));
        if (MODE == PRMT) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(u[c]) : "r"(v[c]), "r"(v[(c + 1) & 7]));
        if (MODE == ISETP_SEL) { if (j == 0) { asm volatile("{ .reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %2, %0, p; }" : "+r"(u[c]) : "r"(v[c]), "r"(v[(c + 1) & 7])); } }
      }
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < 16; k++) s += a[k];
#pragma unroll
  for (int k = 0; k < 8; k++) s += u[k] + f[k];
  if (s == 12345.678) sink[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE, int ND, int DCLASS>
void run(const char* name) {
  double* sink; long long* cyc;
  cudaMalloc(&sink, 8); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4000;
  probe<MODE, ND, DCLASS><<<148, 256>>>(sink, cyc, iters, 0.5, 12345u, 77u);
  probe<MODE, ND, DCLASS><<<148, 256>>>(sink, cyc, iters, 0.5, 12345u, 77u);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  printf("%-36s %2d x %s  %.1f cycles per iteration\n", name, ND, DCLASS ? "DADD" : "DFMA", (double)h[0] / iters);
  cudaFree(sink); cudaFree(cyc);
}
#define BOTH(M, NAME) run<M, 0, 0>(NAME " alone"); run<M, 16, 0>(NAME " + FP64"); run<M, 16, 1>(NAME " + FP64")
int main() {
  run<NONE, 16, 0>("FP64 alone");
  run<NONE, 16, 1>("FP64 alone");
  BOTH(IMAD, "32 IMAD (mad.lo.u32)");
  BOTH(LOP3, "32 LOP3");
  BOTH(IADD, "32 add.u32");
  BOTH(SHF, "32 SHF");
  BOTH(FFMA, "32 FFMA");
  BOTH(PRMT, "32 PRMT");
  BOTH(ISETP_SEL, "16 x (ISETP, SEL)");
  return 0;
}
