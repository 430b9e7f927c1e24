// role_probe.cu -- can the four warp schedulers of a B200 SM each run their OWN loop out of their own instruction
// cache?  (Planning probe for a warp-specialised blind rotation: role = warp & 3, one pipeline stage per scheduler.)
// One CTA of 8 warps per SM; warp w runs role w & 3 (warps w and w + 4 share a scheduler).
//   mode 0: every warp runs the same body (one instruction stream per SM, lock-step)                      -- baseline
//   mode 1: each scheduler runs a different body of the same size (four streams per SM), its two warps in lock-step
//   mode 2: like 1, and the second warp of every scheduler starts half a body later (eight streams per SM)
// Bodies: N instructions, either all DFMA (FP64-pipe bound: ideal 4 cycles per instruction per warp with two warps
// per scheduler) or DFMA / integer alternating (ideal 3: 2 + 1).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o role_probe role_probe.cu && ./role_probe
#include <cstdio>
#include <cuda_runtime.h>

// register-only operands: no per-instruction constants (a first version drew a distinct double per DFMA from the constant
// bank and measured the constant cache instead); one integer add with a ROLE immediate per 64 instructions keeps the
// four instantiations from being merged
template <int N, int ROLE, bool MIXED>
__device__ __forceinline__ void body(double (&a)[8], unsigned (&u)[8], const double m, const double c) {
#pragma unroll
  for (int i = 0; i < N; i++) {
    if ((i & 63) == 63) u[0] += (unsigned)(ROLE * 1000003 + 17);
    else if (MIXED && (i & 1)) u[i & 7] = u[i & 7] * 3u + 7u;
    else a[i & 7] = __fma_rn(a[i & 7], m, c);
  }
}

template <int N, bool MIXED>
__global__ void __launch_bounds__(256, 1) probe(double* sink, int iters, int mode, long long* cycles) {
  double a[8];
  unsigned u[8];
  for (int k = 0; k < 8; k++) { a[k] = threadIdx.x + k; u[k] = threadIdx.x * 8 + k; }
  const int warp = threadIdx.x >> 5;
  const int role = mode == 0 ? 0 : (warp & 3);
  const bool late = mode == 2 && warp >= 4;
  __syncthreads();
  const long long t0 = clock64();
  switch (role) {
    case 0:
      if (late) body<N / 2, 0, MIXED>(a, u, 1.0000001, 1e-9);
      for (int it = 0; it < iters; it++) body<N, 0, MIXED>(a, u, 1.0000001, 1e-9);
      break;
    case 1:
      if (late) body<N / 2, 1, MIXED>(a, u, 1.0000001, 1e-9);
      for (int it = 0; it < iters; it++) body<N, 1, MIXED>(a, u, 1.0000001, 1e-9);
      break;
    case 2:
      if (late) body<N / 2, 2, MIXED>(a, u, 1.0000001, 1e-9);
      for (int it = 0; it < iters; it++) body<N, 2, MIXED>(a, u, 1.0000001, 1e-9);
      break;
    default:
      if (late) body<N / 2, 3, MIXED>(a, u, 1.0000001, 1e-9);
      for (int it = 0; it < iters; it++) body<N, 3, MIXED>(a, u, 1.0000001, 1e-9);
      break;
  }
  const long long t1 = clock64();
  double s = 0;
  for (int k = 0; k < 8; k++) s += a[k] + (double)u[k];
  if (s == 1234.5) sink[0] = s;
  if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) atomicMax((unsigned long long*)cycles, (unsigned long long)(t1 - t0));
}

template <int N, bool MIXED>
void run(double* sink, long long* d_cyc) {
  const int iters = 200;
  for (int mode = 0; mode < 3; mode++) {
    probe<N, MIXED><<<148, 256>>>(sink, iters, mode, d_cyc);
    cudaMemset(d_cyc, 0, 8);
    probe<N, MIXED><<<148, 256>>>(sink, iters, mode, d_cyc);
    cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d_cyc, sizeof c, cudaMemcpyDeviceToHost);
    printf("%s body %4d instr (~%2d KB)  mode %d  cycles/instr %.3f (ideal %.1f)\n", MIXED ? "mixed" : "dfma ", N, N * 16 / 1024, mode,
           (double)c / (iters + (mode == 2 ? 0.5 : 0.0)) / N, MIXED ? 3.0 : 4.0);
  }
}

int main() {
  double* sink; long long* d_cyc;
  cudaMalloc(&sink, 8); cudaMalloc(&d_cyc, 8);
  run<512, false>(sink, d_cyc); run<768, false>(sink, d_cyc); run<1024, false>(sink, d_cyc); run<1280, false>(sink, d_cyc);
  run<1536, false>(sink, d_cyc); run<2048, false>(sink, d_cyc);
  run<512, true>(sink, d_cyc); run<768, true>(sink, d_cyc); run<1024, true>(sink, d_cyc); run<1280, true>(sink, d_cyc);
  run<1536, true>(sink, d_cyc); run<2048, true>(sink, d_cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
