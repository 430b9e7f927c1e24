// role_sizes.cu -- code size of the four pipeline roles of a warp-specialised blind rotation, built from the existing
// per-thread blocks (br_core.cuh): planning aid for round 2, never launched.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../fhe_regex_b200/csrc -c role_sizes.cu && cuobjdump -sass role_sizes.o | grep -c "^ *\/\*[0-9a-f]\{4\}\*\/"
#include <cuda_runtime.h>
#include "br_core.cuh"
using namespace fb;
// role 1: decomposition + forward pass 1 (+ twist) + column stores of both planes (twiddle moved behind the transpose)
__global__ void role1(const uint32_t* acc, uint32_t a, double* plane_re, double* plane_im) {
  double xr[32], xi[32];
  const int lane = threadIdx.x & 31;
  phaseA_load32(xr, xi, acc, a, lane);
  fft32_fwd_twist(xr, xi);
  col_store_brev(xr, plane_re, lane);
  col_store_brev(xi, plane_im, lane);
}
// role 2: row loads + inter-pass twiddle + forward pass 2 + Fourier MAC + hand-off stores
__global__ void role2(const double* plane_re, const double* plane_im, const c2* tab_f, const c2* b_own, const c2* b_in, double* hand_re, double* hand_im) {
  double xr[32], xi[32];
  const int lane = threadIdx.x & 31;
  row_load(xr, plane_re, lane);
  row_load(xi, plane_im, lane);
  fwd_twiddle_inplace(xr, xi, tab_f, lane);   // (stand-in: same cost as the post-transpose form)
  fft32_fwd(xr, xi);
#pragma unroll
  for (int q = 0; q < 32; q++) {
    const int k2 = brev5(q);
    const double pr = __shfl_xor_sync(0xffffffffu, xr[q], 16);
    const double pi = __shfl_xor_sync(0xffffffffu, xi[q], 16);
    mac_point2(xr[q], xi[q], pr, pi, b_own[32 * k2], b_in[32 * k2]);
  }
  row_store(xr, hand_re, lane);
  row_store(xi, hand_im, lane);
}
// role 3: hand-off loads + inverse pass 1 + twiddle + row stores
__global__ void role3(const double* hand_re, const double* hand_im, const c2* tab_i, double* plane_re, double* plane_im) {
  double xr[32], xi[32];
  const int lane = threadIdx.x & 31;
  row_load(xr, hand_re, lane);
  row_load(xi, hand_im, lane);
  fft32_inv(xr, xi);
  inv_twiddle_inplace(xr, xi, tab_i, lane);
  row_store(xr, plane_re, lane);
  row_store(xi, plane_im, lane);
}
// role 4: column loads + inverse pass 2 + untwist/round + accumulate (shared copy)
__global__ void role4(const double* plane_re, const double* plane_im, uint32_t* acc) {
  double xr[32], xi[32];
  const int lane = threadIdx.x & 31;
  col_load_brev(xr, plane_re, lane);
  col_load_brev(xi, plane_im, lane);
  fft32_inv(xr, xi);
#pragma unroll
  for (int r = 0; r < 32; r++) {
    uint32_t i0, i1;
    phaseC_increments32(xr, xi, r, i0, i1);
    acc[32 * r + lane] += i0;
    acc[32 * r + lane + 1024] += i1;
  }
}
