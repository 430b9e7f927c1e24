// kernels.h -- launchers of the sm_100a kernels in kernels.cu (internal; the public boundary is include/fhe_b200.h)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "br_core.cuh"

namespace fb {
// cudaFuncSetAttribute is per device: one flag per (kernel, device), so a process driving several GPUs (one context
// each) configures every kernel on each of them
struct PerDeviceOnce {
  bool done[64] = {};
  bool* slot() {
    int d = 0;
    cudaGetDevice(&d);
    return &done[d & 63];
  }
};
size_t br_smem_bytes(int S);
cudaError_t launch_bsk_convert(const uint64_t* bsk_std, c2* fbsk, const c2* tabs, cudaStream_t st);
cudaError_t launch_blind_rotate(const c2* fbsk, const uint64_t* small, const uint64_t* luts, const uint32_t* lut_idx,
                                uint64_t* out, const int32_t* out_rows, const c2* tabs, int count, cudaStream_t st);
// throughput variant with the fused CMUX body, 4 PBS per CTA (br_fused.cu)
// variant: bit 0 digits through I2F, bit 1 Fourier key of the step staged in tensor memory (reads fbsk_lm, the key re-ordered by
// launch_fbsk_lane_major)
cudaError_t launch_blind_rotate_fused(const c2* fbsk, const c2* fbsk_lm, const uint64_t* small, const uint64_t* luts, const uint32_t* lut_idx,
                                      uint64_t* out, const int32_t* out_rows, const c2* tabs, int count, int variant, int stagger, int samples, cudaStream_t st);
cudaError_t launch_fbsk_lane_major(const c2* fbsk, c2* fbsk_lm, cudaStream_t st);
// latency variant: one PBS per CTA (br_wide.cu)
size_t br_wide_table_bytes();
void br_wide_make_table(c2* host_tab);
cudaError_t launch_blind_rotate_wide(const c2* fbsk, const uint64_t* small, const uint64_t* luts, const uint32_t* lut_idx,
                                     uint64_t* out, const int32_t* out_rows, const c2* wtab, int count, int skew, int prefetch, cudaStream_t st);
// latency variant for levels between one and two waves of SMs: two PBS per CTA of 512 threads, twiddles in tensor memory (br_wide2.cu)
cudaError_t launch_blind_rotate_wide2(const c2* fbsk, const uint64_t* small, const uint64_t* luts, const uint32_t* lut_idx,
                                      uint64_t* out, const int32_t* out_rows, const c2* wtab, int count, int skew, int prefetch, int sample_offset, cudaStream_t st);
// cluster variant: one PBS per pair of CTAs (br_duo.cu)
size_t br_duo_table_bytes();
void br_duo_make_table(c2* host_tab);
cudaError_t launch_blind_rotate_duo(const c2* fbsk, const uint64_t* small, const uint64_t* luts, const uint32_t* lut_idx,
                                    uint64_t* out, const int32_t* out_rows, const c2* dtab, int count, cudaStream_t st);
int br_duo_max_clusters();
cudaError_t launch_lincomb(uint64_t* arena, const int32_t* out_rows, const int32_t* term_off, const int32_t* term_rows,
                           const int64_t* term_coef, const uint64_t* body_const, int n_out, cudaStream_t st);
// keyswitch as an int8 tensor-core contraction (ks_kernels.cu)
size_t ks_key_bytes();
size_t ks_digit_bytes(int count);
cudaError_t launch_ksk_bytes(const uint64_t* ksk, uint8_t* kb, cudaStream_t st);
// variant 0: mma.sync GEMM (ks_kernels.cu), 1: tcgen05.mma kind::i8 GEMM with TMA-staged operands and TMEM accumulators (ks_umma.cu)
cudaError_t launch_keyswitch_mma(const uint8_t* kb, int8_t* dig, const uint64_t* in, const int32_t* in_rows, uint64_t* out, int count,
                                 int variant, int sms, cudaStream_t st);
cudaError_t launch_fp64_peak(double* sink, int ctas, cudaStream_t st);
double fp64_peak_flops_per_launch(int ctas);
int br_samples_per_cta();
}  // namespace fb
