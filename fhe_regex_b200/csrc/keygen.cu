// keygen.cu -- server key generation on the GPU: `ServerKey::new(&client_key)` / `gen_keys_radix`
// (/root/reference/src/regex/engine.rs:252, src/regex/ciphertext.rs:42-45).
//
//   KSK  [2048][5][743] u64: row (i, l) is an LWE encryption under the small key (n = 742) of
//        big_key[i] * 2^(64 - 3 (l + 1)), noise sigma_lwe                      (tfhe-rs LweKeyswitchKey, levels most significant first)
//   BSK  [742][1][2][2][2048] u64: GGSW(small_key[i]) under the GLWE key (= the big key, k = 1, N = 2048), one level of
//        base 2^23: row 0 encrypts -(s_i 2^41) S(X), row 1 encrypts s_i 2^41, noise sigma_glwe
// Both keys are generated in device memory, optionally copied out, and installed like fb_load_server_key_raw does
// (byte planes for the tensor-core keyswitch, Fourier transform of the bootstrapping key).  Randomness is a counter-based
// generator (SplitMix64 finaliser over (seed, stream, index)) with Box-Muller Gaussians: reproducible per seed, and --
// like the CPU generator of client.cpp -- statistically equivalent to, not bit-identical with, tfhe-rs keys (tfhe-rs
// draws from an AES-CTR CSPRNG).  A deployment would seed from the OS; the secret keys travel to the GPU only because
// the reference generates its server key where it holds the client key.
#include <cuda_runtime.h>
#include <stdint.h>
#include "context.h"

namespace {

constexpr int kLweN = 742, kN = 2048, kSmall = 743, kKsLevels = 5, kKsBaseLog = 3, kPbsBaseLog = 23;
constexpr double kSigmaLwe = 7.069849454709433e-06;    // lwe_modular_std_dev of PARAM_MESSAGE_2_CARRY_2
constexpr double kSigmaGlwe = 2.9403601535432533e-16;  // glwe_modular_std_dev

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// uniform 64 bits for (seed, stream, index)
__device__ __forceinline__ uint64_t rnd(uint64_t seed, uint64_t stream, uint64_t idx) {
  return mix64(mix64(seed * 0x9E3779B97F4A7C15ull + stream) + idx * 0xD1B54A32D192ED03ull + 0x9E3779B97F4A7C15ull);
}
// N(0, sigma^2) on the torus, as a wrapping u64
__device__ __forceinline__ uint64_t torus_gauss(uint64_t seed, uint64_t stream, uint64_t idx, double sigma) {
  const double u1 = ((double)(rnd(seed, stream, 2 * idx) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  const double u2 = ((double)(rnd(seed, stream, 2 * idx + 1) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  const double g = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
  return (uint64_t)(int64_t)llrint(g * sigma * 18446744073709551616.0);
}

// one CTA per KSK row (i, l): mask = fresh uniform words, body = <mask, s> + big_key[i] 2^(64 - 3(l+1)) + e
__global__ void __launch_bounds__(256)
keygen_ksk_kernel(const uint64_t* __restrict__ big_key, const uint64_t* __restrict__ small_key, uint64_t seed, uint64_t* __restrict__ ksk) {
  __shared__ uint64_t part[256];
  const int row = blockIdx.x, i = row / kKsLevels, l = row % kKsLevels;
  uint64_t* out = ksk + (size_t)row * kSmall;
  uint64_t acc = 0;
  for (int c = threadIdx.x; c < kLweN; c += 256) {
    const uint64_t a = rnd(seed, 0x10000000ull + row, c);
    out[c] = a;
    acc += a * small_key[c];
  }
  part[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) part[threadIdx.x] += part[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0)
    out[kLweN] = part[0] + (big_key[i] << (64 - kKsBaseLog * (l + 1))) + torus_gauss(seed, 0x18000000ull, row, kSigmaLwe);
}

// one CTA per GGSW row (i, r): A = fresh uniform polynomial, B = A * S (negacyclic, binary S) + plaintext + e
__global__ void __launch_bounds__(256)
keygen_bsk_kernel(const uint64_t* __restrict__ big_key, const uint64_t* __restrict__ small_key, uint64_t seed, uint64_t* __restrict__ bsk) {
  __shared__ uint64_t A[kN];
  __shared__ uint16_t ones[kN];
  __shared__ int n_ones;
  const int row = blockIdx.x, i = row >> 1, r = row & 1;
  uint64_t* outA = bsk + ((size_t)row * 2 + 0) * kN;
  uint64_t* outB = bsk + ((size_t)row * 2 + 1) * kN;
  if (threadIdx.x == 0) {
    int n = 0;
    for (int t = 0; t < kN; t++)
      if (big_key[t] & 1ull) ones[n++] = (uint16_t)t;
    n_ones = n;
  }
  for (int j = threadIdx.x; j < kN; j += 256) {
    const uint64_t a = rnd(seed, 0x20000000ull + row, j);
    A[j] = a;
    outA[j] = a;
  }
  __syncthreads();
  const uint64_t factor = (small_key[i] & 1ull) << (64 - kPbsBaseLog);
  for (int j = threadIdx.x; j < kN; j += 256) {
    uint64_t b = 0;
    for (int u = 0; u < n_ones; u++) {   // (A * S)[j] = sum_{t in S} +A[j - t] (j >= t) or -A[j - t + N]
      const int t = ones[u];
      const uint64_t a = A[(j - t) & (kN - 1)];
      b += (j >= t) ? a : (uint64_t)0 - a;
    }
    const uint64_t pt = (r == 0) ? (uint64_t)0 - factor * (big_key[j] & 1ull) : (j == 0 ? factor : 0ull);
    outB[j] = b + pt + torus_gauss(seed, 0x28000000ull + row, j, kSigmaGlwe);
  }
}

}  // namespace

extern "C" int fb_keygen_server_gpu(fb_ctx* ctx, const uint64_t* h_big_key, const uint64_t* h_small_key, uint64_t seed, uint64_t* h_ksk,
                                    uint64_t* h_bsk_std) {
  if (!ctx || !h_big_key || !h_small_key) return fb_fail(ctx, FB_ERR_ARG, "null argument");
  for (int t = 0; t < kN; t++)
    if (h_big_key[t] > 1) return fb_fail(ctx, FB_ERR_ARG, "secret keys are binary");
  for (int t = 0; t < kLweN; t++)
    if (h_small_key[t] > 1) return fb_fail(ctx, FB_ERR_ARG, "secret keys are binary");
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  uint64_t *d_big = nullptr, *d_small = nullptr, *d_ksk = nullptr, *d_bsk = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_big);
    cudaFree(d_small);
    cudaFree(d_ksk);
    cudaFree(d_bsk);
  };
#define KG_CUDA(call)                                                         \
  do {                                                                        \
    cudaError_t _e = (call);                                                  \
    if (_e != cudaSuccess) { cleanup(); return fb_cuda_fail(ctx, _e, #call); } \
  } while (0)
  KG_CUDA(cudaMalloc(&d_big, kN * 8));
  KG_CUDA(cudaMalloc(&d_small, kLweN * 8));
  KG_CUDA(cudaMalloc(&d_ksk, FB_KSK_WORDS * 8));
  KG_CUDA(cudaMalloc(&d_bsk, FB_BSK_WORDS * 8));
  KG_CUDA(cudaMemcpyAsync(d_big, h_big_key, kN * 8, cudaMemcpyHostToDevice, ctx->stream));
  KG_CUDA(cudaMemcpyAsync(d_small, h_small_key, kLweN * 8, cudaMemcpyHostToDevice, ctx->stream));
  keygen_ksk_kernel<<<kN * kKsLevels, 256, 0, ctx->stream>>>(d_big, d_small, seed, d_ksk);
  KG_CUDA(cudaGetLastError());
  keygen_bsk_kernel<<<kLweN * 2, 256, 0, ctx->stream>>>(d_big, d_small, seed, d_bsk);
  KG_CUDA(cudaGetLastError());
  if (h_ksk) KG_CUDA(cudaMemcpyAsync(h_ksk, d_ksk, FB_KSK_WORDS * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (h_bsk_std) KG_CUDA(cudaMemcpyAsync(h_bsk_std, d_bsk, FB_BSK_WORDS * 8, cudaMemcpyDeviceToHost, ctx->stream));
  int rc = fb_install_server_key_device(ctx, d_ksk, d_bsk);
  // the secret keys do not stay on the device
  cudaMemsetAsync(d_big, 0, kN * 8, ctx->stream);
  cudaMemsetAsync(d_small, 0, kLweN * 8, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  cleanup();
#undef KG_CUDA
  return rc;
}
