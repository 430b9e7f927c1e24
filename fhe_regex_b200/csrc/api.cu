// api.cu -- context management and the batched keyswitch / bootstrap entry points of the C ABI
// (include/fhe_b200.h).  No CPU fallback: every compute entry point needs an sm_100 device.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>
#include "context.h"

using fb::c2;

int fb_fail(fb_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}
int fb_cuda_fail(fb_ctx* ctx, cudaError_t e, const char* what) {
  return fb_fail(ctx, FB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
int fb_reserve(fb_ctx* ctx, fb_devbuf& b, size_t bytes) {
  if (bytes <= b.cap) return FB_OK;
  if (b.p) {
    FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    FB_CUDA(ctx, cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
  }
  size_t cap = bytes + bytes / 4;
  FB_CUDA(ctx, cudaMalloc(&b.p, cap));
  b.cap = cap;
  return FB_OK;
}

static thread_local std::string g_create_err;

extern "C" int fb_ctx_create(fb_ctx** out, int device) {
  if (!out) return FB_ERR_ARG;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0 || device < 0 || device >= n) {
    g_create_err = "no CUDA device available (libfhe_b200 has no CPU fallback)";
    return FB_ERR_NO_DEVICE;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
    g_create_err = "device is not compute capability 10.x (kernels are built for sm_100a only)";
    return FB_ERR_NO_DEVICE;
  }
  fb_ctx* ctx = new (std::nothrow) fb_ctx();
  if (!ctx) return FB_ERR_ARG;
  ctx->device = device;
  ctx->quantum = prop.multiProcessorCount * fb::br_samples_per_cta();
  ctx->sms = prop.multiProcessorCount;
  ctx->dist_shard_min = ctx->sms;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    g_create_err = "cudaSetDevice / cudaStreamCreate failed";
    return FB_ERR_CUDA;
  }
  // twiddle tables
  // (the table of the latency kernel rides behind the two 12x32 tables of the throughput kernel)
  const size_t n_tabs = 2 * fb::kTabEntries * 32;
  const size_t n_wide = fb::br_wide_table_bytes() / sizeof(c2);
  std::vector<c2> tabs(n_tabs + n_wide + fb::br_duo_table_bytes() / sizeof(c2));
  fb::make_twiddle_tables(tabs.data(), tabs.data() + fb::kTabEntries * 32);
  fb::br_wide_make_table(tabs.data() + n_tabs);
  fb::br_duo_make_table(tabs.data() + n_tabs + n_wide);
  // measured on B200: a PBS on a pair of SMs takes 2.39 ms, on one SM 2.37 ms (the step is a chain of dependent
  // stages, not bandwidth) -- the cluster kernel is kept as an option (fb_set_cluster_threshold), off by default
  ctx->duo_pairs = fb::br_duo_max_clusters();
  if (cudaMalloc(&ctx->d_tabs, tabs.size() * sizeof(c2)) != cudaSuccess ||
      cudaMemcpy(ctx->d_tabs, tabs.data(), tabs.size() * sizeof(c2), cudaMemcpyHostToDevice) != cudaSuccess) {
    delete ctx;
    g_create_err = "twiddle table upload failed";
    return FB_ERR_CUDA;
  }
  ctx->d_wtab = ctx->d_tabs + n_tabs;
  ctx->d_dtab = ctx->d_wtab + n_wide;
  *out = ctx;
  return FB_OK;
}

extern "C" void fb_ctx_destroy(fb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  fb_comm_destroy(ctx);
  for (auto& p : ctx->pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
  for (auto& p : ctx->pool) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
  for (auto& e : ctx->pipe_events) cudaEventDestroy(e);
  if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
  if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
  cudaFree(ctx->d_kb);
  cudaFree(ctx->d_fbsk);
  cudaFree(ctx->d_fbsk_lm);
  cudaFree(ctx->d_tabs);
  for (fb_devbuf* b : {&ctx->in, &ctx->small, &ctx->out, &ctx->luts, &ctx->lut_idx, &ctx->digits, &ctx->arena, &ctx->plan_i32,
                       &ctx->plan_i64, &ctx->plan_u64, &ctx->plan_u32, &ctx->regex_luts, &ctx->op_i32, &ctx->op_i64, &ctx->op_u64,
                       &ctx->op_u32, &ctx->op_rows}) cudaFree(b->p);
  for (auto& a : ctx->arenas) cudaFree(a.p);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char* fb_last_error(const fb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }
extern "C" void* fb_ctx_stream(fb_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" int fb_sync(fb_ctx* ctx) {
  if (!ctx) return FB_ERR_ARG;
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FB_OK;
}

// keyswitch key: upload the u64 container, keep only its byte planes resident
static int upload_ksk(fb_ctx* ctx, const uint64_t* h_ksk) {
  if (!ctx->d_fbsk) FB_CUDA(ctx, cudaMalloc(&ctx->d_fbsk, (size_t)fb::kLweN * 4 * fb::kHalfN * sizeof(c2)));
  if (!ctx->d_fbsk_lm) FB_CUDA(ctx, cudaMalloc(&ctx->d_fbsk_lm, (size_t)fb::kLweN * 4 * fb::kHalfN * sizeof(c2)));
  if (!ctx->d_kb) FB_CUDA(ctx, cudaMalloc(&ctx->d_kb, fb::ks_key_bytes()));
  uint64_t* d_ksk = nullptr;  // staging copy of the u64 key; only its byte planes stay resident
  FB_CUDA(ctx, cudaMalloc(&d_ksk, FB_KSK_WORDS * sizeof(uint64_t)));
  cudaError_t e = cudaMemcpyAsync(d_ksk, h_ksk, FB_KSK_WORDS * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = fb::launch_ksk_bytes(d_ksk, ctx->d_kb, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_ksk);
  if (e != cudaSuccess) return fb_cuda_fail(ctx, e, "keyswitch key upload / byte-plane split");
  return FB_OK;
}

// The bootstrapping key as a tfhe-rs 0.2.0 ServerKey holds it: Fourier domain only.  h_fbsk is the key in the serialized
// (plan-independent, natural frequency) order -- see wire.cpp -- which is this library's resident layout: a copy.
extern "C" int fb_load_server_key_fourier(fb_ctx* ctx, const uint64_t* h_ksk, const double* h_fbsk) {
  if (!ctx || !h_ksk || !h_fbsk) return fb_fail(ctx, FB_ERR_ARG, "null argument");
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = upload_ksk(ctx, h_ksk);
  if (rc != FB_OK) return rc;
  FB_CUDA(ctx, cudaMemcpyAsync(ctx->d_fbsk, h_fbsk, (size_t)fb::kLweN * 4 * fb::kHalfN * sizeof(c2), cudaMemcpyHostToDevice, ctx->stream));
  FB_CUDA(ctx, fb::launch_fbsk_lane_major(ctx->d_fbsk, ctx->d_fbsk_lm, ctx->stream));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->have_key = true;
  return FB_OK;
}

// bincode of tfhe::integer::ServerKey (wire.cpp): what the reference's keygen produces (engine.rs:252, ciphertext.rs:44)
extern "C" int fb_load_server_key_bincode(fb_ctx* ctx, const uint8_t* buf, size_t len) {
  if (!ctx || !buf) return fb_fail(ctx, FB_ERR_ARG, "null argument");
  std::vector<uint64_t> ksk(FB_KSK_WORDS);
  std::vector<double> fbsk((size_t)fb::kLweN * 4 * fb::kHalfN * 2);
  int rc = fb_server_key_from_bincode(buf, len, ksk.data(), fbsk.data());
  if (rc != FB_OK) return fb_fail(ctx, rc, "not a bincode tfhe::integer::ServerKey of PARAM_MESSAGE_2_CARRY_2 (layout: wire.cpp)");
  return fb_load_server_key_fourier(ctx, ksk.data(), fbsk.data());
}

// both keys already in device memory (standard domain): byte planes, Fourier transform, lane-major copy
int fb_install_server_key_device(fb_ctx* ctx, uint64_t* d_ksk, uint64_t* d_bsk_std) {
  if (!ctx->d_fbsk) FB_CUDA(ctx, cudaMalloc(&ctx->d_fbsk, (size_t)fb::kLweN * 4 * fb::kHalfN * sizeof(c2)));
  if (!ctx->d_fbsk_lm) FB_CUDA(ctx, cudaMalloc(&ctx->d_fbsk_lm, (size_t)fb::kLweN * 4 * fb::kHalfN * sizeof(c2)));
  if (!ctx->d_kb) FB_CUDA(ctx, cudaMalloc(&ctx->d_kb, fb::ks_key_bytes()));
  cudaError_t e = fb::launch_ksk_bytes(d_ksk, ctx->d_kb, ctx->stream);
  if (e == cudaSuccess) e = fb::launch_bsk_convert(d_bsk_std, ctx->d_fbsk, ctx->d_tabs, ctx->stream);
  if (e == cudaSuccess) e = fb::launch_fbsk_lane_major(ctx->d_fbsk, ctx->d_fbsk_lm, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return fb_cuda_fail(ctx, e, "key conversion (byte planes / Fourier transform)");
  ctx->have_key = true;
  return FB_OK;
}

extern "C" int fb_load_server_key_raw(fb_ctx* ctx, const uint64_t* h_ksk, const uint64_t* h_bsk_std) {
  if (!ctx || !h_ksk || !h_bsk_std) return fb_fail(ctx, FB_ERR_ARG, "null argument");
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  uint64_t *d_ksk = nullptr, *d_std = nullptr;   // staging copies of the u64 keys; only their converted forms stay resident
  FB_CUDA(ctx, cudaMalloc(&d_ksk, FB_KSK_WORDS * sizeof(uint64_t)));
  if (cudaMalloc(&d_std, FB_BSK_WORDS * sizeof(uint64_t)) != cudaSuccess) {
    cudaFree(d_ksk);
    return fb_fail(ctx, FB_ERR_CUDA, "cudaMalloc of the key staging buffer failed");
  }
  cudaError_t e = cudaMemcpyAsync(d_ksk, h_ksk, FB_KSK_WORDS * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_std, h_bsk_std, FB_BSK_WORDS * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream);
  int rc = e == cudaSuccess ? fb_install_server_key_device(ctx, d_ksk, d_std) : fb_cuda_fail(ctx, e, "key upload");
  cudaStreamSynchronize(ctx->stream);
  cudaFree(d_ksk);
  cudaFree(d_std);
  return rc;
}

extern "C" int fb_get_fourier_bsk(fb_ctx* ctx, double* h_out) {
  if (!ctx || !h_out) return fb_fail(ctx, FB_ERR_ARG, "null argument");
  if (!ctx->have_key) return fb_fail(ctx, FB_ERR_NO_KEY, "server key not loaded");
  FB_CUDA(ctx, cudaMemcpyAsync(h_out, ctx->d_fbsk, (size_t)fb::kLweN * 4 * fb::kHalfN * sizeof(c2), cudaMemcpyDeviceToHost, ctx->stream));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FB_OK;
}

// FP64 FMA pipe throughput of this device (TFLOP/s), best of `reps` launches timed with CUDA events
extern "C" int fb_measure_fp64_peak(fb_ctx* ctx, int reps, double* tflops) {
  if (!ctx || !tflops) return fb_fail(ctx, FB_ERR_ARG, "null argument");
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaDeviceProp prop;
  FB_CUDA(ctx, cudaGetDeviceProperties(&prop, ctx->device));
  const int ctas = prop.multiProcessorCount * 8;
  double* sink = nullptr;
  FB_CUDA(ctx, cudaMalloc(&sink, 8));
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  double best = 0;
  cudaError_t e = cudaSuccess;
  for (int r = 0; r < reps + 2 && e == cudaSuccess; r++) {
    cudaEventRecord(a, ctx->stream);
    e = fb::launch_fp64_peak(sink, ctas, ctx->stream);
    cudaEventRecord(b, ctx->stream);
    if (e == cudaSuccess) e = cudaEventSynchronize(b);
    float ms = 0;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, a, b);
    if (e == cudaSuccess && r >= 2 && ms > 0) best = std::max(best, fb::fp64_peak_flops_per_launch(ctas) / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(sink);
  if (e != cudaSuccess) return fb_cuda_fail(ctx, e, "fp64 peak probe");
  *tflops = best;
  return FB_OK;
}

// batch sizes that fill the GPU evenly are multiples of this (SM count x samples per CTA)
extern "C" int fb_pbs_batch_quantum(fb_ctx* ctx) {
  if (!ctx) return FB_ERR_ARG;
  return ctx->quantum;
}

// ---- timed launches ---------------------------------------------------------------------------
static bool timing_begin(fb_ctx* ctx, int kind, fb_event_pair& ev) {
  if (!ctx->timing) return false;
  if (!ctx->pool.empty()) {
    ev = ctx->pool.back();
    ctx->pool.pop_back();
  } else if (cudaEventCreate(&ev.a) != cudaSuccess || cudaEventCreate(&ev.b) != cudaSuccess) {
    return false;
  }
  ev.kind = kind;
  cudaEventRecord(ev.a, ctx->stream);
  return true;
}
static void timing_end(fb_ctx* ctx, bool on, fb_event_pair& ev) {
  if (!on) return;
  cudaEventRecord(ev.b, ctx->stream);
  ctx->pending.push_back(ev);
}

int fb_run_keyswitch(fb_ctx* ctx, const uint64_t* d_in, const int32_t* d_in_rows, uint64_t* d_small, int count) {
  int rc = fb_reserve(ctx, ctx->digits, fb::ks_digit_bytes(count));
  if (rc) return rc;
  fb_event_pair ev;
  bool t = timing_begin(ctx, 0, ev);
  cudaError_t e = fb::launch_keyswitch_mma(ctx->d_kb, (int8_t*)ctx->digits.p, d_in, d_in_rows, d_small, count, ctx->ks_variant, ctx->sms, ctx->stream);
  timing_end(ctx, t, ev);
  if (e != cudaSuccess) return fb_cuda_fail(ctx, e, "keyswitch_kernel launch");
  ctx->ks.ks_launches++;
  ctx->ks.ks_samples += count;
  return FB_OK;
}
int fb_run_blind_rotate(fb_ctx* ctx, const uint64_t* d_small, const uint64_t* d_luts, const uint32_t* d_lut_idx,
                        uint64_t* d_out, const int32_t* d_out_rows, int count) {
  fb_event_pair ev;
  bool t = timing_begin(ctx, 1, ev);
  // a level narrower than two waves of SMs is latency: one PBS per CTA (br_wide.cu); otherwise throughput
  // (kernels.cu, up to 4 PBS per CTA).  A short tail behind full throughput waves (e.g. 620 = 592 + 28) would
  // cost a whole extra throughput wave: it goes to the latency kernel instead.
  // Narrower still -- at most as many PBS as the device runs CTA pairs -- a PBS gets two SMs (br_duo.cu).
  // Between one and two waves of SMs (149 .. 296 PBS on a B200) two PBS share a CTA (br_wide2.cu): both waves at once.
  auto narrow = [&](const uint64_t* sm, const uint32_t* li, uint64_t* o, const int32_t* orows, int n) {
    if (n <= ctx->duo_max) return fb::launch_blind_rotate_duo(ctx->d_fbsk, sm, d_luts, li, o, orows, ctx->d_dtab, n, ctx->stream);
    if (ctx->wide_pair == 2 || (ctx->wide_pair == 1 && n > ctx->sms && n <= 2 * ctx->sms))
      return fb::launch_blind_rotate_wide2(ctx->d_fbsk, sm, d_luts, li, o, orows, ctx->d_wtab, n, ctx->wide_skew, ctx->wide_pair_prefetch, ctx->wide_pair_offset, ctx->stream);
    return fb::launch_blind_rotate_wide(ctx->d_fbsk, sm, d_luts, li, o, orows, ctx->d_wtab, n, ctx->wide_skew, ctx->wide_prefetch, ctx->stream);
  };
  const int narrow_max = ctx->wide_max > ctx->duo_max ? ctx->wide_max : ctx->duo_max;
  cudaError_t e;
  if (count <= narrow_max) {
    e = narrow(d_small, d_lut_idx, d_out, d_out_rows, count);
  } else {
    // PBS per CTA of the throughput launch (fused body only, option "br_samples"): 4, 6, or 0 = whichever the wave arithmetic
    // favours.  Measured wave times on a B200 (ms): 4 per SM 6.24, 6 per SM 9.47, latency kernel 2.30 (up to one PBS per SM),
    // pair kernel 4.30 (up to two per SM); per PBS 6 per SM is 1 % slower, so it only wins through the quantisation, e.g.
    // 1 020 = 888 + 132: 11.8 ms against two waves of 592 = 12.5.
    auto est = [&](int Sx) {
      const int qx = ctx->sms * Sx;
      const double tw = Sx == 6 ? 9.47 : 6.24;
      const int full = count / qx, rem = count % qx;
      if (rem == 0) return full * tw;
      if (full > 0 && rem <= narrow_max) return full * tw + (rem <= ctx->sms ? 2.30 : 4.30);
      return (full + 1) * tw;
    };
    int S = fb::br_samples_per_cta();
    if (ctx->br_variant >= 1 && count > 4 * ctx->sms) {
      if (ctx->br_samples == 6) S = 6;
      else if (ctx->br_samples == 0 && est(6) < 0.97 * est(4)) S = 6;
    }
    const int q = ctx->sms * S;
    const int tail = (q > 0) ? count % q : 0;
    const int head = (tail > 0 && tail <= narrow_max) ? count - tail : count;
    // the fused body is built for full SMs: batches that do not fill the GPU at 4 per SM keep the phase-by-phase body
    const bool fused = ctx->br_variant >= 1 && head > 3 * ctx->sms;
    const int fv = (ctx->br_variant - 1) | (ctx->br_barriers ? 4 : 0) | ((ctx->br_planes == 2 && S == 4) ? 8 : 0) | ((ctx->br_planes == 3 && S == 4 && ctx->br_variant <= 2) ? 64 : 0);
    e = fused ? fb::launch_blind_rotate_fused(ctx->d_fbsk, ctx->d_fbsk_lm, d_small, d_luts, d_lut_idx, d_out, d_out_rows, ctx->d_tabs, head, fv, ctx->br_stagger | (ctx->br_stagger_groups << 24) | (ctx->br_sync << 25) | (ctx->br_resync << 26), S, ctx->stream)
              : fb::launch_blind_rotate(ctx->d_fbsk, d_small, d_luts, d_lut_idx, d_out, d_out_rows, ctx->d_tabs, head, ctx->stream);
    if (e == cudaSuccess && head < count)
      e = narrow(d_small + (size_t)head * FB_LWE_SMALL_WORDS, d_lut_idx + head, d_out_rows ? d_out : d_out + (size_t)head * FB_LWE_BIG_WORDS,
                 d_out_rows ? d_out_rows + head : nullptr, tail);
  }
  timing_end(ctx, t, ev);
  if (e != cudaSuccess) return fb_cuda_fail(ctx, e, "blind_rotate_kernel launch");
  ctx->ks.br_launches++;
  ctx->ks.br_samples += count;
  return FB_OK;
}
int fb_run_lincomb(fb_ctx* ctx, uint64_t* d_arena, const int32_t* out_rows, const int32_t* term_off,
                   const int32_t* term_rows, const int64_t* term_coef, const uint64_t* body_const, int n_out) {
  fb_event_pair ev;
  bool t = timing_begin(ctx, 2, ev);
  cudaError_t e = fb::launch_lincomb(d_arena, out_rows, term_off, term_rows, term_coef, body_const, n_out, ctx->stream);
  timing_end(ctx, t, ev);
  if (e != cudaSuccess) return fb_cuda_fail(ctx, e, "lincomb_kernel launch");
  ctx->ks.lin_launches++;
  return FB_OK;
}

extern "C" int fb_host_alloc(size_t bytes, void** out) {
  if (!out) return FB_ERR_ARG;
  *out = nullptr;
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();   // not sticky: the caller falls back to pageable memory
    return FB_ERR_CUDA;
  }
  *out = p;
  return FB_OK;
}
extern "C" void fb_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

extern "C" int fb_kernel_timing_enable(fb_ctx* ctx, int on) {
  if (!ctx) return FB_ERR_ARG;
  ctx->timing = on != 0;
  return FB_OK;
}
static int resolve_pending(fb_ctx* ctx) {
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (auto& p : ctx->pending) {
    float ms = 0.f;
    FB_CUDA(ctx, cudaEventElapsedTime(&ms, p.a, p.b));
    if (p.kind == 0) ctx->ks.ks_ms += ms;
    else if (p.kind == 1) ctx->ks.br_ms += ms;
    else ctx->ks.lin_ms += ms;
    ctx->pool.push_back(p);
  }
  ctx->pending.clear();
  return FB_OK;
}
extern "C" int fb_set_latency_threshold(fb_ctx* ctx, int max_count) {
  if (!ctx || max_count < 0) return FB_ERR_ARG;
  const int prev = ctx->wide_max;
  ctx->wide_max = max_count;
  return prev;
}
extern "C" int fb_set_cluster_threshold(fb_ctx* ctx, int max_count) {
  if (!ctx || max_count < 0) return FB_ERR_ARG;
  const int prev = ctx->duo_max;
  ctx->duo_max = max_count;
  return prev;
}
// ---- options: every knob of the library is per context and set explicitly (nothing is read from the environment) ----
namespace {
struct OptionDesc { const char* name; int64_t lo, hi; };
const OptionDesc kOptions[] = {
    {"latency_threshold", 0, 1 << 30},      // = fb_set_latency_threshold
    {"cluster_threshold", 0, 1 << 30},      // = fb_set_cluster_threshold
    {"br_variant", 0, 4},                   // throughput blind rotation: 0 phase-by-phase body, 1 fused, 2 fused + I2F digits, 3 fused + TMEM key, 4 both
    {"wide_skew", 0, 100000},               // latency kernel: hold-back of one CTA half after the MAC, cycles
    {"wide_prefetch", 0, 4},                // latency kernel: GGSW groups fetched before the pre-MAC barrier
    {"plan_reference_shaped", 0, 1},        // 1: evaluate every variant the reference enumerates (no absorption)
    {"plan_timing", 0, 1},                  // 1: planner phase times on stderr
    {"br_stagger", 0, 100000},              // fused throughput kernel: start skew between the samples of a CTA, cycles per sample index
    {"ks_variant", 0, 1},                   // keyswitch GEMM: 0 mma.sync, 1 tcgen05.mma kind::i8 (TMA operands, TMEM accumulators)
    {"wide_pair", 0, 2},                    // latency kernel with two PBS per CTA: 0 never, 1 for batches between one and two waves of SMs, 2 for every narrow batch
    {"wide_pair_prefetch", 0, 2},           // pair kernel: GGSW groups fetched before the pre-MAC barrier
    {"wide_pair_offset", 0, 100000},        // pair kernel: cycles the second sample of a CTA starts late
    {"br_samples", 0, 6},                   // fused throughput kernel: PBS per CTA, 4 or 6 (6: transpose planes inside the accumulator copies), 0 = by wave arithmetic
    {"br_stagger_groups", 0, 1},            // 1: "br_stagger" delays the odd samples of a CTA only (two scheduler groups, one instruction stream per scheduler)
    {"br_planes", 1, 3},                    // fused throughput kernel at 4 PBS per CTA: 2 = a transpose plane per component (one barrier per transpose)
    {"pbs_chunks", 3, 5},                   // fb_pbs_batch: 3 = chunks of 4, rest, 4 waves (default), 5 = 1, 6, rest, 6, 1 for batches of at least 24 waves
    {"br_sync", 0, 1},                      // fused throughput kernel: 1 = the samples of a CTA start their rotation together (default)
    {"br_resync", 0, 63},                   // fused throughput kernel: the samples of a CTA meet at a barrier every this many CMUX steps (0 = never)
    {"br_barriers", 0, 1},                  // fused throughput kernel: 1 keeps the two per-step barriers that are not needed (A/B measurements)
    {"dist_shard_min", 0, 1 << 30},         // fb_has_match_dist: levels of at most this many PBS are computed by every rank instead of being sharded (default: the SM count)
};
int64_t* option_slot(fb_ctx* ctx, const char* name, int64_t& shadow, int& which) {
  for (int i = 0; i < (int)(sizeof(kOptions) / sizeof(kOptions[0])); i++)
    if (std::strcmp(name, kOptions[i].name) == 0) {
      which = i;
      switch (i) {
        case 0: shadow = ctx->wide_max; break;
        case 1: shadow = ctx->duo_max; break;
        case 2: shadow = ctx->br_variant; break;
        case 3: shadow = ctx->wide_skew; break;
        case 4: shadow = ctx->wide_prefetch; break;
        case 5: shadow = ctx->plan_absorb ? 0 : 1; break;
        case 6: shadow = ctx->plan_timing ? 1 : 0; break;
        case 7: shadow = ctx->br_stagger; break;
        case 8: shadow = ctx->ks_variant; break;
        case 9: shadow = ctx->wide_pair; break;
        case 10: shadow = ctx->wide_pair_prefetch; break;
        case 11: shadow = ctx->wide_pair_offset; break;
        case 12: shadow = ctx->br_samples; break;
        case 13: shadow = ctx->br_stagger_groups; break;
        case 14: shadow = ctx->br_planes; break;
        case 15: shadow = ctx->pbs_chunks; break;
        case 16: shadow = ctx->br_sync; break;
        case 17: shadow = ctx->br_resync; break;
        case 18: shadow = ctx->br_barriers; break;
        case 19: shadow = ctx->dist_shard_min; break;
      }
      return &shadow;
    }
  return nullptr;
}
}  // namespace

extern "C" int fb_get_option(fb_ctx* ctx, const char* name, int64_t* value) {
  if (!ctx || !name || !value) return fb_fail(ctx, FB_ERR_ARG, "null argument");
  int64_t cur = 0;
  int which = -1;
  if (!option_slot(ctx, name, cur, which)) return fb_fail(ctx, FB_ERR_ARG, std::string("unknown option: ") + name);
  *value = cur;
  return FB_OK;
}

extern "C" int fb_set_option(fb_ctx* ctx, const char* name, int64_t value) {
  if (!ctx || !name) return fb_fail(ctx, FB_ERR_ARG, "null argument");
  int64_t cur = 0;
  int which = -1;
  if (!option_slot(ctx, name, cur, which)) return fb_fail(ctx, FB_ERR_ARG, std::string("unknown option: ") + name);
  if (value < kOptions[which].lo || value > kOptions[which].hi) return fb_fail(ctx, FB_ERR_ARG, std::string("option out of range: ") + name);
  switch (which) {
    case 0: ctx->wide_max = (int)value; break;
    case 1: ctx->duo_max = (int)value; break;
    case 2: ctx->br_variant = (int)value; break;
    case 3: ctx->wide_skew = (int)value; break;
    case 4: ctx->wide_prefetch = (int)value; break;
    case 5: ctx->plan_absorb = value == 0; break;
    case 6: ctx->plan_timing = value != 0; break;
    case 7: ctx->br_stagger = (int)value; break;
    case 8: ctx->ks_variant = (int)value; break;
    case 9: ctx->wide_pair = (int)value; break;
    case 10: ctx->wide_pair_prefetch = (int)value; break;
    case 11: ctx->wide_pair_offset = (int)value; break;
    case 12:
      if (value != 0 && value != 4 && value != 6) return fb_fail(ctx, FB_ERR_ARG, "br_samples is 0 (automatic), 4 or 6");
      ctx->br_samples = (int)value;
      ctx->quantum = ctx->sms * (value == 6 ? 6 : 4);
      break;
    case 13: ctx->br_stagger_groups = (int)value; break;
    case 14: ctx->br_planes = (int)value; break;
    case 15: ctx->pbs_chunks = (int)value; break;
    case 16: ctx->br_sync = (int)value; break;
    case 17: ctx->br_resync = (int)value; break;
    case 18: ctx->br_barriers = (int)value; break;
    case 19: ctx->dist_shard_min = (int)value; break;
  }
  return FB_OK;
}

extern "C" int fb_kernel_stats_reset(fb_ctx* ctx) {
  if (!ctx) return FB_ERR_ARG;
  int rc = resolve_pending(ctx);
  ctx->ks = fb_kernel_stats{};
  return rc;
}
extern "C" int fb_kernel_stats_get(fb_ctx* ctx, fb_kernel_stats* out) {
  if (!ctx || !out) return FB_ERR_ARG;
  int rc = resolve_pending(ctx);
  *out = ctx->ks;
  return rc;
}

// ---- batch entry points ---------------------------------------------------------------------------
extern "C" int fb_keyswitch_batch(fb_ctx* ctx, const uint64_t* h_in, size_t count, uint64_t* h_out) {
  if (!ctx || !h_in || !h_out) return fb_fail(ctx, FB_ERR_ARG, "null argument");
  if (!ctx->have_key) return fb_fail(ctx, FB_ERR_NO_KEY, "server key not loaded");
  if (count == 0) return FB_OK;
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc;
  if ((rc = fb_reserve(ctx, ctx->in, count * FB_LWE_BIG_WORDS * 8))) return rc;
  if ((rc = fb_reserve(ctx, ctx->small, count * FB_LWE_SMALL_WORDS * 8))) return rc;
  FB_CUDA(ctx, cudaMemcpyAsync(ctx->in.p, h_in, count * FB_LWE_BIG_WORDS * 8, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = fb_run_keyswitch(ctx, (const uint64_t*)ctx->in.p, nullptr, (uint64_t*)ctx->small.p, (int)count))) return rc;
  FB_CUDA(ctx, cudaMemcpyAsync(h_out, ctx->small.p, count * FB_LWE_SMALL_WORDS * 8, cudaMemcpyDeviceToHost, ctx->stream));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FB_OK;
}

extern "C" int fb_pbs_batch_dev(fb_ctx* ctx, const uint64_t* d_in, const uint64_t* d_luts, const uint32_t* d_lut_idx,
                                size_t count, uint64_t* d_out) {
  if (!ctx || !d_in || !d_luts || !d_lut_idx || !d_out) return fb_fail(ctx, FB_ERR_ARG, "null argument");
  if (!ctx->have_key) return fb_fail(ctx, FB_ERR_NO_KEY, "server key not loaded");
  if (count == 0) return FB_OK;
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc;
  if ((rc = fb_reserve(ctx, ctx->small, count * FB_LWE_SMALL_WORDS * 8))) return rc;
  if ((rc = fb_run_keyswitch(ctx, d_in, nullptr, (uint64_t*)ctx->small.p, (int)count))) return rc;
  return fb_run_blind_rotate(ctx, (const uint64_t*)ctx->small.p, d_luts, d_lut_idx, d_out, nullptr, (int)count);
}

static int upload_luts(fb_ctx* ctx, const uint64_t* h_luts, size_t n_luts, const uint32_t* h_lut_idx, size_t count) {
  for (size_t b = 0; b < count; b++)
    if (h_lut_idx[b] >= n_luts) return fb_fail(ctx, FB_ERR_ARG, "lut index out of range");
  int rc;
  if ((rc = fb_reserve(ctx, ctx->luts, n_luts * FB_POLY_SIZE * 8))) return rc;
  if ((rc = fb_reserve(ctx, ctx->lut_idx, count * 4))) return rc;
  FB_CUDA(ctx, cudaMemcpyAsync(ctx->luts.p, h_luts, n_luts * FB_POLY_SIZE * 8, cudaMemcpyHostToDevice, ctx->stream));
  FB_CUDA(ctx, cudaMemcpyAsync(ctx->lut_idx.p, h_lut_idx, count * 4, cudaMemcpyHostToDevice, ctx->stream));
  return FB_OK;
}

extern "C" int fb_pbs_batch(fb_ctx* ctx, const uint64_t* h_in, const uint64_t* h_luts, size_t n_luts,
                            const uint32_t* h_lut_idx, size_t count, uint64_t* h_out) {
  if (!ctx || !h_in || !h_luts || !h_lut_idx || !h_out || n_luts == 0) return fb_fail(ctx, FB_ERR_ARG, "null argument");
  if (!ctx->have_key) return fb_fail(ctx, FB_ERR_NO_KEY, "server key not loaded");
  if (count == 0) return FB_OK;
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc;
  if ((rc = upload_luts(ctx, h_luts, n_luts, h_lut_idx, count))) return rc;
  if ((rc = fb_reserve(ctx, ctx->in, count * FB_LWE_BIG_WORDS * 8))) return rc;
  if ((rc = fb_reserve(ctx, ctx->out, count * FB_LWE_BIG_WORDS * 8))) return rc;
  uint64_t* d_in = (uint64_t*)ctx->in.p;
  uint64_t* d_out = (uint64_t*)ctx->out.p;
  const uint32_t* d_idx = (const uint32_t*)ctx->lut_idx.p;
  // Large batches are pipelined in chunks of whole throughput waves: the upload of chunk k+1 and the download of chunk k-1
  // run on their own streams (two DMA engines) under the bootstraps of chunk k, so only the first upload and the last
  // download are exposed -- hence a short first and a short last chunk (4 waves each) around one long one.  Few chunks: every
  // extra launch ends in a tail where SMs wait for the slowest CTA (measured: 12 chunks of 4 waves cost more than the copies
  // they hide; option "pbs_chunks" 5 = chunks of 1, 6, rest, 6, 1 waves: 0.6 ms exposed instead of 1.6 ms, but +2.5 ms of
  // launch tails at 28 416 PBS -- tools/e2e_gap_probe.py).
  const size_t q = (size_t)ctx->quantum;
  size_t bounds[6] = {0, count, count, count, count, count};
  size_t n_chunks = 1;
  if (count >= 24 * q && ctx->pbs_chunks >= 5) {
    const size_t mid = (count - 14 * q) / q * q;
    bounds[1] = q;
    bounds[2] = 7 * q;
    bounds[3] = 7 * q + mid;
    bounds[4] = count - q;
    n_chunks = 5;
  } else if (count >= 16 * q) {
    const size_t mid = (count - 8 * q) / q * q;
    bounds[1] = 4 * q;
    bounds[2] = 4 * q + mid;
    n_chunks = 3;
  } else if (count > 8 * q) {
    bounds[1] = (count / 2 + q - 1) / q * q;
    n_chunks = 2;
  }
  if (n_chunks < 2) {
    FB_CUDA(ctx, cudaMemcpyAsync(d_in, h_in, count * FB_LWE_BIG_WORDS * 8, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = fb_pbs_batch_dev(ctx, d_in, (const uint64_t*)ctx->luts.p, d_idx, count, d_out))) return rc;
    FB_CUDA(ctx, cudaMemcpyAsync(h_out, d_out, count * FB_LWE_BIG_WORDS * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FB_OK;
  }
  if (!ctx->h2d_stream) FB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
  if (!ctx->d2h_stream) FB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
  while (ctx->pipe_events.size() < 3 * n_chunks + 1) {   // per chunk: uploaded, bootstrapped, "next upload may go"; + start
    cudaEvent_t e;
    FB_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->pipe_events.push_back(e);
  }
  // the LUT upload (and whatever the caller queued before) precedes the first bootstrap; the copy streams must not
  // overwrite buffers an earlier call on the context stream may still use
  cudaEvent_t start = ctx->pipe_events[2 * n_chunks];
  FB_CUDA(ctx, cudaEventRecord(start, ctx->stream));
  FB_CUDA(ctx, cudaStreamWaitEvent(ctx->h2d_stream, start, 0));
  FB_CUDA(ctx, cudaStreamWaitEvent(ctx->d2h_stream, start, 0));
  // option "plan_timing": the timeline of the pipeline on stderr (when each upload, each chunk's bootstraps and each download
  // ended, in ms from the start of the call on the device)
  const bool trace = ctx->plan_timing;
  std::vector<cudaEvent_t> tev;
  auto mark = [&](cudaStream_t st) {
    if (!trace) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) == cudaSuccess) { cudaEventRecord(e, st); tev.push_back(e); }
  };
  mark(ctx->stream);
  // The upload of chunk k+1 is released by an event on the compute stream placed between the keyswitch and the blind rotation
  // of chunk k, not queued up front: with all uploads in flight from the start of the call the first kernels did not begin
  // before the LAST upload had ended (measured with the timeline below: the first chunk finished 8.5 ms late, exactly the
  // 466 MB / 55 GB/s of the whole input), which was the whole gap between this call and the device-resident one.
  auto upload = [&](size_t k) -> cudaError_t {
    const size_t off = bounds[k], n = bounds[k + 1] - bounds[k];
    cudaError_t e = cudaMemcpyAsync(d_in + off * FB_LWE_BIG_WORDS, h_in + off * FB_LWE_BIG_WORDS, n * FB_LWE_BIG_WORDS * 8, cudaMemcpyHostToDevice,
                                    ctx->h2d_stream);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->pipe_events[2 * k], ctx->h2d_stream);
    mark(ctx->h2d_stream);
    return e;
  };
  FB_CUDA(ctx, upload(0));
  for (size_t k = 0; k < n_chunks; k++) {
    const size_t off = bounds[k], n = bounds[k + 1] - bounds[k];
    FB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->pipe_events[2 * k], 0));
    rc = fb_reserve(ctx, ctx->small, n * FB_LWE_SMALL_WORDS * 8);
    if (!rc) rc = fb_run_keyswitch(ctx, d_in + off * FB_LWE_BIG_WORDS, nullptr, (uint64_t*)ctx->small.p, (int)n);
    if (!rc && k + 1 < n_chunks) {
      cudaEvent_t go = ctx->pipe_events[2 * n_chunks + 1 + k];
      cudaError_t e = cudaEventRecord(go, ctx->stream);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->h2d_stream, go, 0);
      if (e == cudaSuccess) e = upload(k + 1);
      if (e != cudaSuccess) rc = fb_cuda_fail(ctx, e, "fb_pbs_batch upload");
    }
    if (!rc)
      rc = fb_run_blind_rotate(ctx, (const uint64_t*)ctx->small.p, (const uint64_t*)ctx->luts.p, d_idx + off, d_out + off * FB_LWE_BIG_WORDS, nullptr, (int)n);
    if (rc) {
      cudaStreamSynchronize(ctx->h2d_stream);   // nothing of this call may still touch the caller's buffers
      cudaStreamSynchronize(ctx->d2h_stream);
      cudaStreamSynchronize(ctx->stream);
      for (cudaEvent_t e : tev) cudaEventDestroy(e);
      return rc;
    }
    FB_CUDA(ctx, cudaEventRecord(ctx->pipe_events[2 * k + 1], ctx->stream));
    mark(ctx->stream);
    FB_CUDA(ctx, cudaStreamWaitEvent(ctx->d2h_stream, ctx->pipe_events[2 * k + 1], 0));
    FB_CUDA(ctx, cudaMemcpyAsync(h_out + off * FB_LWE_BIG_WORDS, d_out + off * FB_LWE_BIG_WORDS, n * FB_LWE_BIG_WORDS * 8,
                                 cudaMemcpyDeviceToHost, ctx->d2h_stream));
    mark(ctx->d2h_stream);
  }
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->d2h_stream));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (trace && tev.size() == 1 + 3 * n_chunks) {
    // marks in queue order: start, upload 0, then per chunk k: [upload k+1,] bootstraps k, download k
    auto at = [&](size_t i) { float ms = 0.f; cudaEventElapsedTime(&ms, tev[0], tev[i]); return ms; };
    std::fprintf(stderr, "fb_pbs_batch %zu PBS in %zu chunks:", count, n_chunks);
    for (size_t k = 0; k < n_chunks; k++) {
      const size_t up = k == 0 ? 1 : 3 * k - 1, br = k + 1 < n_chunks ? 3 * k + 3 : 3 * k + 2;
      std::fprintf(stderr, "  [%zu: %zu PBS, upload done %.2f, bootstraps done %.2f, download done %.2f]", k, bounds[k + 1] - bounds[k], at(up), at(br),
                   at(br + 1));
    }
    std::fprintf(stderr, " ms\n");
  }
  for (cudaEvent_t e : tev) cudaEventDestroy(e);
  return FB_OK;
}

extern "C" int fb_bootstrap_small_batch(fb_ctx* ctx, const uint64_t* h_small, const uint64_t* h_luts, size_t n_luts,
                                        const uint32_t* h_lut_idx, size_t count, uint64_t* h_out) {
  if (!ctx || !h_small || !h_luts || !h_lut_idx || !h_out || n_luts == 0) return fb_fail(ctx, FB_ERR_ARG, "null argument");
  if (!ctx->have_key) return fb_fail(ctx, FB_ERR_NO_KEY, "server key not loaded");
  if (count == 0) return FB_OK;
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc;
  if ((rc = upload_luts(ctx, h_luts, n_luts, h_lut_idx, count))) return rc;
  if ((rc = fb_reserve(ctx, ctx->small, count * FB_LWE_SMALL_WORDS * 8))) return rc;
  if ((rc = fb_reserve(ctx, ctx->out, count * FB_LWE_BIG_WORDS * 8))) return rc;
  FB_CUDA(ctx, cudaMemcpyAsync(ctx->small.p, h_small, count * FB_LWE_SMALL_WORDS * 8, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = fb_run_blind_rotate(ctx, (const uint64_t*)ctx->small.p, (const uint64_t*)ctx->luts.p, (const uint32_t*)ctx->lut_idx.p,
                                (uint64_t*)ctx->out.p, nullptr, (int)count)))
    return rc;
  FB_CUDA(ctx, cudaMemcpyAsync(h_out, ctx->out.p, count * FB_LWE_BIG_WORDS * 8, cudaMemcpyDeviceToHost, ctx->stream));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FB_OK;
}
