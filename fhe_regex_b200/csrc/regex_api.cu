// regex_api.cu -- has_match() behind the C ABI: build the level-synchronous PBS plan on the host
// (regex_host.cpp) and run it on the GPU, every level as lincomb -> keyswitch -> blind-rotate launches
// on the context stream with no host synchronisation in between.
//
// Replaces has_match (engine.rs:8-42) + Execution (execution.rs:37-223) of the reference.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include "context.h"
#include "regex_host.h"

using namespace fbre;

namespace {

struct DevPlan {
  // all per-level arrays concatenated into one upload each
  std::vector<int32_t> i32;
  std::vector<int64_t> i64;
  std::vector<uint64_t> u64;
  std::vector<uint32_t> u32;
  struct Off { size_t lin_out, lin_off, lin_rows, lin_coef, lin_const, in_rows, out_rows, lut_idx; int n_lin, n_pbs; int32_t out_base; };
  std::vector<Off> levels;
};

// m instances of one plan (m contents against one pattern): instance k lives in arena rows [k * n_rows, (k+1) * n_rows),
// every level's arrays are the instances' arrays back to back, so a level is still one lincomb + one keyswitch +
// one blind-rotation launch, m times as wide
void flatten_plan(const Plan& plan, DevPlan& d, size_t m) {
  const int32_t R = plan.n_rows;
  for (auto& l : plan.levels) {
    DevPlan::Off o;
    const size_t nl = l.lin_out_rows.size(), np = l.in_rows.size(), nt = l.lin_term_rows.size();
    o.n_lin = (int)(nl * m);
    o.n_pbs = (int)(np * m);
    o.out_base = l.out_row_base;
    o.lin_out = d.i32.size();
    for (size_t k = 0; k < m; k++) for (int32_t r : l.lin_out_rows) d.i32.push_back(r + (int32_t)k * R);
    o.lin_off = d.i32.size();
    for (size_t k = 0; k < m; k++) for (size_t j = 0; j < nl; j++) d.i32.push_back(l.lin_term_off[j] + (int32_t)(k * nt));
    d.i32.push_back((int32_t)(m * nt));
    o.lin_rows = d.i32.size();
    for (size_t k = 0; k < m; k++) for (int32_t r : l.lin_term_rows) d.i32.push_back(r + (int32_t)k * R);
    o.in_rows = d.i32.size();
    for (size_t k = 0; k < m; k++) for (int32_t r : l.in_rows) d.i32.push_back(r + (int32_t)k * R);
    o.out_rows = d.i32.size();
    if (m > 1)
      for (size_t k = 0; k < m; k++) for (size_t j = 0; j < np; j++) d.i32.push_back(l.out_row_base + (int32_t)j + (int32_t)k * R);
    o.lin_coef = d.i64.size();
    for (size_t k = 0; k < m; k++) d.i64.insert(d.i64.end(), l.lin_coef.begin(), l.lin_coef.end());
    o.lin_const = d.u64.size();
    for (size_t k = 0; k < m; k++) d.u64.insert(d.u64.end(), l.lin_const.begin(), l.lin_const.end());
    o.lut_idx = d.u32.size();
    for (size_t k = 0; k < m; k++) d.u32.insert(d.u32.end(), l.lut_idx.begin(), l.lut_idx.end());
    d.levels.push_back(o);
  }
}

// fixed accumulator table (shortint generate_accumulator layout), LutId order
void build_lut_table(std::vector<uint64_t>& luts) {
  luts.resize((size_t)LUT_COUNT * FB_POLY_SIZE);
  for (uint32_t id = 0; id < LUT_COUNT; id++) {
    uint64_t f16[16];
    for (uint32_t x = 0; x < 16; x++) f16[x] = lut_value(id, x);
    fb_make_lut(f16, luts.data() + (size_t)id * FB_POLY_SIZE);
  }
}

void write_trivial_radix(uint64_t* h_out, uint64_t bit) {
  std::memset(h_out, 0, sizeof(uint64_t) * 4 * FB_LWE_BIG_WORDS);
  h_out[FB_POLY_SIZE] = bit << 59;
}

// run m instances of a plan; instance k's input (n_in_rows x 2049, at h_in + k * n_in_rows rows) is uploaded to the
// first n_in_rows arena rows of its block, its result radix goes to h_out_radix + k * 4 rows
// dist: the context is a rank of a communicator (comm.cu) and every rank runs this call with the same plan and input:
// rank r uploads and bootstraps only slice r of the input rows and of every PBS level, the slices are exchanged in place
// in the arena over NVLink, the (cheap) linear combinations are evaluated by every rank on its full copy of the arena.
int run_plan(fb_ctx* ctx, const Plan& plan, const uint64_t* h_in, size_t n_in_rows, uint64_t* h_out_radix, double* gpu_ms, size_t m = 1,
             bool dist = false) {
  if (plan.result_kind < 2) {
    for (size_t k = 0; k < m; k++) write_trivial_radix(h_out_radix + k * 4 * FB_LWE_BIG_WORDS, (uint64_t)plan.result_kind);
    if (gpu_ms) *gpu_ms = 0;
    return FB_OK;
  }
  if (!ctx->have_key) return fb_fail(ctx, FB_ERR_NO_KEY, "server key not loaded");
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  DevPlan d;
  flatten_plan(plan, d, m);
  int rc;
  const size_t row_bytes = (size_t)FB_LWE_BIG_WORDS * 8;
  // at least 1 GiB from the first match on (65 536 ciphertext rows; the device has 180 GB): growing the arena means
  // a device-wide cudaFree + cudaMalloc in the middle of a serving loop
  if ((rc = fb_reserve(ctx, ctx->arena, std::max<size_t>(m * (size_t)plan.n_rows * row_bytes, (size_t)1 << 30)))) return rc;
  if ((rc = fb_reserve(ctx, ctx->small, (m * (size_t)plan.stats.max_level_width + 1) * FB_LWE_SMALL_WORDS * 8))) return rc;
  if ((rc = fb_reserve(ctx, ctx->plan_i32, (d.i32.size() + 1) * 4))) return rc;
  if ((rc = fb_reserve(ctx, ctx->plan_i64, (d.i64.size() + 1) * 8))) return rc;
  if ((rc = fb_reserve(ctx, ctx->plan_u64, (d.u64.size() + 1) * 8))) return rc;
  if ((rc = fb_reserve(ctx, ctx->plan_u32, (d.u32.size() + 1) * 4))) return rc;
  if ((rc = fb_reserve(ctx, ctx->regex_luts, (size_t)LUT_COUNT * FB_POLY_SIZE * 8))) return rc;
  uint64_t* d_arena = (uint64_t*)ctx->arena.p;
  uint64_t* d_small = (uint64_t*)ctx->small.p;
  int32_t* d_i32 = (int32_t*)ctx->plan_i32.p;
  int64_t* d_i64 = (int64_t*)ctx->plan_i64.p;
  uint64_t* d_u64 = (uint64_t*)ctx->plan_u64.p;
  uint32_t* d_u32 = (uint32_t*)ctx->plan_u32.p;
  uint64_t* d_luts = (uint64_t*)ctx->regex_luts.p;
  cudaStream_t st = ctx->stream;
  cudaEvent_t ev0, ev1;
  FB_CUDA(ctx, cudaEventCreate(&ev0));
  FB_CUDA(ctx, cudaEventCreate(&ev1));
  auto cleanup = [&]() { cudaEventDestroy(ev0); cudaEventDestroy(ev1); };
#define RP_CUDA(call)                                                         \
  do {                                                                        \
    cudaError_t _e = (call);                                                  \
    if (_e != cudaSuccess) { cleanup(); return fb_cuda_fail(ctx, _e, #call); } \
  } while (0)
  dist = dist && ctx->comm && ctx->comm_world > 1 && m == 1;
  if (dist) {   // every rank uploads its slice of the content over PCIe; NVLink does the rest
    size_t lo, hi;
    fb_comm_slice(n_in_rows, ctx->comm_rank, ctx->comm_world, &lo, &hi);
    if (hi > lo)
      RP_CUDA(cudaMemcpyAsync(d_arena + lo * FB_LWE_BIG_WORDS, h_in + lo * FB_LWE_BIG_WORDS, (hi - lo) * row_bytes, cudaMemcpyHostToDevice, st));
  } else {
    RP_CUDA(cudaMemcpy2DAsync(d_arena, (size_t)plan.n_rows * row_bytes, h_in, n_in_rows * row_bytes, n_in_rows * row_bytes, m,
                              cudaMemcpyHostToDevice, st));
  }
  RP_CUDA(cudaMemcpyAsync(d_i32, d.i32.data(), d.i32.size() * 4, cudaMemcpyHostToDevice, st));
  RP_CUDA(cudaMemcpyAsync(d_i64, d.i64.data(), d.i64.size() * 8, cudaMemcpyHostToDevice, st));
  RP_CUDA(cudaMemcpyAsync(d_u64, d.u64.data(), d.u64.size() * 8, cudaMemcpyHostToDevice, st));
  RP_CUDA(cudaMemcpyAsync(d_u32, d.u32.data(), d.u32.size() * 4, cudaMemcpyHostToDevice, st));
  if (!ctx->regex_luts_ready) {   // the accumulator table is fixed: upload it once per context
    std::vector<uint64_t> luts;
    build_lut_table(luts);
    RP_CUDA(cudaMemcpyAsync(d_luts, luts.data(), luts.size() * 8, cudaMemcpyHostToDevice, st));
    RP_CUDA(cudaStreamSynchronize(st));
    ctx->regex_luts_ready = true;
  }
  RP_CUDA(cudaEventRecord(ev0, st));
  if (dist) {
    int rc = fb_comm_exchange_rows(ctx, d_arena, n_in_rows, FB_LWE_BIG_WORDS);
    if (rc) { cleanup(); return rc; }
  }
  for (auto& o : d.levels) {
    int rc;
    if (dist) {
      if (o.n_lin > 0) {
        rc = fb_run_lincomb(ctx, d_arena, d_i32 + o.lin_out, d_i32 + o.lin_off, d_i32 + o.lin_rows, d_i64 + o.lin_coef,
                            d_u64 + o.lin_const, o.n_lin);
        if (rc) { cleanup(); return rc; }
      }
      if (o.n_pbs > 0) {
        uint64_t* d_level_out = d_arena + (size_t)o.out_base * FB_LWE_BIG_WORDS;
        // a level that one GPU bootstraps in a single wave of SMs gains nothing from being cut up: every rank computes all of
        // it (the kernels are deterministic, the arenas stay identical) and the exchange is skipped
        if (o.n_pbs <= ctx->dist_shard_min) {
          rc = fb_run_keyswitch(ctx, d_arena, d_i32 + o.in_rows, d_small, o.n_pbs);
          if (rc) { cleanup(); return rc; }
          rc = fb_run_blind_rotate(ctx, d_small, d_luts, d_u32 + o.lut_idx, d_level_out, nullptr, o.n_pbs);
          if (rc) { cleanup(); return rc; }
          continue;
        }
        size_t lo, hi;
        fb_comm_slice((size_t)o.n_pbs, ctx->comm_rank, ctx->comm_world, &lo, &hi);
        if (hi > lo) {
          rc = fb_run_keyswitch(ctx, d_arena, d_i32 + o.in_rows + lo, d_small, (int)(hi - lo));
          if (rc) { cleanup(); return rc; }
          rc = fb_run_blind_rotate(ctx, d_small, d_luts, d_u32 + o.lut_idx + lo, d_level_out + lo * FB_LWE_BIG_WORDS, nullptr, (int)(hi - lo));
          if (rc) { cleanup(); return rc; }
        }
        rc = fb_comm_exchange_rows(ctx, d_level_out, (size_t)o.n_pbs, FB_LWE_BIG_WORDS);
        if (rc) { cleanup(); return rc; }
      }
      continue;
    }
    if (o.n_lin > 0) {
      rc = fb_run_lincomb(ctx, d_arena, d_i32 + o.lin_out, d_i32 + o.lin_off, d_i32 + o.lin_rows, d_i64 + o.lin_coef,
                          d_u64 + o.lin_const, o.n_lin);
      if (rc) { cleanup(); return rc; }
    }
    if (o.n_pbs > 0) {
      rc = fb_run_keyswitch(ctx, d_arena, d_i32 + o.in_rows, d_small, o.n_pbs);
      if (rc) { cleanup(); return rc; }
      rc = (m == 1) ? fb_run_blind_rotate(ctx, d_small, d_luts, d_u32 + o.lut_idx, d_arena + (size_t)o.out_base * FB_LWE_BIG_WORDS, nullptr, o.n_pbs)
                    : fb_run_blind_rotate(ctx, d_small, d_luts, d_u32 + o.lut_idx, d_arena, d_i32 + o.out_rows, o.n_pbs);
      if (rc) { cleanup(); return rc; }
    }
  }
  RP_CUDA(cudaEventRecord(ev1, st));
  std::memset(h_out_radix, 0, m * 4 * row_bytes);
  RP_CUDA(cudaMemcpy2DAsync(h_out_radix, 4 * row_bytes, d_arena + (size_t)plan.result_row * FB_LWE_BIG_WORDS, (size_t)plan.n_rows * row_bytes,
                            row_bytes, m, cudaMemcpyDeviceToHost, st));
  RP_CUDA(cudaStreamSynchronize(st));
  float ms = 0.f;
  RP_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
  if (gpu_ms) *gpu_ms = ms;
  cleanup();
#undef RP_CUDA
  return FB_OK;
}

}  // namespace

extern "C" int fb_parse_debug(const char* pattern, char* out, size_t cap) {
  if (!pattern || !out || cap == 0) return FB_ERR_ARG;
  RegExpr re;
  std::string err;
  int rc = parse(pattern, re, err);
  const std::string s = rc == FB_OK ? debug_fmt(re) : err;
  const size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
  std::memcpy(out, s.data(), n);
  out[n] = 0;
  return rc;
}

static PlanOptions options_of(uint32_t flags) {
  PlanOptions o;
  o.absorb = (flags & FB_PLAN_REFERENCE_SHAPED) == 0;
  return o;
}

extern "C" int fb_plan_stats(const char* pattern, size_t n_chars, uint32_t flags, fb_match_stats* stats) {
  if (!pattern || !stats) return FB_ERR_ARG;
  Plan plan;
  std::string err;
  int rc = build_plan(pattern, n_chars, 0, 1, options_of(flags), plan, err);
  if (rc != FB_OK) return rc;
  *stats = plan.stats;
  return FB_OK;
}

extern "C" int fb_plan_level_widths(const char* pattern, size_t n_chars, int rank, int world, uint32_t flags, int32_t* widths, size_t cap) {
  if (!pattern || (!widths && cap)) return FB_ERR_ARG;
  Plan plan;
  std::string err;
  int rc = build_plan(pattern, n_chars, rank, world, options_of(flags), plan, err);
  if (rc != FB_OK) return rc;
  int n = 0;
  for (auto& l : plan.levels) {
    if (l.in_rows.empty()) continue;
    if ((size_t)n < cap) widths[n] = (int32_t)l.in_rows.size();
    n++;
  }
  return n;
}

extern "C" int fb_plan_eval_plain(const char* pattern, const uint8_t* content, size_t n_chars, int rank, int world, uint32_t flags, int* result) {
  if (!pattern || !result || (!content && n_chars)) return FB_ERR_ARG;
  Plan plan;
  std::string err;
  int rc = build_plan(pattern, n_chars, rank, world, options_of(flags), plan, err);
  if (rc != FB_OK) return rc;
  return eval_plan_plain(plan, content, 0, result, err);
}

static int cached_plan(fb_ctx* ctx, const char* pattern, size_t n_chars, int rank, int world, std::shared_ptr<const Plan>& plan) {
  const std::string key = std::string(pattern) + '\n' + std::to_string(n_chars) + '/' + std::to_string(rank) + '/' + std::to_string(world) +
                          (ctx->plan_absorb ? "/absorbed" : "/ref-shaped");
  for (size_t i = 0; i < ctx->plan_cache.size(); i++)
    if (ctx->plan_cache[i].first == key) {
      plan = ctx->plan_cache[i].second;
      std::rotate(ctx->plan_cache.begin(), ctx->plan_cache.begin() + i, ctx->plan_cache.begin() + i + 1);
      return FB_OK;
    }
  auto fresh = std::make_shared<Plan>();
  std::string err;
  PlanOptions opt;
  opt.absorb = ctx->plan_absorb;
  opt.timing = ctx->plan_timing;
  int rc = build_plan(pattern, n_chars, rank, world, opt, *fresh, err);
  if (rc != FB_OK) return fb_fail(ctx, rc, err);
  plan = fresh;
  ctx->plan_cache.insert(ctx->plan_cache.begin(), std::make_pair(key, plan));
  if (ctx->plan_cache.size() > 8) ctx->plan_cache.pop_back();
  return FB_OK;
}

extern "C" int fb_has_match_shard(fb_ctx* ctx, const uint64_t* h_content, size_t n_chars, const char* pattern, int rank,
                                  int world, uint64_t* h_out, fb_match_stats* stats) {
  if (!ctx || !pattern || !h_out || (!h_content && n_chars)) return fb_fail(ctx, FB_ERR_ARG, "null argument");
  std::shared_ptr<const Plan> plan;
  int rc = cached_plan(ctx, pattern, n_chars, rank, world, plan);
  if (rc != FB_OK) return rc;
  double ms = 0;
  rc = run_plan(ctx, *plan, h_content, 4 * n_chars, h_out, &ms);
  if (rc != FB_OK) return rc;
  if (stats) {
    *stats = plan->stats;
    stats->gpu_ms = ms;
  }
  return FB_OK;
}

extern "C" int fb_has_match_many(fb_ctx* ctx, const uint64_t* h_contents, size_t n_contents, size_t n_chars, const char* pattern,
                                 uint64_t* h_out, fb_match_stats* stats) {
  if (!ctx || !pattern || (!h_out && n_contents) || (!h_contents && n_chars && n_contents)) return fb_fail(ctx, FB_ERR_ARG, "null argument");
  std::shared_ptr<const Plan> plan;
  int rc = cached_plan(ctx, pattern, n_chars, 0, 1, plan);
  if (rc != FB_OK) return rc;
  double ms = 0;
  if (n_contents > 0) {
    // bound the arena: instances are processed in groups of at most ~8 GiB of ciphertext rows
    const size_t per = std::max<size_t>(1, (size_t)plan->n_rows) * FB_LWE_BIG_WORDS * 8;
    const size_t group = std::max<size_t>(1, std::min<size_t>(n_contents, ((size_t)8 << 30) / per));
    for (size_t k0 = 0; k0 < n_contents; k0 += group) {
      const size_t mk = std::min(group, n_contents - k0);
      double part = 0;
      rc = run_plan(ctx, *plan, h_contents + k0 * 4 * n_chars * FB_LWE_BIG_WORDS, 4 * n_chars, h_out + k0 * 4 * FB_LWE_BIG_WORDS, &part, mk);
      if (rc != FB_OK) return rc;
      ms += part;
    }
  }
  if (stats) {
    *stats = plan->stats;
    stats->gpu_ms = ms;
  }
  return FB_OK;
}

// has_match across the communicator of fb_comm_init: collective -- every rank calls it with the same content and pattern
// and gets the result.  The plan is the single-GPU plan; its levels are sharded (run_plan, dist).
extern "C" int fb_has_match_dist(fb_ctx* ctx, const uint64_t* h_content, size_t n_chars, const char* pattern, uint64_t* h_out,
                                 fb_match_stats* stats) {
  if (!ctx || !pattern || !h_out || (!h_content && n_chars)) return fb_fail(ctx, FB_ERR_ARG, "null argument");
  if (!ctx->comm) return fb_fail(ctx, FB_ERR_ARG, "fb_has_match_dist needs fb_comm_init first");
  std::shared_ptr<const Plan> plan;
  int rc = cached_plan(ctx, pattern, n_chars, 0, 1, plan);
  if (rc != FB_OK) return rc;
  double ms = 0;
  rc = run_plan(ctx, *plan, h_content, 4 * n_chars, h_out, &ms, 1, true);
  if (rc != FB_OK) return rc;
  if (stats) {
    *stats = plan->stats;
    stats->gpu_ms = ms;
  }
  return FB_OK;
}

extern "C" int fb_has_match(fb_ctx* ctx, const uint64_t* h_content, size_t n_chars, const char* pattern, uint64_t* h_out,
                            fb_match_stats* stats) {
  return fb_has_match_shard(ctx, h_content, n_chars, pattern, 0, 1, h_out, stats);
}

extern "C" int fb_or_fold(fb_ctx* ctx, const uint64_t* h_in, size_t n, uint64_t* h_out) {
  if (!ctx || !h_out || (!h_in && n)) return fb_fail(ctx, FB_ERR_ARG, "null argument");
  if (n == 1) {
    std::memset(h_out, 0, sizeof(uint64_t) * 4 * FB_LWE_BIG_WORDS);
    std::memcpy(h_out, h_in, sizeof(uint64_t) * FB_LWE_BIG_WORDS);
    return FB_OK;
  }
  Plan plan;
  build_or_fold_plan(n, plan);
  return run_plan(ctx, plan, h_in, n, h_out, nullptr);
}
