// handles.cu -- op-level boundary for a host that keeps the reference's own executor: device-resident ciphertext
// arenas behind opaque handles, linear combinations and batched KS -> PBS on rows of an arena, no host round trip
// between the operations of a level or between levels.
//
// A maintainer who keeps /root/reference/src/regex/execution.rs:37-223 (`Execution`, its structural cache `with_cache`
// :212-222 and the six smart_* call sites :76,93,110,143,173,190) turns every cache-missing op into rows of an arena and
// flushes a level as one fb_lincomb + one fb_pbs_rows call.  fb_has_match (regex_api.cu) is exactly that, driven by the
// library's own planner; fb_plan_export hands the planner's levels to the host so that the two can be compared level by
// level (tests/test_gpu_parity.py::test_handle_api_replays_has_match_level_by_level).
#include <cuda_runtime.h>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include "context.h"
#include "regex_host.h"

namespace {

fb_arena* arena_of(fb_ctx* ctx, fb_handle h) {
  if (h == 0 || h > ctx->arenas.size()) return nullptr;
  fb_arena& a = ctx->arenas[h - 1];
  return a.p ? &a : nullptr;
}

// host index arrays -> one device scratch buffer per kind (grown as needed, reused by every call)
template <class T>
int stage(fb_ctx* ctx, fb_devbuf& b, const T* h, size_t n, T** d) {
  int rc = fb_reserve(ctx, b, (n + 1) * sizeof(T));
  if (rc) return rc;
  *d = static_cast<T*>(b.p);
  if (n) FB_CUDA(ctx, cudaMemcpyAsync(b.p, h, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  return FB_OK;
}

}  // namespace

extern "C" int fb_ct_alloc(fb_ctx* ctx, size_t rows, size_t row_words, fb_handle* out) {
  if (!ctx || !out || rows == 0 || (row_words != FB_LWE_BIG_WORDS && row_words != FB_POLY_SIZE))
    return fb_fail(ctx, FB_ERR_ARG, "fb_ct_alloc: rows of 2049 words (ciphertexts) or 2048 words (accumulator polynomials)");
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  fb_arena a;
  a.rows = rows;
  a.row_words = row_words;
  FB_CUDA(ctx, cudaMalloc(&a.p, rows * row_words * sizeof(uint64_t)));
  size_t slot = ctx->arenas.size();
  for (size_t i = 0; i < ctx->arenas.size(); i++)
    if (!ctx->arenas[i].p) { slot = i; break; }
  if (slot == ctx->arenas.size()) ctx->arenas.push_back(a);
  else ctx->arenas[slot] = a;
  *out = (fb_handle)(slot + 1);
  return FB_OK;
}

extern "C" int fb_ct_free(fb_ctx* ctx, fb_handle h) {
  if (!ctx) return FB_ERR_ARG;
  fb_arena* a = arena_of(ctx, h);
  if (!a) return fb_fail(ctx, FB_ERR_ARG, "bad handle");
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(a->p);
  *a = fb_arena{};
  return FB_OK;
}

extern "C" int fb_ct_upload(fb_ctx* ctx, fb_handle h, size_t first_row, const uint64_t* h_rows, size_t count) {
  if (!ctx || (!h_rows && count)) return fb_fail(ctx, FB_ERR_ARG, "null argument");
  fb_arena* a = arena_of(ctx, h);
  if (!a || first_row > a->rows || count > a->rows - first_row) return fb_fail(ctx, FB_ERR_ARG, "bad handle or row range");
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  // pageable host memory: the copy is staged by the runtime before the call returns; pinned memory: the caller must not
  // touch the buffer before fb_sync
  FB_CUDA(ctx, cudaMemcpyAsync(a->p + first_row * a->row_words, h_rows, count * a->row_words * 8, cudaMemcpyHostToDevice, ctx->stream));
  return FB_OK;
}

extern "C" int fb_ct_download(fb_ctx* ctx, fb_handle h, size_t first_row, size_t count, uint64_t* h_rows) {
  if (!ctx || (!h_rows && count)) return fb_fail(ctx, FB_ERR_ARG, "null argument");
  fb_arena* a = arena_of(ctx, h);
  if (!a || first_row > a->rows || count > a->rows - first_row) return fb_fail(ctx, FB_ERR_ARG, "bad handle or row range");
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  FB_CUDA(ctx, cudaMemcpyAsync(h_rows, a->p + first_row * a->row_words, count * a->row_words * 8, cudaMemcpyDeviceToHost, ctx->stream));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FB_OK;
}

extern "C" int fb_lincomb(fb_ctx* ctx, fb_handle h, const int32_t* h_out_rows, const int32_t* h_term_off, const int32_t* h_term_rows,
                          const int64_t* h_term_coef, const uint64_t* h_body_const, size_t n_out) {
  if (!ctx) return FB_ERR_ARG;
  if (n_out == 0) return FB_OK;
  if (!h_out_rows || !h_term_off || !h_body_const) return fb_fail(ctx, FB_ERR_ARG, "null argument");
  fb_arena* a = arena_of(ctx, h);
  if (!a || a->row_words != FB_LWE_BIG_WORDS) return fb_fail(ctx, FB_ERR_ARG, "bad handle");
  const size_t n_terms = (size_t)h_term_off[n_out];
  if (n_terms && (!h_term_rows || !h_term_coef)) return fb_fail(ctx, FB_ERR_ARG, "null argument");
  for (size_t o = 0; o < n_out; o++)
    if (h_out_rows[o] < 0 || (size_t)h_out_rows[o] >= a->rows || h_term_off[o] < 0 || h_term_off[o] > h_term_off[o + 1])
      return fb_fail(ctx, FB_ERR_ARG, "fb_lincomb: output row or term offsets out of range");
  for (size_t t = 0; t < n_terms; t++)
    if (h_term_rows[t] < 0 || (size_t)h_term_rows[t] >= a->rows) return fb_fail(ctx, FB_ERR_ARG, "fb_lincomb: term row out of range");
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  // (the index buffers of the previous call may still be read by its kernels: same stream, so the copies below queue behind them)
  int32_t *d_out = nullptr, *d_off = nullptr, *d_rows = nullptr;
  int64_t* d_coef = nullptr;
  uint64_t* d_const = nullptr;
  std::vector<int32_t> i32;
  i32.insert(i32.end(), h_out_rows, h_out_rows + n_out);
  i32.insert(i32.end(), h_term_off, h_term_off + n_out + 1);
  if (n_terms) i32.insert(i32.end(), h_term_rows, h_term_rows + n_terms);
  int rc;
  int32_t* d_i32 = nullptr;
  if ((rc = stage(ctx, ctx->op_i32, i32.data(), i32.size(), &d_i32))) return rc;
  d_out = d_i32;
  d_off = d_i32 + n_out;
  d_rows = d_off + n_out + 1;
  if ((rc = stage(ctx, ctx->op_i64, h_term_coef, n_terms, &d_coef))) return rc;
  if ((rc = stage(ctx, ctx->op_u64, h_body_const, n_out, &d_const))) return rc;
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // i32 is a host temporary
  return fb_run_lincomb(ctx, a->p, d_out, d_off, d_rows, d_coef, d_const, (int)n_out);
}

extern "C" int fb_pbs_rows(fb_ctx* ctx, fb_handle h, const int32_t* h_in_rows, fb_handle luts, const uint32_t* h_lut_idx, size_t count,
                           size_t out_row_base) {
  if (!ctx) return FB_ERR_ARG;
  if (count == 0) return FB_OK;
  if (!h_in_rows || !h_lut_idx) return fb_fail(ctx, FB_ERR_ARG, "null argument");
  if (!ctx->have_key) return fb_fail(ctx, FB_ERR_NO_KEY, "server key not loaded");
  fb_arena* a = arena_of(ctx, h);
  fb_arena* l = arena_of(ctx, luts);
  if (!a || a->row_words != FB_LWE_BIG_WORDS || !l || l->row_words != FB_POLY_SIZE) return fb_fail(ctx, FB_ERR_ARG, "bad handle");
  if (out_row_base > a->rows || count > a->rows - out_row_base) return fb_fail(ctx, FB_ERR_ARG, "fb_pbs_rows: output rows out of range");
  for (size_t b = 0; b < count; b++)
    if (h_in_rows[b] < 0 || (size_t)h_in_rows[b] >= a->rows || h_lut_idx[b] >= l->rows) return fb_fail(ctx, FB_ERR_ARG, "fb_pbs_rows: row or LUT index out of range");
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc;
  int32_t* d_in = nullptr;
  uint32_t* d_idx = nullptr;
  if ((rc = stage(ctx, ctx->op_rows, h_in_rows, count, &d_in))) return rc;
  if ((rc = stage(ctx, ctx->op_u32, h_lut_idx, count, &d_idx))) return rc;
  if ((rc = fb_reserve(ctx, ctx->small, count * FB_LWE_SMALL_WORDS * 8))) return rc;
  if ((rc = fb_run_keyswitch(ctx, a->p, d_in, (uint64_t*)ctx->small.p, (int)count))) return rc;
  return fb_run_blind_rotate(ctx, (const uint64_t*)ctx->small.p, l->p, d_idx, a->p + out_row_base * FB_LWE_BIG_WORDS, nullptr, (int)count);
}

// ---- the library's own plan, for a host that wants to drive (or check) it level by level -----------------------
// stream of int64: n_rows, result_kind, result_row, n_levels, then per level:
//   n_lin, n_terms, n_pbs, out_row_base, lin_out_rows[n_lin], lin_term_off[n_lin + 1], lin_term_rows[n_terms],
//   lin_coef[n_terms], lin_const[n_lin] (u64 bit patterns), in_rows[n_pbs], lut_idx[n_pbs]
extern "C" int fb_plan_export(const char* pattern, size_t n_chars, uint32_t flags, int64_t* out, size_t cap, size_t* n_words) {
  if (!pattern || !n_words) return FB_ERR_ARG;
  fbre::Plan plan;
  fbre::PlanOptions opt;
  opt.absorb = (flags & FB_PLAN_REFERENCE_SHAPED) == 0;
  std::string err;
  int rc = fbre::build_plan(pattern, n_chars, 0, 1, opt, plan, err);
  if (rc != FB_OK) return rc;
  std::vector<int64_t> w;
  w.push_back(plan.n_rows);
  w.push_back(plan.result_kind);
  w.push_back(plan.result_row);
  w.push_back((int64_t)plan.levels.size());
  for (auto& l : plan.levels) {
    w.push_back((int64_t)l.lin_out_rows.size());
    w.push_back((int64_t)l.lin_term_rows.size());
    w.push_back((int64_t)l.in_rows.size());
    w.push_back(l.out_row_base);
    for (auto v : l.lin_out_rows) w.push_back(v);
    if (l.lin_out_rows.empty()) w.push_back(0);
    else for (auto v : l.lin_term_off) w.push_back(v);
    for (auto v : l.lin_term_rows) w.push_back(v);
    for (auto v : l.lin_coef) w.push_back(v);
    for (auto v : l.lin_const) w.push_back((int64_t)v);
    for (auto v : l.in_rows) w.push_back(v);
    for (auto v : l.lut_idx) w.push_back((int64_t)v);
  }
  *n_words = w.size();
  if (!out) return FB_OK;
  if (cap < w.size()) return FB_ERR_ARG;
  std::memcpy(out, w.data(), w.size() * sizeof(int64_t));
  return FB_OK;
}

// the accumulator table fb_has_match bootstraps through (regex_host.h LutId order): h_out[FB_REGEX_LUTS][2048]
extern "C" int fb_regex_lut_table(uint64_t* h_out) {
  if (!h_out) return FB_ERR_ARG;
  for (uint32_t id = 0; id < fbre::LUT_COUNT; id++) {
    uint64_t f16[16];
    for (uint32_t x = 0; x < 16; x++) f16[x] = fbre::lut_value(id, x);
    fb_make_lut(f16, h_out + (size_t)id * FB_POLY_SIZE);
  }
  return FB_OK;
}
