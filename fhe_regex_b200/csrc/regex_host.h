// regex_host.h -- host side of the hot path, mirroring the reference's own interfaces:
//   parse()            src/regex/parser.rs:146-184      RegExpr           parser.rs:9-41
//   build_branches()   src/regex/engine.rs:45-214       has_match()       engine.rs:8-42
//   Execution          src/regex/execution.rs:37-223 (ct_eq/ct_ge/ct_le/ct_and/ct_or/ct_not/ct_constant,
//                      structural cache, ct_operations_count / cache_hits)
// and, below that surface, the part that is new: lowering of the executed boolean DAG to batched
// programmable bootstraps (level-synchronous plan) for the CUDA backend.
#pragma once
#include <stdint.h>
#include <string>
#include <unordered_map>
#include <vector>
#include "../../include/fhe_b200.h"

namespace fbre {

// ---- RegExpr (parser.rs:9-41) -----------------------------------------------------------------
struct RegExpr {
  enum Kind { SOF, EOF_, Char, AnyChar, Between, Range, Not, Either, Optional, Repeated, Seq };
  Kind kind = Seq;
  uint8_t c = 0, from = 0, to = 0;   // Char / Between
  std::vector<uint8_t> cs;           // Range
  std::vector<RegExpr> sub;          // Not/Optional/Repeated: 1; Either: 2; Seq: n
  bool has_lo = false, has_hi = false;
  size_t lo = 0, hi = 0;             // Repeated at_least / at_most
};

// FB_OK, FB_ERR_PARSE (anyhow error of parse) or FB_ERR_PANIC (reference panics, e.g. /a{}/)
int parse(const std::string& pattern, RegExpr& out, std::string& err);
std::string debug_fmt(const RegExpr& re);  // impl Debug for RegExpr, parser.rs:87-144

// ---- lowered plan -----------------------------------------------------------------------------
// LUT ids of the fixed accumulator table uploaded once per context
enum LutId : uint32_t {
  LUT_NIB_EQ = 0,    // +v: x == v            (v < 16) on a packed 4-bit nibble
  LUT_NIB_GT = 16,   // +v: x > v
  LUT_SUM_EQ = 32,   // +k: x == k            (k-ary AND of booleans summed)
  LUT_GE1 = 48,      // x >= 1                (k-ary OR)
  LUT_GE2 = 49,      // x >= 2                (gt combine over 2*G1 + E1 + G0)
  LUT_LT2 = 50,      // x < 2                 (le combine)
  LUT_COUNT = 51
};
uint64_t lut_value(uint32_t lut_id, uint32_t x);  // f(x) in {0,1} for x < 16

struct LinTerm { int32_t row; int64_t coef; };
struct PlanLevel {
  // linear combinations evaluated before this level's PBS batch (sums of booleans, packing)
  std::vector<int32_t> lin_out_rows, lin_term_off, lin_term_rows;
  std::vector<int64_t> lin_coef;
  std::vector<uint64_t> lin_const;
  // PBS batch: input row per PBS, LUT per PBS; outputs are rows out_row_base .. out_row_base + n
  std::vector<int32_t> in_rows;
  std::vector<uint32_t> lut_idx;
  int32_t out_row_base = 0;
};
struct Plan {
  size_t n_chars = 0;
  int32_t n_rows = 0;               // arena rows: [0, 4n) content blocks, then packs, PBS outputs, scratch
  std::vector<PlanLevel> levels;    // levels[0] has only the packing lincombs (no PBS) when n_chars > 0
  int result_kind = 0;              // 0: constant false, 1: constant true, 2: arena row result_row
  int32_t result_row = -1;
  fb_match_stats stats{};
};

// Parses, enumerates the variants of every start offset i with i % world == rank, runs the reference's
// executor bookkeeping and lowers to a plan.  Returns FB_OK / FB_ERR_PARSE / FB_ERR_PANIC.
struct PlanOptions {
  bool absorb = true;    // false: reference-shaped plan, every variant the reference enumerates is evaluated
  bool timing = false;   // print the planner's phase times to stderr
};
int build_plan(const std::string& pattern, size_t n_chars, int rank, int world, const PlanOptions& opt, Plan& plan, std::string& err);
// plan for OR-folding n boolean ciphertexts placed in arena rows 0..n-1
void build_or_fold_plan(size_t n, Plan& plan);

// plaintext dry run of a plan (host only; validates the lowering, never touches ciphertexts)
int eval_plan_plain(const Plan& plan, const uint8_t* content, size_t n_in_rows_are_blocks, int* result, std::string& err);

}  // namespace fbre
