// regex_host.cpp -- parser, variant generator, executor bookkeeping and PBS lowering (host, C++).
// See regex_host.h for the reference interfaces each part mirrors.  Written from the reference's
// behaviour (incl. its quirks), not from its code: the reference evaluates Rc closures
// against deep-cloned provenance trees; here variants are integer thunks and provenance is hash-consed.
#include "regex_host.h"
#include <chrono>
#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>

namespace fbre {

// =================================================================================================
// parser (grammar: parser.rs:208-351; entry: parser.rs:146-184; case folding: parser.rs:43-81)
// =================================================================================================
namespace {

struct PanicEx { std::string msg; };

struct Parser {
  const std::string& s;
  explicit Parser(const std::string& str) : s(str) {}
  static bool is_letter(uint8_t b) { return (b >= 'A' && b <= 'Z') || (b >= 'a' && b <= 'z'); }
  static bool is_digit(uint8_t b) { return b >= '0' && b <= '9'; }
  static bool non_escapable(uint8_t b) { return b && std::strchr("&;:,`~-_!@#%'\"", b) != nullptr; }  // parser.rs:238-240
  bool at(size_t i, char ch) const { return i < s.size() && s[i] == ch; }

  // regex := term '|' regex | term
  bool regex(size_t i, RegExpr& out, size_t& end) {
    RegExpr l;
    size_t j;
    if (term(i, l, j) && at(j, '|')) {
      RegExpr r;
      size_t k;
      if (regex(j + 1, r, k)) {
        out = RegExpr();
        out.kind = RegExpr::Either;
        out.sub.push_back(std::move(l));
        out.sub.push_back(std::move(r));
        end = k;
        return true;
      }
    }
    return term(i, out, end);
  }
  // term := factor*  (one factor -> itself, otherwise Seq, possibly empty)
  bool term(size_t i, RegExpr& out, size_t& end) {
    std::vector<RegExpr> xs;
    for (;;) {
      RegExpr f;
      size_t j;
      if (!factor(i, f, j)) break;
      xs.push_back(std::move(f));
      i = j;
    }
    if (xs.size() == 1) {
      out = std::move(xs[0]);
    } else {
      out = RegExpr();
      out.kind = RegExpr::Seq;
      out.sub = std::move(xs);
    }
    end = i;
    return true;
  }
  // factor := atom '?' | repeated | atom
  bool factor(size_t i, RegExpr& out, size_t& end) {
    RegExpr a;
    size_t j;
    if (atom(i, a, j) && at(j, '?')) {
      out = RegExpr();
      out.kind = RegExpr::Optional;
      out.sub.push_back(std::move(a));
      end = j + 1;
      return true;
    }
    if (repeated(i, out, end)) return true;
    return atom(i, out, end);
  }
  bool atom(size_t i, RegExpr& out, size_t& end) {
    if (i >= s.size()) return false;
    const uint8_t b = (uint8_t)s[i];
    out = RegExpr();
    if (b == '.') { out.kind = RegExpr::AnyChar; end = i + 1; return true; }
    if (b == '\\' && i + 1 < s.size()) { out.kind = RegExpr::Char; out.c = (uint8_t)s[i + 1]; end = i + 2; return true; }
    if (is_letter(b) || non_escapable(b)) { out.kind = RegExpr::Char; out.c = b; end = i + 1; return true; }
    if (b == '[') {
      size_t j;
      if (!range(i + 1, out, j) || !at(j, ']')) return false;
      end = j + 1;
      return true;
    }
    if (b == '(') {
      size_t j;
      if (!regex(i + 1, out, j) || !at(j, ')')) return false;
      end = j + 1;
      return true;
    }
    return false;
  }
  // range := '^' range | letter '-' letter | letter+
  bool range(size_t i, RegExpr& out, size_t& end) {
    if (at(i, '^')) {
      RegExpr r;
      size_t j;
      if (!range(i + 1, r, j)) return false;
      out = RegExpr();
      out.kind = RegExpr::Not;
      out.sub.push_back(std::move(r));
      end = j;
      return true;
    }
    if (i + 2 < s.size() && is_letter(s[i]) && s[i + 1] == '-' && is_letter(s[i + 2])) {
      out = RegExpr();
      out.kind = RegExpr::Between;
      out.from = (uint8_t)s[i];
      out.to = (uint8_t)s[i + 2];
      end = i + 3;
      return true;
    }
    size_t j = i;
    while (j < s.size() && is_letter(s[j])) j++;
    if (j == i) return false;
    out = RegExpr();
    out.kind = RegExpr::Range;
    out.cs.assign(s.begin() + i, s.begin() + j);
    end = j;
    return true;
  }
  static size_t parse_digits(const std::string& d) {
    if (d.empty()) throw PanicEx{"parse_digits(\"\").unwrap() panics (parser.rs:349-351)"};
    size_t v = 0;
    for (char ch : d) {
      if (v > (SIZE_MAX - 9) / 10) throw PanicEx{"repeat count overflows usize (parser.rs:349-351)"};
      v = v * 10 + (size_t)(ch - '0');
    }
    return v;
  }
  bool repeated(size_t i, RegExpr& out, size_t& end) {
    RegExpr a;
    size_t j;
    if (!atom(i, a, j)) return false;
    auto make = [&](bool has_lo, size_t lo, bool has_hi, size_t hi, size_t e) {
      out = RegExpr();
      out.kind = RegExpr::Repeated;
      out.sub.push_back(std::move(a));
      out.has_lo = has_lo; out.lo = lo; out.has_hi = has_hi; out.hi = hi;
      end = e;
      return true;
    };
    if (at(j, '*')) return make(false, 0, false, 0, j + 1);
    if (at(j, '+')) return make(true, 1, false, 0, j + 1);
    if (!at(j, '{')) return false;
    size_t k = j + 1, k1 = k;
    while (k1 < s.size() && is_digit(s[k1])) k1++;
    const std::string d1 = s.substr(k, k1 - k);
    if (at(k1, '}')) {
      size_t n = parse_digits(d1);
      return make(true, n, true, n, k1 + 1);
    }
    if (!at(k1, ',')) return false;
    size_t k2 = k1 + 1, k3 = k2;
    while (k3 < s.size() && is_digit(s[k3])) k3++;
    if (!at(k3, '}')) return false;
    const std::string d2 = s.substr(k2, k3 - k2);
    const bool hl = !d1.empty(), hh = !d2.empty();
    return make(hl, hl ? parse_digits(d1) : 0, hh, hh ? parse_digits(d2) : 0, k3 + 1);
  }
};

void fold_case(RegExpr& re) {  // parser.rs:43-81: only Char is rewritten
  if (re.kind == RegExpr::Char) {
    const uint8_t c = re.c;
    re.kind = RegExpr::Range;
    re.cs.clear();
    re.cs.push_back(c);
    if (c >= 'a' && c <= 'z') re.cs.push_back((uint8_t)(c - 32));
    else if (c >= 'A' && c <= 'Z') re.cs.push_back((uint8_t)(c + 32));
    return;
  }
  if (re.kind == RegExpr::Not || re.kind == RegExpr::Either || re.kind == RegExpr::Optional ||
      re.kind == RegExpr::Repeated || re.kind == RegExpr::Seq)
    for (auto& x : re.sub) fold_case(x);
}

}  // namespace

int parse(const std::string& pattern, RegExpr& out, std::string& err) {
  try {
    Parser p(pattern);
    size_t i = 0;
    if (!p.at(i, '/')) { err = "failed to parse regular expression: expected '/'"; return FB_ERR_PARSE; }
    i++;
    const bool sof = p.at(i, '^');
    if (sof) i++;
    RegExpr re;
    size_t j;
    if (!p.regex(i, re, j)) { err = "failed to parse regular expression"; return FB_ERR_PARSE; }
    i = j;
    const bool eof = p.at(i, '$');
    if (eof) i++;
    if (!p.at(i, '/')) { err = "failed to parse regular expression: expected closing '/'"; return FB_ERR_PARSE; }
    i++;
    if (sof || eof) {
      RegExpr seq;
      seq.kind = RegExpr::Seq;
      if (sof) { RegExpr a; a.kind = RegExpr::SOF; seq.sub.push_back(a); }
      seq.sub.push_back(std::move(re));
      if (eof) { RegExpr a; a.kind = RegExpr::EOF_; seq.sub.push_back(a); }
      re = std::move(seq);
    }
    if (p.at(i, 'i')) { i++; fold_case(re); }
    if (i != pattern.size()) {
      err = "failed to parse regular expression, unexpected token at start of: " + pattern.substr(i);
      return FB_ERR_PARSE;
    }
    out = std::move(re);
    return FB_OK;
  } catch (const PanicEx& e) {
    err = e.msg;
    return FB_ERR_PANIC;
  }
}

std::string debug_fmt(const RegExpr& re) {
  switch (re.kind) {
    case RegExpr::SOF: return "^";
    case RegExpr::EOF_: return "$";
    case RegExpr::Char: return std::string(1, (char)re.c);
    case RegExpr::AnyChar: return ".";
    case RegExpr::Not: return "[^" + debug_fmt(re.sub[0]) + "]";
    case RegExpr::Between: return std::string("[") + (char)re.from + "->" + (char)re.to + "]";
    case RegExpr::Range: return "[" + std::string(re.cs.begin(), re.cs.end()) + "]";
    case RegExpr::Either: return "(" + debug_fmt(re.sub[0]) + "|" + debug_fmt(re.sub[1]) + ")";
    case RegExpr::Repeated:
      return debug_fmt(re.sub[0]) + "{" + (re.has_lo ? std::to_string(re.lo) : "*") + "," +
             (re.has_hi ? std::to_string(re.hi) : "*") + "}";
    case RegExpr::Optional: return debug_fmt(re.sub[0]) + "?";
    case RegExpr::Seq: {
      std::string o = "<";
      for (auto& x : re.sub) o += debug_fmt(x);
      return o + ">";
    }
  }
  return "";
}

// =================================================================================================
// variant generator (engine.rs:45-214): branches are integer thunks instead of Rc closures
// =================================================================================================
namespace {

struct Thunk {
  enum K : uint8_t { TRUE_, CHAR, NOT, BETWEEN, RANGE, ANDTHEN } k;
  int32_t a, b;  // CHAR: pos, c | NOT: thunk | BETWEEN: pos, from|to<<8 | RANGE: pos, range id | ANDTHEN: prev, x
};
typedef std::pair<int32_t, size_t> Branch;  // (thunk, next content position)

struct Builder {
  size_t n;
  std::vector<Thunk> thunks;
  std::vector<std::vector<uint8_t>> ranges;
  int32_t true_thunk;
  explicit Builder(size_t n_chars) : n(n_chars) { true_thunk = add(Thunk::TRUE_, 0, 0); }
  int32_t add(Thunk::K k, int32_t a, int32_t b) {
    thunks.push_back(Thunk{k, a, b});
    return (int32_t)thunks.size() - 1;
  }
  std::vector<Branch> seq_continue(std::vector<Branch> conts, const RegExpr& re_x) {
    std::vector<Branch> nxt;
    for (auto& bp : conts)
      for (auto& bx : build(re_x, bp.second)) nxt.emplace_back(add(Thunk::ANDTHEN, bp.first, bx.first), bx.second);
    return nxt;
  }
  std::vector<Branch> build(const RegExpr& re, size_t c_pos) {
    std::vector<Branch> res;
    if (re.kind == RegExpr::SOF) {  // engine.rs:52-58
      if (c_pos == 0) res.emplace_back(true_thunk, c_pos);
      return res;
    }
    if (re.kind == RegExpr::EOF_) {  // engine.rs:59-65
      if (c_pos == n) res.emplace_back(true_thunk, c_pos);
      return res;
    }
    if (c_pos >= n) return res;  // engine.rs:69-71
    switch (re.kind) {
      case RegExpr::Char:
        res.emplace_back(add(Thunk::CHAR, (int32_t)c_pos, re.c), c_pos + 1);
        return res;
      case RegExpr::AnyChar:
        res.emplace_back(true_thunk, c_pos + 1);
        return res;
      case RegExpr::Not:
        for (auto& b : build(re.sub[0], c_pos)) res.emplace_back(add(Thunk::NOT, b.first, 0), b.second);
        return res;
      case RegExpr::Either: {
        res = build(re.sub[0], c_pos);
        auto r = build(re.sub[1], c_pos);
        res.insert(res.end(), r.begin(), r.end());
        return res;
      }
      case RegExpr::Between:
        res.emplace_back(add(Thunk::BETWEEN, (int32_t)c_pos, (int32_t)re.from | ((int32_t)re.to << 8)), c_pos + 1);
        return res;
      case RegExpr::Range:
        ranges.push_back(re.cs);
        res.emplace_back(add(Thunk::RANGE, (int32_t)c_pos, (int32_t)ranges.size() - 1), c_pos + 1);
        return res;
      case RegExpr::Repeated: {  // engine.rs:127-183
        const size_t at_least = re.has_lo ? re.lo : 0;
        const size_t at_most = re.has_hi ? re.hi : n - c_pos;
        if (at_least > at_most) return res;
        if (at_least == 0) res.emplace_back(true_thunk, c_pos);
        // Seq of max(1, at_least) copies; more copies than remaining characters can never match
        // (every non-anchor node consumes >= 1 position or yields nothing), so cap the fold early.
        std::vector<Branch> last = build(re.sub[0], c_pos);
        for (size_t t = 1; t < std::max<size_t>(1, at_least) && !last.empty(); t++) last = seq_continue(std::move(last), re.sub[0]);
        res.insert(res.end(), last.begin(), last.end());
        for (size_t t = at_least + 1; t < at_most + 1 && !last.empty(); t++) {
          last = seq_continue(std::move(last), re.sub[0]);
          res.insert(res.end(), last.begin(), last.end());
        }
        return res;
      }
      case RegExpr::Optional:
        res = build(re.sub[0], c_pos);
        res.emplace_back(true_thunk, c_pos);
        return res;
      case RegExpr::Seq: {  // engine.rs:189-211
        if (re.sub.empty()) throw PanicEx{"Seq{re_xs: []}: index out of bounds (engine.rs:189-190)"};
        res = build(re.sub[0], c_pos);
        for (size_t t = 1; t < re.sub.size(); t++) res = seq_continue(std::move(res), re.sub[t]);
        return res;
      }
      default:
        throw PanicEx{"unmatched regex variant"};
    }
  }
};

// =================================================================================================
// Execution (execution.rs:37-223) with hash-consed provenance keys + boolean value DAG
// =================================================================================================
struct Triple {
  int32_t t, a, b;
  bool operator==(const Triple& o) const { return t == o.t && a == o.a && b == o.b; }
};
struct TripleHash {
  size_t operator()(const Triple& x) const {
    uint64_t h = (uint64_t)(uint32_t)x.t * 0x9E3779B97F4A7C15ull;
    h ^= ((uint64_t)(uint32_t)x.a + 0x7F4A7C15ull + (h << 6) + (h >> 2));
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= ((uint64_t)(uint32_t)x.b + 0x94D049BBull + (h << 6) + (h >> 2));
    h *= 0x94D049BB133111EBull;
    return (size_t)(h ^ (h >> 31));
  }
};
struct Interner {  // hash-consing table: open addressing over a power-of-two slot array (ids index `items`)
  std::vector<int32_t> slots;
  std::vector<Triple> items;
  size_t mask = 0;
  Interner() { slots.assign(1 << 12, -1); mask = slots.size() - 1; }
  void grow() {
    std::vector<int32_t> bigger(slots.size() * 2, -1);
    const size_t m = bigger.size() - 1;
    TripleHash h;
    for (int32_t id = 0; id < (int32_t)items.size(); id++) {
      size_t p = h(items[id]) & m;
      while (bigger[p] >= 0) p = (p + 1) & m;
      bigger[p] = id;
    }
    slots.swap(bigger);
    mask = m;
  }
  int32_t get(int32_t t, int32_t a, int32_t b) {
    const Triple k{t, a, b};
    size_t p = TripleHash()(k) & mask;
    while (slots[p] >= 0) {
      if (items[slots[p]] == k) return slots[p];
      p = (p + 1) & mask;
    }
    const int32_t id = (int32_t)items.size();
    items.push_back(k);
    slots[p] = id;
    if (items.size() * 2 > slots.size()) grow();
    return id;
  }
};

enum KeyTag { K_CONST, K_POS, K_AND, K_OR, K_EQ, K_GE, K_LE, K_NOT };
enum ValTag { V_CONST, V_POS, V_EQ, V_GT, V_LE, V_AND, V_OR, V_NOT };  // V_EQ/GT/LE: a = pos, b = constant byte

struct Res { int32_t val, key; };

struct Execution {
  Interner keys, vals;
  std::vector<int32_t> cache;  // key id -> value id or -1 (HashMap<Executed, RadixCiphertext>)
  uint64_t ct_ops = 0, cache_hits = 0, calls = 0;
  uint64_t by_type[8] = {0};

  int key_const(int32_t key) const {  // Executed::get_trivial_constant
    const Triple& k = keys.items[key];
    return k.t == K_CONST ? k.a : -1;
  }
  Res ct_constant(uint8_t c) { return Res{vals.get(V_CONST, c, 0), keys.get(K_CONST, c, 0)}; }
  Res ct_true() { return ct_constant(1); }
  Res ct_false() { return ct_constant(0); }
  Res ct_pos(int32_t at) { return Res{vals.get(V_POS, at, 0), keys.get(K_POS, at, 0)}; }

  // value-level constructors with the obvious boolean simplifications (decrypt-equivalent)
  int32_t v_const_of(int32_t v) const { return vals.items[v].t == V_CONST ? vals.items[v].a : -1; }
  int32_t v_cmp(int tag, int32_t a, int32_t b) {
    // the reference only ever compares a content position against a pattern constant (engine.rs:77,103-106,117-119)
    return vals.get(tag, vals.items[a].a, vals.items[b].a);
  }
  int32_t v_not(int32_t a) {
    int c = v_const_of(a);
    if (c >= 0) return vals.get(V_CONST, c ^ 1, 0);
    if (vals.items[a].t == V_NOT) return vals.items[a].a;
    return vals.get(V_NOT, a, 0);
  }
  int32_t v_and(int32_t a, int32_t b) {
    int ca = v_const_of(a), cb = v_const_of(b);
    if (ca >= 0) return (ca & 1) ? b : a;
    if (cb >= 0) return (cb & 1) ? a : b;
    if (a == b) return a;
    return vals.get(V_AND, std::min(a, b), std::max(a, b));
  }
  int32_t v_or(int32_t a, int32_t b) {
    int ca = v_const_of(a), cb = v_const_of(b);
    if (ca >= 0) return (ca & 1) ? a : b;
    if (cb >= 0) return (cb & 1) ? b : a;
    if (a == b) return a;
    return vals.get(V_OR, std::min(a, b), std::max(a, b));
  }

  template <class F>
  Res with_cache(int32_t key, int type, F f) {  // execution.rs:212-222
    calls++;
    if ((size_t)key >= cache.size()) cache.resize(std::max<size_t>((size_t)key + 1, cache.size() * 2), -1);
    if (cache[key] >= 0) {
      cache_hits++;
      return Res{cache[key], key};
    }
    ct_ops++;
    by_type[type]++;
    int32_t v = f();
    cache[key] = v;
    return Res{v, key};
  }
  Res ct_eq(Res a, Res b) { return with_cache(keys.get(K_EQ, a.key, b.key), K_EQ, [&] { return v_cmp(V_EQ, a.val, b.val); }); }
  Res ct_ge(Res a, Res b) {  // execution.rs:93: smart_gt, i.e. strictly greater
    return with_cache(keys.get(K_GE, a.key, b.key), K_GE, [&] { return v_cmp(V_GT, a.val, b.val); });
  }
  Res ct_le(Res a, Res b) { return with_cache(keys.get(K_LE, a.key, b.key), K_LE, [&] { return v_cmp(V_LE, a.val, b.val); }); }
  Res ct_and(Res a, Res b) {  // execution.rs:115-146
    const int32_t key = keys.get(K_AND, a.key, b.key);
    const int ca = key_const(a.key), cb = key_const(b.key);
    if (ca == 1) return Res{b.val, key};
    if (ca == 0) return Res{a.val, key};
    if (cb == 1) return Res{a.val, key};
    if (cb == 0) return Res{b.val, key};
    return with_cache(key, K_AND, [&] { return v_and(a.val, b.val); });
  }
  Res ct_or(Res a, Res b) {  // execution.rs:148-176
    const int32_t key = keys.get(K_OR, a.key, b.key);
    const int ca = key_const(a.key), cb = key_const(b.key);
    if (ca == 1) return Res{a.val, key};
    if (cb == 1) return Res{b.val, key};
    if (ca == 0 && cb == 0) return Res{a.val, key};
    return with_cache(key, K_OR, [&] { return v_or(a.val, b.val); });
  }
  Res ct_not(Res a) {  // execution.rs:178-195: smart_bitxor(a, trivial 1)
    return with_cache(keys.get(K_NOT, a.key, 0), K_NOT, [&] { return v_not(a.val); });
  }
};

struct ThunkEval {
  const Builder& B;
  Execution& ex;
  struct Memo { Res res; uint64_t calls; bool done; };
  std::vector<Memo> memo;
  ThunkEval(const Builder& b, Execution& e) : B(b), ex(e), memo(b.thunks.size(), Memo{{0, 0}, 0, false}) {}

  // Re-running a closure in the reference repeats exactly the same with_cache calls, all of which hit;
  // so a thunk is walked once and re-evaluations only replay its call count into cache_hits.
  Res eval(int32_t t) {
    Memo& m = memo[t];
    if (m.done) {
      ex.calls += m.calls;
      ex.cache_hits += m.calls;
      return m.res;
    }
    const uint64_t before = ex.calls;
    const Thunk th = B.thunks[t];
    Res r{0, 0};
    switch (th.k) {
      case Thunk::TRUE_: r = ex.ct_true(); break;
      case Thunk::CHAR: r = ex.ct_eq(ex.ct_pos(th.a), ex.ct_constant((uint8_t)th.b)); break;  // engine.rs:74-80
      case Thunk::NOT: { Res a = eval(th.a); r = ex.ct_not(a); break; }                        // engine.rs:82-93
      case Thunk::BETWEEN: {                                                                   // engine.rs:99-111
        Res ch = ex.ct_pos(th.a);
        Res f = ex.ct_constant((uint8_t)(th.b & 255)), to = ex.ct_constant((uint8_t)(th.b >> 8));
        Res ge = ex.ct_ge(ch, f);
        Res le = ex.ct_le(ch, to);
        r = ex.ct_and(ge, le);
        break;
      }
      case Thunk::RANGE: {                                                                     // engine.rs:112-126
        const auto& cs = B.ranges[th.b];
        Res ch = ex.ct_pos(th.a);
        r = ex.ct_eq(ch, ex.ct_constant(cs[0]));
        for (size_t i = 1; i < cs.size(); i++) {
          Res e = ex.ct_eq(ch, ex.ct_constant(cs[i]));
          r = ex.ct_or(r, e);
        }
        break;
      }
      case Thunk::ANDTHEN: {                                                                   // engine.rs:165-177,196-208
        Res p = eval(th.a);
        Res x = eval(th.b);
        r = ex.ct_and(p, x);
        break;
      }
    }
    Memo& m2 = memo[t];
    m2.res = r;
    m2.calls = ex.calls - before;
    m2.done = true;
    return r;
  }
};

// =================================================================================================
// lowering: boolean value DAG -> PBS nodes (sum of <= 15 boolean LWEs, then one LUT)
// =================================================================================================
struct Lit { int32_t node; bool neg; };  // node: PBS node id (>= 0) ; constants handled separately
struct LitOrConst { int kind; Lit lit; };  // kind 0: const false, 1: const true, 2: literal

struct PbsNode {
  uint32_t lut;
  std::vector<LinTerm> terms;   // rows are *node refs*: >= 0 PBS node id, < 0: pack ref -(1 + 2*pos + half)
  uint64_t add_const;           // units of delta (2^59)
  int level;
  int32_t row;                  // assigned later
};

struct Lowering {
  Execution& ex;
  size_t n;
  std::vector<PbsNode> nodes;
  std::map<std::vector<int64_t>, int32_t> node_memo;      // canonical (lut, const, terms) -> node
  std::vector<LitOrConst> lowered;                          // value id -> literal (kind -1: not lowered yet)
  std::vector<uint32_t> seen_stamp;                         // flattening scratch: value id -> generation
  uint32_t seen_gen = 0;
  bool absorb = true;                                       // drop OR operands implied by another operand (x | (x & y) = x)
  // multi-GPU with absorption: every rank lowers the WHOLE match, absorbs globally, and keeps a contiguous slice
  // (by leftmost content position) of the surviving operands of the final OR -- so a rank neither evaluates variants
  // that a variant of another rank implies nor bootstraps leaves outside its slice
  int shard_rank = 0, shard_world = 1;
  int32_t shard_root = -1;
  bool sharded = false;
  std::unordered_map<int32_t, std::vector<Lit>> and_sets;   // PBS node of a lowered AND -> its full literal set
  uint64_t absorbed = 0;
  std::vector<std::pair<int32_t, int32_t>> eq_parts;         // eq node -> its two nibble-test nodes (eq = their AND), or (-1,-1)
  std::map<std::vector<int64_t>, int32_t> shape_intern;    // shift-invariant structure -> shape id
  std::unordered_map<int32_t, std::pair<int32_t, int32_t>> node_shape;  // PBS node -> (shape, base position)
  std::map<int32_t, std::map<int32_t, int32_t>> shape_nodes;           // shape -> base position -> PBS node
  std::map<std::vector<int64_t>, Lit> run_memo;             // (shape, neg, start, len) -> literal of the AND over the run

  Lowering(Execution& e, size_t n_chars)
      : ex(e), n(n_chars), lowered(e.vals.items.size(), LitOrConst{-1, {0, false}}), seen_stamp(e.vals.items.size(), 0) {}
  bool is_lowered(int32_t v) const { return lowered[v].kind >= 0; }

  int level_of(int32_t ref) const { return ref < 0 ? 0 : nodes[ref].level; }

  int32_t make_node(uint32_t lut, std::vector<LinTerm> terms, uint64_t add_const) {
    std::sort(terms.begin(), terms.end(), [](const LinTerm& x, const LinTerm& y) { return x.row < y.row || (x.row == y.row && x.coef < y.coef); });
    std::vector<int64_t> key;
    key.reserve(2 + 2 * terms.size());
    key.push_back(lut);
    key.push_back((int64_t)add_const);
    for (auto& t : terms) { key.push_back(t.row); key.push_back(t.coef); }
    auto it = node_memo.find(key);
    if (it != node_memo.end()) return it->second;
    int lvl = 0;
    for (auto& t : terms) lvl = std::max(lvl, level_of(t.row));
    PbsNode nd{lut, std::move(terms), add_const, lvl + 1, -1};
    nodes.push_back(std::move(nd));
    int32_t id = (int32_t)nodes.size() - 1;
    node_memo.emplace(std::move(key), id);
    return id;
  }
  static int32_t pack_ref(int32_t pos, int half) { return -(1 + 2 * pos + half); }
  int32_t nibble(int32_t pos, int half, uint32_t lut_base, uint32_t v) {
    return make_node(lut_base + v, {LinTerm{pack_ref(pos, half), 1}}, 0);
  }

  // Two materialised booleans with the same shape id are position shifts of one another
  // (e.g. eq(content[i], 'a') for different i); runs of consecutive shifts inside an AND are what
  // the overlap-doubling below compresses.
  void register_shape(int32_t node, std::vector<int64_t> key, int32_t base) {
    if (node_shape.count(node)) return;
    auto it = shape_intern.find(key);
    int32_t sid;
    if (it == shape_intern.end()) { sid = (int32_t)shape_intern.size(); shape_intern.emplace(std::move(key), sid); }
    else sid = it->second;
    node_shape[node] = {sid, base};
    shape_nodes[sid].emplace(base, node);
  }
  void register_set_shape(int32_t node, const std::vector<Lit>& lits, bool is_and) {
    if (lits.size() > 32 || node_shape.count(node)) return;
    int32_t base = INT32_MAX;
    for (auto& l : lits) {
      auto it = node_shape.find(l.node);
      if (it == node_shape.end()) return;
      base = std::min(base, it->second.second);
    }
    std::vector<std::vector<int64_t>> parts;
    for (auto& l : lits) {
      auto sb = node_shape[l.node];
      parts.push_back({sb.first, l.neg ? 1 : 0, sb.second - base});
    }
    std::sort(parts.begin(), parts.end());
    std::vector<int64_t> key{is_and ? -1 : -2};
    for (auto& p : parts) key.insert(key.end(), p.begin(), p.end());
    register_shape(node, std::move(key), base);
  }

  // sum-then-LUT over a literal list (<= 15 literals): AND -> sum == k, OR -> sum >= 1
  Lit combine_small(const std::vector<Lit>& lits, bool is_and) {
    std::vector<LinTerm> terms;
    uint64_t add = 0;
    for (auto& l : lits) {
      if (l.neg) { terms.push_back(LinTerm{l.node, -1}); add += 1; }
      else terms.push_back(LinTerm{l.node, 1});
    }
    const uint32_t lut = is_and ? LUT_SUM_EQ + (uint32_t)lits.size() : (uint32_t)LUT_GE1;
    return Lit{make_node(lut, std::move(terms), add), false};
  }
  Lit combine(std::vector<Lit> lits, bool is_and) {
    while (lits.size() > 15) {  // balanced fan-in-15 tree
      std::vector<Lit> nxt;
      const size_t groups = (lits.size() + 14) / 15;
      const size_t per = (lits.size() + groups - 1) / groups;
      for (size_t g = 0; g < lits.size(); g += per) {
        std::vector<Lit> grp(lits.begin() + g, lits.begin() + std::min(lits.size(), g + per));
        nxt.push_back(grp.size() == 1 ? grp[0] : combine_small(grp, is_and));
      }
      lits.swap(nxt);
    }
    if (lits.size() == 1) return lits[0];
    return combine_small(lits, is_and);
  }

  // AND over a run of `len` consecutive shifts of one literal shape, shared across all variants.
  // len <= 15: one sum-then-LUT node over the literals themselves; len = 15 * 2^k: overlap doubling of
  // two half runs (AND is idempotent).  Fan-in 15 at the bottom keeps the dependent depth of a run of L
  // positions at 1 + ceil(log2(L / 75)) levels instead of log2(L).
  Lit run_block(int32_t shape, bool neg, int32_t start, int32_t len, const std::map<int32_t, int32_t>& by_base) {
    std::vector<int64_t> key{shape, neg ? 1 : 0, start, len};
    auto it = run_memo.find(key);
    if (it != run_memo.end()) return it->second;
    Lit r;
    if (len == 1) {
      r = Lit{by_base.at(start), neg};
    } else if (len <= 15) {
      std::vector<Lit> lits;
      for (int32_t t = 0; t < len; t++) lits.push_back(Lit{by_base.at(start + t), neg});
      r = combine_small(lits, true);
    } else {
      Lit a = run_block(shape, neg, start, len / 2, by_base);
      Lit b = run_block(shape, neg, start + len / 2, len / 2, by_base);
      r = combine_small({a, b}, true);
    }
    run_memo.emplace(std::move(key), r);
    return r;
  }
  // cover [start, start+len) with at most 5 equal blocks (the last one overlapping its neighbour)
  void cover_run(int32_t shape, bool neg, int32_t start, int32_t len, const std::map<int32_t, int32_t>& by_base, std::vector<Lit>& out) {
    if (len <= 15) { out.push_back(run_block(shape, neg, start, len, by_base)); return; }
    int32_t B = 15;
    while ((len + B - 1) / B > 5) B *= 2;
    const int32_t m = (len + B - 1) / B;
    for (int32_t t = 0; t + 1 < m; t++) out.push_back(run_block(shape, neg, start + t * B, B, by_base));
    out.push_back(run_block(shape, neg, start + len - B, B, by_base));
  }

  // Absorption over the operands of an OR: an operand that is the AND of a superset of another operand's
  // literals is implied by it (x | (x & y) = x) and is dropped; the boolean function, hence the decrypted
  // result, is unchanged.  This is what turns the reference's O(n^2) variants of an unanchored /a+.../ into
  // O(n) bootstraps: the variant that starts at i with run length L contains the literals of the variant
  // that starts at i+L-1 with run length 1.  Sets are sorted literal lists; candidates are visited by
  // increasing size and looked up through the kept sets indexed by their rarest literal.
  static uint64_t lit_key(const Lit& l) { return ((uint64_t)(uint32_t)l.node << 1) | (l.neg ? 1u : 0u); }
  double absorb_ms = 0;
  void absorb_or_operands(std::vector<Lit>& ops) {
    const auto t_begin = std::chrono::steady_clock::now();
    struct Guard { double& acc; std::chrono::steady_clock::time_point t; ~Guard() { acc += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); } } guard{absorb_ms, t_begin};
    struct Item { std::vector<uint64_t> set; Lit lit; };
    std::vector<Item> items;
    items.reserve(ops.size());
    for (auto& l : ops) {
      Item it{{}, l};
      auto f = l.neg ? and_sets.end() : and_sets.find(l.node);
      if (f != and_sets.end()) for (auto& x : f->second) it.set.push_back(lit_key(x));
      else it.set.push_back(lit_key(l));
      std::sort(it.set.begin(), it.set.end());
      items.push_back(std::move(it));
    }
    std::unordered_map<uint64_t, uint32_t> freq;                  // literal -> number of operand sets holding it
    for (auto& it : items)
      for (uint64_t x : it.set) freq[x]++;
    std::vector<size_t> order(items.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return items[a].set.size() < items[b].set.size(); });
    std::unordered_map<uint64_t, std::vector<size_t>> by_rare;    // rarest literal of a kept set -> kept items
    std::vector<Lit> kept;
    uint64_t budget = 2000000000ull;                               // literal comparisons; beyond it the rest is kept as is
    for (size_t oi : order) {
      const auto& B = items[oi].set;
      bool implied = false;
      for (size_t xi2 = 0; xi2 < B.size() && !implied && budget > 0; xi2++) {
        auto f = by_rare.find(B[xi2]);
        if (f == by_rare.end()) continue;
        for (size_t ai : f->second) {
          const auto& A = items[ai].set;
          if (A.size() > B.size()) continue;
          budget -= std::min<uint64_t>(budget, A.size() + B.size());
          if (std::includes(B.begin(), B.end(), A.begin(), A.end())) { implied = true; break; }
        }
      }
      if (implied) { absorbed++; continue; }
      uint64_t rare = B[0];
      for (uint64_t x : B)
        if (freq[x] < freq[rare]) rare = x;
      by_rare[rare].push_back(oi);
      kept.push_back(items[oi].lit);
    }
    std::sort(kept.begin(), kept.end(), [](const Lit& x, const Lit& y) { return x.node < y.node || (x.node == y.node && x.neg < y.neg); });
    ops.swap(kept);
  }

  // Absorption on the VALUE level, before any operand of a wide OR is lowered: an operand that is an AND over a
  // superset of another operand's conjuncts is implied by it and is neither lowered nor kept (for the 256-character
  // /a+b?c/ this spares lowering 64 771 of the 65 025 variants; sound for the same reason as absorb_or_operands).
  // Sets are sorted value ids; candidates go by increasing size against kept sets indexed by their rarest member.
  std::unordered_map<int32_t, std::vector<int32_t>> pre_absorbed;   // OR value -> surviving operand values
  const std::vector<int32_t>& pre_absorb(int32_t v, const std::vector<int32_t>& ops) {
    auto cached = pre_absorbed.find(v);
    if (cached != pre_absorbed.end()) return cached->second;
    const auto t_begin = std::chrono::steady_clock::now();
    const size_t nv = ex.vals.items.size();
    // sorted conjunct set of every AND value under the operands, bottom-up and memoised: the operands of a lazily
    // enumerated match share their prefixes (hash-consed values), so a set is one merge of its children's sets
    std::vector<std::vector<int32_t>> memo(nv);
    std::vector<char> state(nv, 0);                  // 0 = not computed, 1 = set ready, 2 = contains a false conjunct
    auto conj = [&](int32_t root) {
      std::vector<int32_t> stack{root};
      while (!stack.empty()) {
        const int32_t x = stack.back();
        if (state[x]) { stack.pop_back(); continue; }
        const Triple tx = ex.vals.items[x];
        if (tx.t == V_CONST) { state[x] = (tx.a & 1) ? 1 : 2; stack.pop_back(); continue; }   // true: empty set (neutral)
        if (tx.t != V_AND) { memo[x].assign(1, x); state[x] = 1; stack.pop_back(); continue; }
        if (!state[tx.a]) { stack.push_back(tx.a); continue; }
        if (!state[tx.b]) { stack.push_back(tx.b); continue; }
        if (state[tx.a] == 2 || state[tx.b] == 2) state[x] = 2;
        else {
          const auto &A = memo[tx.a], &B = memo[tx.b];
          memo[x].resize(A.size() + B.size());
          memo[x].erase(std::set_union(A.begin(), A.end(), B.begin(), B.end(), memo[x].begin()), memo[x].end());
          state[x] = 1;
        }
        stack.pop_back();
      }
    };
    std::vector<const std::vector<int32_t>*> setp(ops.size());
    std::vector<char> dead(ops.size(), 0);
    std::vector<uint32_t> freq(nv, 0);
    for (size_t o = 0; o < ops.size(); o++) {
      conj(ops[o]);
      dead[o] = state[ops[o]] == 2;                  // a false conjunct: the operand is false, neutral in the OR
      setp[o] = &memo[ops[o]];
      if (!dead[o]) for (int32_t x : *setp[o]) freq[x]++;
    }
    auto sets = [&](size_t o) -> const std::vector<int32_t>& { return *setp[o]; };
    std::vector<uint32_t> order;
    for (size_t o = 0; o < ops.size(); o++) if (!dead[o]) order.push_back((uint32_t)o);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return sets(a).size() < sets(b).size(); });
    std::vector<std::vector<uint32_t>> by_rare(nv);
    std::vector<char> keep(ops.size(), 0);
    for (uint32_t o : order) {
      const auto& B = sets(o);
      bool implied = false;
      for (size_t bi = 0; bi < B.size() && !implied; bi++)
        for (uint32_t a : by_rare[B[bi]]) {
          const auto& A = sets(a);
          if (A.size() <= B.size() && std::includes(B.begin(), B.end(), A.begin(), A.end())) { implied = true; break; }
        }
      if (implied) { absorbed++; continue; }
      keep[o] = 1;
      if (B.empty()) continue;               // an empty conjunction is the constant true: the lowering decides the OR
      int32_t rare = B[0];
      for (int32_t x : B) if (freq[x] < freq[rare]) rare = x;
      by_rare[rare].push_back(o);
    }
    std::vector<int32_t> out;
    for (size_t o = 0; o < ops.size(); o++) if (keep[o]) out.push_back(ops[o]);
    absorb_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    return pre_absorbed.emplace(v, std::move(out)).first->second;
  }

  // this rank's contiguous slice of the final OR's operands, ordered by the leftmost content position they read
  void shard_operands(std::vector<Lit>& ops) {
    std::vector<std::pair<int64_t, Lit>> keyed;
    keyed.reserve(ops.size());
    for (auto& l : ops) {
      auto it = node_shape.find(l.node);
      keyed.emplace_back(it == node_shape.end() ? (int64_t)INT32_MAX : (int64_t)it->second.second, l);
    }
    std::stable_sort(keyed.begin(), keyed.end(), [](const std::pair<int64_t, Lit>& a, const std::pair<int64_t, Lit>& b) { return a.first < b.first; });
    const size_t n = keyed.size();
    const size_t lo = n * (size_t)shard_rank / (size_t)shard_world, hi = n * ((size_t)shard_rank + 1) / (size_t)shard_world;
    std::vector<Lit> mine;
    for (size_t i = lo; i < hi; i++) mine.push_back(keyed[i].second);
    std::sort(mine.begin(), mine.end(), [](const Lit& x, const Lit& y) { return x.node < y.node || (x.node == y.node && x.neg < y.neg); });
    ops.swap(mine);
  }

  LitOrConst lower(int32_t root) {
    if (shard_world > 1) shard_root = root;
    // iterative post-order over the value DAG (the sequential OR fold is tens of thousands deep)
    std::vector<int32_t> stack{root};
    while (!stack.empty()) {
      const int32_t v = stack.back();
      if (is_lowered(v)) { stack.pop_back(); continue; }
      const Triple t = ex.vals.items[v];
      if (t.t == V_CONST) { lowered[v] = LitOrConst{t.a & 1, {0, false}}; stack.pop_back(); continue; }
      if (t.t == V_EQ) {  // eq = [e_lo + e_hi == 2]
        Lit e0{nibble(t.a, 0, LUT_NIB_EQ, t.b & 15), false}, e1{nibble(t.a, 1, LUT_NIB_EQ, (t.b >> 4) & 15), false};
        Lit r = combine_small({e0, e1}, true);
        lowered[v] = LitOrConst{2, r};
        register_shape(r.node, {V_EQ, t.b}, t.a);
        register_shape(e0.node, {-10, t.b & 15}, t.a);          // nibble tests are shiftable shapes too
        register_shape(e1.node, {-11, (t.b >> 4) & 15}, t.a);
        if ((size_t)r.node >= eq_parts.size()) eq_parts.resize(std::max<size_t>((size_t)r.node + 1, eq_parts.size() * 2), {-1, -1});
        eq_parts[r.node] = {e0.node, e1.node};
        stack.pop_back();
        continue;
      }
      if (t.t == V_GT || t.t == V_LE) {  // a > c  <=>  2*[hi > c_hi] + [hi == c_hi] + [lo > c_lo] >= 2
        int32_t g1 = nibble(t.a, 1, LUT_NIB_GT, (t.b >> 4) & 15), e1 = nibble(t.a, 1, LUT_NIB_EQ, (t.b >> 4) & 15);
        int32_t g0 = nibble(t.a, 0, LUT_NIB_GT, t.b & 15);
        int32_t nd = make_node(t.t == V_GT ? LUT_GE2 : LUT_LT2, {LinTerm{g1, 2}, LinTerm{e1, 1}, LinTerm{g0, 1}}, 0);
        lowered[v] = LitOrConst{2, Lit{nd, false}};
        register_shape(nd, {t.t, t.b}, t.a);
        stack.pop_back();
        continue;
      }
      if (t.t == V_NOT) {
        if (!is_lowered(t.a)) { stack.push_back(t.a); continue; }
        LitOrConst a = lowered[t.a];
        if (a.kind < 2) lowered[v] = LitOrConst{a.kind ^ 1, {0, false}};
        else lowered[v] = LitOrConst{2, Lit{a.lit.node, !a.lit.neg}};
        stack.pop_back();
        continue;
      }
      // V_AND / V_OR: flatten nested nodes of the same operator into an operand set
      std::vector<int32_t> ops, work{t.a, t.b};
      bool missing = false;
      {
        seen_gen++;
        while (!work.empty()) {
          int32_t x = work.back();
          work.pop_back();
          if (seen_stamp[x] == seen_gen) continue;
          seen_stamp[x] = seen_gen;
          const Triple tx = ex.vals.items[x];
          if (tx.t == t.t) { work.push_back(tx.a); work.push_back(tx.b); continue; }
          ops.push_back(x);
        }
      }
      const bool is_and = t.t == V_AND;
      if (!is_and && absorb && ops.size() > 16) ops = pre_absorb(v, ops);
      for (int32_t x : ops)
        if (!is_lowered(x)) { stack.push_back(x); missing = true; }
      if (missing) continue;
      std::vector<Lit> lits;
      bool decided = false;
      // operand literals: drop neutral constants, dedupe (idempotence), detect x & !x / x | !x
      for (int32_t x : ops) {
        const LitOrConst& lx = lowered[x];
        if (lx.kind < 2) {
          if ((lx.kind == 0) == is_and) { decided = true; break; }  // absorbing element
          continue;                                                 // neutral element
        }
        lits.push_back(lx.lit);
      }
      if (decided) {
        lowered[v] = LitOrConst{is_and ? 0 : 1, {0, false}};
      } else {
        std::sort(lits.begin(), lits.end(), [](const Lit& x, const Lit& y) { return x.node < y.node || (x.node == y.node && x.neg < y.neg); });
        size_t o = 0;
        for (size_t i2 = 0; i2 < lits.size(); i2++) {
          if (o > 0 && lits[o - 1].node == lits[i2].node) {
            if (lits[o - 1].neg != lits[i2].neg) { decided = true; break; }  // both polarities of one node
            continue;                                                       // duplicate
          }
          lits[o++] = lits[i2];
        }
        if (decided) lowered[v] = LitOrConst{is_and ? 0 : 1, {0, false}};
        else lits.resize(o);
      }
      if (decided) { stack.pop_back(); continue; }
      if (is_and && lits.size() > 1) {
        // a plain eq under an AND is itself the AND of two nibble tests: use those directly, which takes the
        // eq-combine level off the dependent path (the eq node stays for consumers that need it whole)
        std::vector<Lit> flat;
        bool changed = false;
        for (auto& l : lits) {
          if (l.neg || (size_t)l.node >= eq_parts.size() || eq_parts[l.node].first < 0) { flat.push_back(l); continue; }
          flat.push_back(Lit{eq_parts[l.node].first, false});
          flat.push_back(Lit{eq_parts[l.node].second, false});
          changed = true;
        }
        if (changed) {
          std::sort(flat.begin(), flat.end(), [](const Lit& x, const Lit& y) { return x.node < y.node || (x.node == y.node && x.neg < y.neg); });
          flat.erase(std::unique(flat.begin(), flat.end(), [](const Lit& x, const Lit& y) { return x.node == y.node && x.neg == y.neg; }), flat.end());
          lits.swap(flat);
        }
      }
      if (lits.empty()) { lowered[v] = LitOrConst{is_and ? 1 : 0, {0, false}}; stack.pop_back(); continue; }
      if (!is_and && absorb && lits.size() > 1) absorb_or_operands(lits);
      if (!is_and && v == shard_root && shard_world > 1) {
        shard_operands(lits);
        sharded = true;
        if (lits.empty()) { lowered[v] = LitOrConst{0, {0, false}}; stack.pop_back(); continue; }
      }
      const std::vector<Lit> full = lits;
      // one sum-then-LUT node takes up to 15 literals: shared run blocks are only worth a level beyond that
      if (is_and && lits.size() > 15) lits = compress_runs(lits, 3);
      const Lit r = combine(lits, is_and);
      lowered[v] = LitOrConst{2, r};
      if (is_and && absorb && !r.neg && full.size() > 1) and_sets.emplace(r.node, full);
      if (full.size() > 1 && !r.neg) register_set_shape(r.node, full, is_and);
      stack.pop_back();
    }
    return lowered[root];
  }

  // replace runs (>= 3 consecutive base positions of one literal shape) by <= 5 shared run-block nodes
  std::vector<Lit> compress_runs(const std::vector<Lit>& lits, int32_t min_run) {
    std::map<std::pair<int32_t, bool>, std::vector<std::pair<int32_t, Lit>>> groups;  // (shape,neg) -> (base, lit)
    std::vector<Lit> out;
    for (auto& l : lits) {
      auto it = node_shape.find(l.node);
      if (it == node_shape.end()) { out.push_back(l); continue; }
      groups[{it->second.first, l.neg}].emplace_back(it->second.second, l);
    }
    for (auto& g : groups) {
      auto& v = g.second;
      std::sort(v.begin(), v.end(), [](const std::pair<int32_t, Lit>& a, const std::pair<int32_t, Lit>& b) { return a.first < b.first; });
      const auto& by_base = shape_nodes[g.first.first];
      size_t i = 0;
      while (i < v.size()) {
        size_t j = i;
        while (j + 1 < v.size() && v[j + 1].first == v[j].first + 1) j++;
        const int32_t len = (int32_t)(j - i + 1), start = v[i].first;
        bool ok = len >= min_run;   // short runs are cheaper as plain literals (a block node costs a level)
        // every position of the run must map to the very node we hold
        for (size_t t = i; ok && t <= j; t++) {
          auto gi = by_base.find(v[t].first);
          ok = gi != by_base.end() && gi->second == v[t].second.node;
        }
        if (!ok) {
          for (size_t t = i; t <= j; t++) out.push_back(v[t].second);
        } else {
          cover_run(g.first.first, g.first.second, start, len, by_base, out);
        }
        i = j + 1;
      }
    }
    return out;
  }
};

void emit_plan(Lowering& L, const LitOrConst& result, size_t n_chars, Plan& plan) {
  // keep only nodes reachable from the result
  std::vector<char> live(L.nodes.size(), 0);
  if (result.kind == 2) {
    std::vector<int32_t> st{result.lit.node};
    while (!st.empty()) {
      int32_t x = st.back();
      st.pop_back();
      if (live[x]) continue;
      live[x] = 1;
      for (auto& t : L.nodes[x].terms)
        if (t.row >= 0) st.push_back(t.row);
    }
  }
  int max_level = 0;
  for (size_t i = 0; i < L.nodes.size(); i++)
    if (live[i]) max_level = std::max(max_level, L.nodes[i].level);
  std::vector<std::vector<int32_t>> by_level(max_level + 1);
  std::vector<char> pack_used(2 * n_chars, 0);
  for (size_t i = 0; i < L.nodes.size(); i++) {
    if (!live[i]) continue;
    by_level[L.nodes[i].level].push_back((int32_t)i);
    for (auto& t : L.nodes[i].terms)
      if (t.row < 0) pack_used[-t.row - 1] = 1;
  }
  plan.n_chars = n_chars;
  plan.levels.clear();
  const int32_t content_rows = (int32_t)(4 * n_chars);
  const int32_t pack_base = content_rows;
  int32_t next_row = pack_base + (int32_t)(2 * n_chars);
  // level 0: nibble packing p = b_even + 4 * b_odd (engine keeps characters as 4 two-bit blocks, ciphertext.rs:8-30)
  PlanLevel l0;
  l0.lin_term_off.push_back(0);
  for (size_t p = 0; p < 2 * n_chars; p++) {
    if (!pack_used[p]) continue;
    const int32_t pos = (int32_t)(p / 2), half = (int32_t)(p % 2);
    l0.lin_out_rows.push_back(pack_base + (int32_t)p);
    l0.lin_term_rows.push_back(4 * pos + 2 * half);
    l0.lin_coef.push_back(1);
    l0.lin_term_rows.push_back(4 * pos + 2 * half + 1);
    l0.lin_coef.push_back(4);
    l0.lin_term_off.push_back((int32_t)l0.lin_term_rows.size());
    l0.lin_const.push_back(0);
  }
  plan.levels.push_back(std::move(l0));
  size_t max_width = 0, total = 0;
  for (int lv = 1; lv <= max_level; lv++) {
    max_width = std::max(max_width, by_level[lv].size());
    total += by_level[lv].size();
    for (int32_t id : by_level[lv]) L.nodes[id].row = next_row++;
  }
  const int32_t scratch_base = next_row;
  next_row += (int32_t)max_width + 1;
  auto row_of = [&](int32_t ref) -> int32_t { return ref < 0 ? pack_base + (-ref - 1) : L.nodes[ref].row; };
  for (int lv = 1; lv <= max_level; lv++) {
    PlanLevel pl;
    pl.lin_term_off.push_back(0);
    pl.out_row_base = by_level[lv].empty() ? 0 : L.nodes[by_level[lv][0]].row;
    int32_t scratch = scratch_base;
    for (int32_t id : by_level[lv]) {
      const PbsNode& nd = L.nodes[id];
      pl.lut_idx.push_back(nd.lut);
      if (nd.terms.size() == 1 && nd.terms[0].coef == 1 && nd.add_const == 0) {
        pl.in_rows.push_back(row_of(nd.terms[0].row));
        continue;
      }
      pl.lin_out_rows.push_back(scratch);
      for (auto& t : nd.terms) { pl.lin_term_rows.push_back(row_of(t.row)); pl.lin_coef.push_back(t.coef); }
      pl.lin_term_off.push_back((int32_t)pl.lin_term_rows.size());
      pl.lin_const.push_back(nd.add_const << 59);
      pl.in_rows.push_back(scratch++);
    }
    plan.levels.push_back(std::move(pl));
  }
  plan.result_kind = result.kind;
  plan.result_row = -1;
  if (result.kind == 2) {
    if (!result.lit.neg) {
      plan.result_row = L.nodes[result.lit.node].row;
    } else {  // 1 - x into the spare scratch row
      PlanLevel pl;
      pl.lin_term_off.push_back(0);
      pl.lin_out_rows.push_back(scratch_base + (int32_t)max_width);
      pl.lin_term_rows.push_back(L.nodes[result.lit.node].row);
      pl.lin_coef.push_back(-1);
      pl.lin_term_off.push_back(1);
      pl.lin_const.push_back(1ull << 59);
      plan.levels.push_back(std::move(pl));
      plan.result_row = scratch_base + (int32_t)max_width;
    }
  }
  plan.n_rows = next_row;
  plan.stats.pbs = total;
  plan.stats.levels = (uint64_t)max_level;
  plan.stats.max_level_width = max_width;
}

}  // namespace

uint64_t lut_value(uint32_t lut_id, uint32_t x) {
  if (lut_id < LUT_NIB_GT) return x == lut_id - LUT_NIB_EQ;
  if (lut_id < LUT_SUM_EQ) return x > lut_id - LUT_NIB_GT;
  if (lut_id < LUT_GE1) return x == lut_id - LUT_SUM_EQ;
  if (lut_id == LUT_GE1) return x >= 1;
  if (lut_id == LUT_GE2) return x >= 2;
  return x < 2;
}

int build_plan(const std::string& pattern, size_t n_chars, int rank, int world, const PlanOptions& opt, Plan& plan, std::string& err) {
  RegExpr re;
  int rc = parse(pattern, re, err);
  if (rc != FB_OK) return rc;
  if (world < 1 || rank < 0 || rank >= world) { err = "bad rank/world"; return FB_ERR_ARG; }
  try {
    const bool timing = opt.timing;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    Builder B(n_chars);
    std::vector<int32_t> branches;
    const bool absorb = opt.absorb;   // false: reference-shaped plan (every variant evaluated)
    const bool shard_at_root = absorb && world > 1;                    // see Lowering::shard_operands
    for (size_t i = 0; i < n_chars; i++) {  // engine.rs:15-18
      if (!shard_at_root && (int)(i % (size_t)world) != rank) continue;
      for (auto& b : B.build(re, i)) branches.push_back(b.first);
    }
    const double t1 = now();
    Execution ex;
    ThunkEval ev(B, ex);
    Res res;
    if (branches.size() <= 1) {  // engine.rs:22-26
      res = branches.empty() ? ex.ct_false() : ev.eval(branches[0]);
    } else {                     // engine.rs:28-34
      res = ev.eval(branches[0]);
      for (size_t i = 1; i < branches.size(); i++) {
        Res br = ev.eval(branches[i]);
        res = ex.ct_or(res, br);
      }
    }
    const double t2 = now();
    Lowering L(ex, n_chars);
    L.absorb = absorb;
    if (shard_at_root) { L.shard_rank = rank; L.shard_world = world; }
    LitOrConst out = L.lower(res.val);
    // no final OR to slice (a single variant, a constant): rank 0 owns the whole result
    if (shard_at_root && !L.sharded && rank != 0) out = LitOrConst{0, {0, false}};
    const double t3 = now();
    plan = Plan();
    emit_plan(L, out, n_chars, plan);
    if (timing)
      fprintf(stderr, "[plan] enumerate %.1f ms, evaluate (executor bookkeeping) %.1f ms, lower %.1f ms (absorption %.1f ms), emit %.1f ms\n",
              t1 - t0, t2 - t1, t3 - t2, L.absorb_ms, now() - t3);
    plan.stats.variants = branches.size();
    plan.stats.ct_ops = ex.ct_ops;
    plan.stats.cache_hits = ex.cache_hits;
    plan.stats.ops_eq = ex.by_type[K_EQ];
    plan.stats.ops_gt = ex.by_type[K_GE];
    plan.stats.ops_le = ex.by_type[K_LE];
    plan.stats.ops_and = ex.by_type[K_AND];
    plan.stats.ops_or = ex.by_type[K_OR];
    plan.stats.ops_not = ex.by_type[K_NOT];
    return FB_OK;
  } catch (const PanicEx& e) {
    err = e.msg;
    return FB_ERR_PANIC;
  }
}

void build_or_fold_plan(size_t n, Plan& plan) {
  // rows 0..n-1 hold the booleans; reuse the lowering's fan-in-15 tree with "content rows" = inputs
  plan = Plan();
  plan.n_chars = 0;
  std::vector<int32_t> cur;
  for (size_t i = 0; i < n; i++) cur.push_back((int32_t)i);
  int32_t next_row = (int32_t)n;
  plan.levels.push_back(PlanLevel());  // level 0: nothing to pack
  plan.levels[0].lin_term_off.push_back(0);
  size_t total = 0, max_width = 0;
  std::vector<std::vector<std::vector<int32_t>>> level_groups;
  while (cur.size() > 1) {
    const size_t groups = (cur.size() + 14) / 15;
    const size_t per = (cur.size() + groups - 1) / groups;
    std::vector<std::vector<int32_t>> grp;
    std::vector<int32_t> nxt;
    for (size_t g = 0; g < cur.size(); g += per) {
      std::vector<int32_t> one(cur.begin() + g, cur.begin() + std::min(cur.size(), g + per));
      if (one.size() == 1) { nxt.push_back(one[0]); continue; }
      grp.push_back(one);
      nxt.push_back(next_row++);
    }
    level_groups.push_back(grp);
    max_width = std::max(max_width, grp.size());
    total += grp.size();
    cur.swap(nxt);
  }
  const int32_t scratch_base = next_row;
  next_row += (int32_t)max_width + 1;
  int32_t out_row = (int32_t)n;
  for (auto& grp : level_groups) {
    PlanLevel pl;
    pl.lin_term_off.push_back(0);
    pl.out_row_base = out_row;
    int32_t scratch = scratch_base;
    for (auto& one : grp) {
      pl.lut_idx.push_back(LUT_GE1);
      pl.lin_out_rows.push_back(scratch);
      for (int32_t r : one) { pl.lin_term_rows.push_back(r); pl.lin_coef.push_back(1); }
      pl.lin_term_off.push_back((int32_t)pl.lin_term_rows.size());
      pl.lin_const.push_back(0);
      pl.in_rows.push_back(scratch++);
      out_row++;
    }
    plan.levels.push_back(std::move(pl));
  }
  plan.result_kind = n == 0 ? 0 : 2;
  plan.result_row = n == 0 ? -1 : cur[0];
  plan.n_rows = next_row;
  plan.stats.pbs = total;
  plan.stats.levels = level_groups.size();
  plan.stats.max_level_width = max_width;
}

// Plaintext dry run of a plan: rows hold messages instead of ciphertexts, a PBS is its LUT.  Checks
// the lowering (and that no PBS input leaves the 4-bit message+carry space) without any GPU.
int eval_plan_plain(const Plan& plan, const uint8_t* content, size_t n_in_rows_are_blocks, int* result, std::string& err) {
  if (plan.result_kind < 2) { *result = plan.result_kind; return FB_OK; }
  std::vector<int64_t> rows((size_t)plan.n_rows, 0);
  if (plan.n_chars) {
    for (size_t i = 0; i < plan.n_chars; i++)
      for (int b = 0; b < 4; b++) rows[4 * i + b] = (content[i] >> (2 * b)) & 3;
  } else {
    for (size_t i = 0; i < n_in_rows_are_blocks; i++) rows[i] = content[i];
  }
  for (auto& l : plan.levels) {
    for (size_t o = 0; o < l.lin_out_rows.size(); o++) {
      int64_t v = (int64_t)(l.lin_const[o] >> 59);
      for (int32_t t = l.lin_term_off[o]; t < l.lin_term_off[o + 1]; t++) v += l.lin_coef[t] * rows[l.lin_term_rows[t]];
      rows[l.lin_out_rows[o]] = v;
    }
    std::vector<int64_t> outs(l.in_rows.size());
    for (size_t b = 0; b < l.in_rows.size(); b++) {
      const int64_t x = rows[l.in_rows[b]];
      if (x < 0 || x > 15) { err = "PBS input outside the 4-bit message space"; return FB_ERR_ARG; }
      outs[b] = (int64_t)lut_value(l.lut_idx[b], (uint32_t)x);
    }
    for (size_t b = 0; b < outs.size(); b++) rows[l.out_row_base + b] = outs[b];
  }
  const int64_t r = rows[plan.result_row];
  if (r != 0 && r != 1) { err = "result is not a boolean"; return FB_ERR_ARG; }
  *result = (int)r;
  return FB_OK;
}

}  // namespace fbre
