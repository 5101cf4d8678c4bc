// Constraint programs: a table's `Stark::eval_packed_generic` (+ its lookup checks) exported ONCE as
// straight-line SSA code over the base field, so that the quotient kernel of ANY starky table — the
// arithmetic, byte-packing, CPU, keccak, keccak-sponge, logic and memory STARKs of evm_arithmetization 0.1.3
// (/root/reference/Cargo.lock:1675; reached from /root/reference/ops/src/lib.rs:52) — can be compiled for
// sm_100a without this library knowing the table.  The patched starky records the program by running the
// table's evaluator on a symbolic PackedField (rust/etp_b200_sys); tests build programs with
// eth_tx_proof_b200/cprog.py.
//
// Wire format (u64 words, little endian):
//   [0] magic "ETPCPRG1"  [1] n_ops  [2] n_trace_cols  [3] n_aux_cols  [4] n_public_inputs
//   [5] n_challenges (lookup challenge scalars)  [6] constraint_degree  [7] n_constraints
//   then n_ops x { w0 = opcode | a << 8 | b << 36 ; w1 = immediate }.  The value of op k is "v_k".
//   opcodes: 0 CONST imm | 1 LV col | 2 NV col | 3 LA aux_col | 4 NA aux_col | 5 PI i | 6 CH i |
//            7 ADD a b | 8 SUB a b | 9 MUL a b |
//            10 EMIT a (ConstraintConsumer::constraint) | 11 EMIT_TRANSITION a | 12 EMIT_FIRST_ROW a |
//            13 EMIT_LAST_ROW a                      (emitted in the table's source order)
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <utility>
#include <vector>

namespace cprog {

constexpr uint64_t MAGIC = 0x3147525043505445ULL;  // "ETPCPRG1"
enum Op { CONST = 0, LV, NV, LA, NA, PI, CH, ADD, SUB, MUL, EMIT, EMIT_TRANSITION, EMIT_FIRST_ROW, EMIT_LAST_ROW, N_OPCODES };
constexpr int HEADER_WORDS = 8;
constexpr uint32_t COMPACT_THRESHOLD_OPS = 600;

struct Program {
  uint32_t n_ops = 0, n_trace = 0, n_aux = 0, n_pi = 0, n_ch = 0, degree = 0, n_constraints = 0;
  std::vector<uint64_t> ops;  // 2 words per op
  int opcode(uint32_t k) const { return (int)(ops[2 * k] & 0xFF); }
  uint32_t a(uint32_t k) const { return (uint32_t)((ops[2 * k] >> 8) & 0xFFFFFFF); }
  uint32_t b(uint32_t k) const { return (uint32_t)((ops[2 * k] >> 36) & 0xFFFFFFF); }
  uint64_t imm(uint32_t k) const { return ops[2 * k + 1]; }
};

// validates the stream (bounds, SSA order, operand kinds); returns an empty string or the reason
inline std::string parse(const uint64_t* w, size_t n_words, int max_pi, int max_ch, Program* out) {
  char buf[160];
  if (!w || n_words < (size_t)HEADER_WORDS || w[0] != MAGIC) return "constraint program: bad magic / truncated header";
  Program p;
  if (w[1] > (1u << 27)) return "constraint program: too many ops";
  p.n_ops = (uint32_t)w[1]; p.n_trace = (uint32_t)w[2]; p.n_aux = (uint32_t)w[3]; p.n_pi = (uint32_t)w[4];
  p.n_ch = (uint32_t)w[5]; p.degree = (uint32_t)w[6]; p.n_constraints = (uint32_t)w[7];
  if (n_words != (size_t)HEADER_WORDS + 2 * (size_t)p.n_ops) return "constraint program: length does not match n_ops";
  if (p.n_trace == 0 || p.n_trace > 65535 || p.n_aux > 65535) return "constraint program: bad column counts";
  if ((int)p.n_pi > max_pi || (int)p.n_ch > max_ch) return "constraint program: too many public inputs / challenges";
  if (p.degree < 1 || p.degree > 16) return "constraint program: bad constraint degree";
  p.ops.assign(w + HEADER_WORDS, w + n_words);
  std::vector<uint8_t> is_value(p.n_ops, 0);
  uint32_t emitted = 0;
  for (uint32_t k = 0; k < p.n_ops; k++) {
    const int op = p.opcode(k);
    const uint32_t a = p.a(k), b = p.b(k);
    bool ok = true;
    switch (op) {
      case CONST: break;
      case LV: case NV: ok = a < p.n_trace; break;
      case LA: case NA: ok = a < p.n_aux; break;
      case PI: ok = a < p.n_pi; break;
      case CH: ok = a < p.n_ch; break;
      case ADD: case SUB: case MUL: ok = a < k && b < k && is_value[a] && is_value[b]; break;
      case EMIT: case EMIT_TRANSITION: case EMIT_FIRST_ROW: case EMIT_LAST_ROW: ok = a < k && is_value[a]; emitted++; break;
      default: ok = false;
    }
    if (!ok) {
      snprintf(buf, sizeof buf, "constraint program: op %u (opcode %d, a=%u, b=%u) is malformed", k, op, a, b);
      return buf;
    }
    is_value[k] = op < EMIT;
  }
  if (emitted != p.n_constraints) return "constraint program: n_constraints does not match the EMIT ops";
  *out = std::move(p);
  return "";
}

// FNV-1a over the whole stream: the key of the per-context kernel cache
inline uint64_t hash(const Program& p) {
  uint64_t h = 1469598103934665603ULL;
  auto mix = [&](uint64_t x) { for (int i = 0; i < 8; i++) { h ^= (x >> (8 * i)) & 0xFF; h *= 1099511628211ULL; } };
  mix(p.n_trace); mix(p.n_aux); mix(p.n_pi); mix(p.n_ch); mix(p.degree);
  for (uint64_t x : p.ops) mix(x);
  return h;
}

// ---- large programs ------------------------------------------------------------------------------------------------------
// ptxas time is superlinear in the size of one function: a 50 k-op table (a bit-level Keccak-f round: 7 k constraints) does not
// finish as ONE straight-line kernel.  Above SEGMENT_THRESHOLD_OPS the program is cut, at EMIT boundaries, into segments of about
// `segment_ops` ops; every segment becomes a __noinline__ device function that contains the transitive dependencies of ITS emits
// (loads, constants and shared subexpressions are re-materialised per segment — constraints are local, so little is duplicated)
// and folds them into the consumer of the row context it receives by reference.  The kernel calls the segments in order, so the
// consumer sees the emissions in the table's source order, exactly as in the one-function form.  Programs up to the threshold
// (every table and circuit this repository ships today) keep the one-function form.  ETP_CPROG_SEGMENT_OPS overrides both the
// threshold and the segment size (tests force the segmented form on small programs with it).
constexpr uint32_t SEGMENT_THRESHOLD_OPS = 16384, SEGMENT_OPS = 3072;
inline uint32_t segment_ops_override() {
  const char* e = getenv("ETP_CPROG_SEGMENT_OPS");
  if (!e || !*e) return 0;
  const long v = strtol(e, nullptr, 10);
  return v > 0 ? (uint32_t)v : 0;
}

inline void emit_op(std::string& s, const Program& p, uint32_t k) {
  char buf[128];
  const uint32_t a = p.a(k), b = p.b(k);
  switch (p.opcode(k)) {
    case CONST: snprintf(buf, sizeof buf, "  const uint64_t v%u = 0x%llxULL;\n", k, (unsigned long long)p.imm(k)); break;
    case LV: snprintf(buf, sizeof buf, "  const uint64_t v%u = r.lv(q, %u);\n", k, a); break;
    case NV: snprintf(buf, sizeof buf, "  const uint64_t v%u = r.nv(q, %u);\n", k, a); break;
    case LA: snprintf(buf, sizeof buf, "  const uint64_t v%u = r.la(q, %u);\n", k, a); break;
    case NA: snprintf(buf, sizeof buf, "  const uint64_t v%u = r.na(q, %u);\n", k, a); break;
    case PI: snprintf(buf, sizeof buf, "  const uint64_t v%u = q.pi[%u];\n", k, a); break;
    case CH: snprintf(buf, sizeof buf, "  const uint64_t v%u = q.lookup_ch[%u];\n", k, a); break;
    case ADD: snprintf(buf, sizeof buf, "  const uint64_t v%u = gl::add(v%u, v%u);\n", k, a, b); break;
    case SUB: snprintf(buf, sizeof buf, "  const uint64_t v%u = gl::sub(v%u, v%u);\n", k, a, b); break;
    case MUL: snprintf(buf, sizeof buf, "  const uint64_t v%u = gl::mul(v%u, v%u);\n", k, a, b); break;
    case EMIT: snprintf(buf, sizeof buf, "  r.cs.constraint(v%u);\n", a); break;
    case EMIT_TRANSITION: snprintf(buf, sizeof buf, "  r.cs.transition(v%u);\n", a); break;
    case EMIT_FIRST_ROW: snprintf(buf, sizeof buf, "  r.cs.first_row(v%u);\n", a); break;
    case EMIT_LAST_ROW: snprintf(buf, sizeof buf, "  r.cs.last_row(v%u);\n", a); break;
    default: buf[0] = 0;
  }
  s += buf;
}

inline std::string generate_cuda_segmented(const Program& p, bool split_columns, uint32_t segment_ops) {
  std::string s;
  s.reserve(80 * (size_t)p.n_ops + 4096);
  s += "#define ETP_COMPACT_CODE 1\n";
  if (split_columns) s += "#define ETP_SPLIT_COLUMNS 1\n";
  s += "#include \"quotient_rt.cuh\"\n";
  // segment boundaries: [begin, end) op ranges, cut after the first EMIT at or past the size limit
  std::vector<std::pair<uint32_t, uint32_t>> segs;
  uint32_t begin = 0;
  for (uint32_t k = 0; k < p.n_ops; k++) {
    const bool is_emit = p.opcode(k) >= EMIT;
    if ((is_emit && k + 1 - begin >= segment_ops) || k + 1 == p.n_ops) {
      segs.emplace_back(begin, k + 1);
      begin = k + 1;
    }
  }
  std::vector<uint8_t> need(p.n_ops);
  std::string calls;
  char buf[192];
  for (size_t si = 0; si < segs.size(); si++) {
    // transitive dependencies of this segment's emits (ops are in SSA order: one backward sweep)
    std::fill(need.begin(), need.end(), 0);
    bool any = false;
    for (uint32_t k = segs[si].first; k < segs[si].second; k++)
      if (p.opcode(k) >= EMIT) { need[k] = 1; any = true; }
    if (!any) continue;  // trailing ops without an emit contribute nothing
    for (uint32_t k = segs[si].second; k-- > 0;) {
      if (!need[k]) continue;
      const int op = p.opcode(k);
      if (op >= EMIT) need[p.a(k)] = 1;
      else if (op == ADD || op == SUB || op == MUL) need[p.a(k)] = need[p.b(k)] = 1;
    }
    snprintf(buf, sizeof buf, "static __device__ __noinline__ void etp_seg_%zu(const stark::QuotientParams& q, stark::RowCtx& r) {\n", si);
    s += buf;
    for (uint32_t k = 0; k < segs[si].second; k++)
      if (need[k]) emit_op(s, p, k);
    s += "}\n";
    snprintf(buf, sizeof buf, "  etp_seg_%zu(q, r);\n", si);
    calls += buf;
  }
  s += "extern \"C\" __global__ void __launch_bounds__(128) etp_cprog_quotient(const __grid_constant__ stark::QuotientParams q) {\n"
       "  stark::RowCtx r;\n"
       "  if (!stark::quotient_begin(q, r)) return;\n";
  s += calls;
  s += "  stark::quotient_end(q, r);\n}\n";
  return s;
}

// CUDA source of the quotient kernel of this program (body between quotient_begin and quotient_end of
// quotient_rt.cuh).  Straight-line SSA: register allocation and scheduling are ptxas's job.
inline std::string generate_cuda(const Program& p, bool split_columns = false) {
  const uint32_t forced = segment_ops_override();
  if (forced ? p.n_ops > forced : p.n_ops > SEGMENT_THRESHOLD_OPS) return generate_cuda_segmented(p, split_columns, forced ? forced : SEGMENT_OPS);
  std::string s;
  s.reserve(64 * (size_t)p.n_ops + 1024);
  // small programs are fully inlined (as fast as a built-in table); large ones share one copy of the field
  // multiplication and of the consumer so that code size and ptxas time stay bounded
  if (p.n_ops > COMPACT_THRESHOLD_OPS) s += "#define ETP_COMPACT_CODE 1\n";
  if (split_columns) s += "#define ETP_SPLIT_COLUMNS 1\n";  // trace columns through QuotientParams::trace_cols (etp_shard)
  s += "#include \"quotient_rt.cuh\"\n"
       "extern \"C\" __global__ void __launch_bounds__(128) etp_cprog_quotient(const __grid_constant__ stark::QuotientParams q) {\n"
       "  stark::RowCtx r;\n"
       "  if (!stark::quotient_begin(q, r)) return;\n";
  char buf[128];
  for (uint32_t k = 0; k < p.n_ops; k++) {
    const uint32_t a = p.a(k), b = p.b(k);
    switch (p.opcode(k)) {
      case CONST: snprintf(buf, sizeof buf, "  const uint64_t v%u = 0x%llxULL;\n", k, (unsigned long long)p.imm(k)); break;
      case LV: snprintf(buf, sizeof buf, "  const uint64_t v%u = r.lv(q, %u);\n", k, a); break;
      case NV: snprintf(buf, sizeof buf, "  const uint64_t v%u = r.nv(q, %u);\n", k, a); break;
      case LA: snprintf(buf, sizeof buf, "  const uint64_t v%u = r.la(q, %u);\n", k, a); break;
      case NA: snprintf(buf, sizeof buf, "  const uint64_t v%u = r.na(q, %u);\n", k, a); break;
      case PI: snprintf(buf, sizeof buf, "  const uint64_t v%u = q.pi[%u];\n", k, a); break;
      case CH: snprintf(buf, sizeof buf, "  const uint64_t v%u = q.lookup_ch[%u];\n", k, a); break;
      case ADD: snprintf(buf, sizeof buf, "  const uint64_t v%u = gl::add(v%u, v%u);\n", k, a, b); break;
      case SUB: snprintf(buf, sizeof buf, "  const uint64_t v%u = gl::sub(v%u, v%u);\n", k, a, b); break;
      case MUL: snprintf(buf, sizeof buf, "  const uint64_t v%u = gl::mul(v%u, v%u);\n", k, a, b); break;
      case EMIT: snprintf(buf, sizeof buf, "  r.cs.constraint(v%u);\n", a); break;
      case EMIT_TRANSITION: snprintf(buf, sizeof buf, "  r.cs.transition(v%u);\n", a); break;
      case EMIT_FIRST_ROW: snprintf(buf, sizeof buf, "  r.cs.first_row(v%u);\n", a); break;
      case EMIT_LAST_ROW: snprintf(buf, sizeof buf, "  r.cs.last_row(v%u);\n", a); break;
      default: buf[0] = 0;
    }
    s += buf;
  }
  s += "  stark::quotient_end(q, r);\n}\n";
  return s;
}

}  // namespace cprog
