// libetp_b200: the starky / FRI half of the C ABI — auxiliary columns (logUp helper columns, CTL running sums),
// compute_quotient_polys, openings, PolynomialBatch::prove_openings for a general FRI instance, the FRI prover step by
// step (commit phase with evaluation-domain folding, proof of work, query rounds), starky::prover::prove_with_commitment
// with the challenger state passed in and out, and the stand-alone `starky::prover::prove`, all under
// StarkConfig::standard_fast_config() (/root/reference/common/src/prover_state/circuit.rs:204; reached from
// /root/reference/ops/src/lib.rs:52).
// See include/etp_b200.h for the upstream item behind each entry point and DESIGN.md for the proof
// wire format.  Product code: no oracle, no CPU fallback; the only host arithmetic is the transcript
// (Challenger) and the <= 2^8-coefficient FRI final polynomial.
#include <time.h>

#include <algorithm>
#include <cstdlib>
#include <functional>
#include <memory>

#include "cprog.h"
#include "ctx.cuh"
#include "host_field.h"
#include "plonk_kernels.cuh"
#include "shard.cuh"
#include "stark_kernels.cuh"

namespace {

// StarkConfig::standard_fast_config()
constexpr int NUM_CHALLENGES = 2, RATE_BITS = 1, CAP_HEIGHT = 4, POW_BITS = 16, ARITY_BITS = 4, FINAL_POLY_BITS = 5,
              NUM_QUERIES = 84;
constexpr uint64_t PROOF_MAGIC = 0x4232303053544B32ULL;  // "B200STK2"
constexpr int HEADER_WORDS = 24;
constexpr uint64_t AUXSPEC_MAGIC = 0x3153585541505445ULL;  // "ETPAUXS1"

// ---- auxiliary-column descriptors (starky/src/lookup.rs Lookup / Column / Filter, cross_table_lookup.rs CtlZData) ----
// The descriptors stay in their wire form (include/etp_b200.h, etp_table_register_ex); parsing validates them and records
// where each Column / Filter starts, which is all the device kernels need.
struct LookupE {
  std::vector<size_t> col_off, filt_off;  // per looking column
  size_t table_off = 0, freq_off = 0;
  // every column a plain trace column with coefficient 1 and every filter the default: the dedicated fast kernels apply
  bool simple = false;
  std::vector<int> simple_cols;
  int simple_table = 0, simple_freq = 0;
  int n() const { return (int)col_off.size(); }
};
struct ColsetE {
  size_t cols_off = 0, filt_off = 0;
  int n_cols = 0;
};
struct CtlZE {
  int challenge = 0;
  std::vector<ColsetE> sets;
};
struct AuxSpec {
  std::vector<uint64_t> words;
  std::vector<LookupE> lookups;
  std::vector<CtlZE> zs;
};
struct SpecReader {
  const std::vector<uint64_t>& w;
  size_t pos = 0;
  int n_trace;
  bool bad = false;
  std::vector<size_t>* col_pos = nullptr;  // if set: the position of every trace-column index word
  uint64_t rd() { if (pos >= w.size()) { bad = true; return 0; } return w[pos++]; }
  // advances over one Column; *single = trace column index if it is Column::single(c), else -1
  void column(int* single) {
    const uint64_t nl = rd();
    if (nl > 65536) { bad = true; return; }
    int one_col = -1;
    bool plain = nl == 1;
    for (uint64_t i = 0; i < nl && !bad; i++) {
      if (col_pos) col_pos->push_back(pos);
      const uint64_t c = rd(), f = rd();
      if (c >= (uint64_t)n_trace) bad = true;
      if (gl::canon(f) != 1) plain = false;
      one_col = (int)c;
    }
    const uint64_t nn = rd();
    if (nn > 65536) { bad = true; return; }
    for (uint64_t i = 0; i < nn && !bad; i++) { if (col_pos) col_pos->push_back(pos); if (rd() >= (uint64_t)n_trace) bad = true; rd(); }
    const uint64_t k = rd();
    if (single) *single = (plain && nn == 0 && gl::canon(k) == 0) ? one_col : -1;
  }
  // advances over one Filter; *is_default = it is Filter::default() (the constant 1)
  void filter(bool* is_default) {
    const size_t start = pos;
    const uint64_t np = rd();
    if (np > 4096) { bad = true; return; }
    for (uint64_t i = 0; i < np && !bad; i++) { column(nullptr); column(nullptr); }
    const uint64_t nc = rd();
    if (nc > 4096) { bad = true; return; }
    for (uint64_t i = 0; i < nc && !bad; i++) column(nullptr);
    // default filter words: 0 products, 1 constant: Column{0 local, 0 next, constant 1} = [0, 1, 0, 0, 1]
    if (is_default) *is_default = !bad && pos - start == 5 && w[start] == 0 && w[start + 1] == 1 && w[start + 2] == 0 && w[start + 3] == 0 &&
                                  gl::canon(w[start + 4]) == 1;
  }
};
std::string parse_aux_spec(const uint64_t* words, size_t n, int n_trace, int num_challenges, AuxSpec* out, std::vector<size_t>* col_pos = nullptr) {
  AuxSpec a;
  a.words.assign(words, words + n);
  SpecReader r{a.words, 0, n_trace};
  r.col_pos = col_pos;
  if (r.rd() != AUXSPEC_MAGIC) return "auxiliary-column spec: bad magic";
  const uint64_t nl = r.rd(), nz = r.rd();
  if (r.bad || nl > 256 || nz > 256) return "auxiliary-column spec: bad lookup / CTL counts";
  for (uint64_t i = 0; i < nl; i++) {
    LookupE l;
    const uint64_t m = r.rd();
    if (r.bad || m < 1 || m > 65536) return "auxiliary-column spec: bad number of looking columns";
    l.simple = true;
    for (uint64_t j = 0; j < m; j++) {
      int single = -1;
      l.col_off.push_back(r.pos);
      r.column(&single);
      if (single < 0) l.simple = false;
      l.simple_cols.push_back(single);
    }
    for (uint64_t j = 0; j < m; j++) {
      bool def = false;
      l.filt_off.push_back(r.pos);
      r.filter(&def);
      if (!def) l.simple = false;
    }
    l.table_off = r.pos; r.column(&l.simple_table);
    l.freq_off = r.pos; r.column(&l.simple_freq);
    if (l.simple_table < 0 || l.simple_freq < 0) l.simple = false;
    if (r.bad) return "auxiliary-column spec: malformed lookup";
    a.lookups.push_back(std::move(l));
  }
  for (uint64_t i = 0; i < nz; i++) {
    CtlZE z;
    z.challenge = (int)r.rd();
    const uint64_t ns = r.rd();
    if (r.bad || z.challenge < 0 || z.challenge >= num_challenges || ns < 1 || ns > 4096) return "auxiliary-column spec: bad CTL Z header";
    for (uint64_t s = 0; s < ns; s++) {
      ColsetE cs;
      const uint64_t nc = r.rd();
      if (r.bad || nc < 1 || nc > 65536) return "auxiliary-column spec: bad CTL column count";
      cs.n_cols = (int)nc;
      cs.cols_off = r.pos;
      for (uint64_t j = 0; j < nc; j++) r.column(nullptr);
      cs.filt_off = r.pos;
      r.filter(nullptr);
      z.sets.push_back(cs);
    }
    if (r.bad) return "auxiliary-column spec: malformed CTL Z";
    a.zs.push_back(std::move(z));
  }
  if (r.bad || r.pos != n) return "auxiliary-column spec: trailing or missing words";
  *out = std::move(a);
  return "";
}
// spec words of plain lookups: [(looking columns, table column, frequencies column)], default filters
void push_single(std::vector<uint64_t>& w, int c) { w.insert(w.end(), {1, (uint64_t)c, 1, 0, 0}); }
std::vector<uint64_t> simple_spec_words(const std::vector<std::tuple<std::vector<int>, int, int>>& lookups) {
  std::vector<uint64_t> w{AUXSPEC_MAGIC, lookups.size(), 0};
  for (auto& l : lookups) {
    w.push_back(std::get<0>(l).size());
    for (int c : std::get<0>(l)) push_single(w, c);
    for (size_t j = 0; j < std::get<0>(l).size(); j++) w.insert(w.end(), {0, 1, 0, 0, 1});
    push_single(w, std::get<1>(l));
    push_single(w, std::get<2>(l));
  }
  return w;
}

// The trace columns the auxiliary polynomials of a table are computed from (sorted), and the same spec over a matrix that
// holds only those columns, in that order (column-split tables: the leader recovers just these columns).
std::string compact_aux_spec(const AuxSpec& a, int n_trace, int num_challenges, std::vector<int>* used, AuxSpec* compact) {
  std::vector<size_t> pos;
  AuxSpec tmp;
  if (a.words.empty()) { used->clear(); *compact = a; return ""; }
  std::string why = parse_aux_spec(a.words.data(), a.words.size(), n_trace, num_challenges, &tmp, &pos);
  if (!why.empty()) return why;
  std::vector<int> cols;
  for (size_t p : pos) cols.push_back((int)a.words[p]);
  std::sort(cols.begin(), cols.end());
  cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
  std::vector<uint64_t> w = a.words;
  for (size_t p : pos) w[p] = (uint64_t)(std::lower_bound(cols.begin(), cols.end(), (int)a.words[p]) - cols.begin());
  why = parse_aux_spec(w.data(), w.size(), (int)cols.size(), num_challenges, compact);
  *used = cols;
  return why;
}

struct TableInfo {
  int cols = 0, degree = 0, n_pi = 0;
  std::shared_ptr<AuxSpec> aux;    // never null after table_info()
  RegisteredTable* reg = nullptr;  // program-defined table
  int chunk() const { return degree - 1 < 1 ? 1 : degree - 1; }
  int helpers(const LookupE& l) const { return (l.n() + chunk() - 1) / chunk(); }
  int ctl_helpers(const CtlZE& z) const { return z.sets.size() > 1 ? ((int)z.sets.size() + chunk() - 1) / chunk() : 0; }
  int n_lookup_cols(int n_ch) const { int a = 0; for (auto& l : aux->lookups) a += helpers(l) + 1; return a * n_ch; }
  int n_ctl_helpers() const { int a = 0; for (auto& z : aux->zs) a += ctl_helpers(z); return a; }
  int n_ctl_zs() const { return (int)aux->zs.size(); }
  int n_aux(int n_ch) const { return n_lookup_cols(n_ch) + n_ctl_helpers() + n_ctl_zs(); }
  bool lookup() const { return !aux->lookups.empty(); }
  bool ctl() const { return !aux->zs.empty(); }
};
}  // namespace

// a table registered at run time: its description + constraint program + the kernel compiled from it
struct RegisteredTable {
  TableInfo info;
  cprog::Program prog;
  JitKernel kernel;
  JitKernel kernel_split;  // the same program reading its trace columns through a pointer table (etp_shard), compiled on first use
  bool standalone = false;  // etp_program_register: not a starky table (no StarkConfig limits); only the pointer-table kernel exists
  uint64_t* d_spec = nullptr;  // the aux spec words on the device (general lookups / CTL), uploaded on first use
};
void free_registered_tables(etp_ctx* ctx) {
  for (auto* t : ctx->tables) { jit_unload(&t->kernel); jit_unload(&t->kernel_split); cudaFree(t->d_spec); delete t; }
  ctx->tables.clear();
}

namespace {
std::shared_ptr<AuxSpec> empty_spec() { static auto s = std::make_shared<AuxSpec>(); return s; }
std::shared_ptr<AuxSpec> memory_spec() {
  static std::shared_ptr<AuxSpec> s = [] {
    auto a = std::make_shared<AuxSpec>();
    const auto w = simple_spec_words({std::make_tuple(std::vector<int>{stark::M_RANGE_CHECK}, (int)stark::M_COUNTER, (int)stark::M_FREQ)});
    parse_aux_spec(w.data(), w.size(), 21, NUM_CHALLENGES, a.get());
    return a;
  }();
  return s;
}
bool table_info(const etp_ctx* ctx, int t, TableInfo* o) {
  if (t == ETP_TABLE_FIBONACCI) { *o = TableInfo(); o->cols = 2; o->degree = 2; o->n_pi = 3; o->aux = empty_spec(); return true; }
  if (t == ETP_TABLE_MEMORY) { *o = TableInfo(); o->cols = 21; o->degree = 3; o->n_pi = 0; o->aux = memory_spec(); return true; }
  if (ctx && t >= ETP_TABLE_FIRST_REGISTERED && (size_t)(t - ETP_TABLE_FIRST_REGISTERED) < ctx->tables.size()) {
    *o = ctx->tables[t - ETP_TABLE_FIRST_REGISTERED]->info;
    return true;
  }
  return false;
}
int quotient_factor(const TableInfo& ti) { return ti.degree - 1 < 1 ? 1 : ti.degree - 1; }
int log2_ceil(int x) { int l = 0; while ((1 << l) < x) l++; return l; }
int fri_total_arities(const etp_fri_params& p) { int t = 0; for (int i = 0; i < p.n_reductions; i++) t += p.reduction_arity_bits[i]; return t; }
// FriConfig::fri_params with FriReductionStrategy::ConstantArityBits(4, 5)
void fri_params_make(int degree_bits, int rate_bits, int cap_height, int pow_bits, int num_queries, etp_fri_params* p) {
  memset(p, 0, sizeof *p);
  p->degree_bits = degree_bits; p->rate_bits = rate_bits; p->cap_height = cap_height; p->proof_of_work_bits = pow_bits;
  p->num_query_rounds = num_queries;
  int db = degree_bits;
  while (db > FINAL_POLY_BITS && db + rate_bits - ARITY_BITS >= cap_height && p->n_reductions < 16) {
    p->reduction_arity_bits[p->n_reductions++] = ARITY_BITS;
    db -= ARITY_BITS;
  }
}
etp_fri_params standard_fast_params(int degree_bits) {
  etp_fri_params p;
  fri_params_make(degree_bits, RATE_BITS, CAP_HEIGHT, POW_BITS, NUM_QUERIES, &p);
  return p;
}

struct PhaseTimer {
  etp_ctx* ctx;
  std::vector<std::pair<const char*, cudaEvent_t>> marks;
  std::vector<double> host_ms;  // host clock at each mark (ETP_TRACE=1 prints it: tells a device phase from a host stall)
  explicit PhaseTimer(etp_ctx* c) : ctx(c) { mark("start"); }
  void mark(const char* name) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, ctx->stream);
    marks.emplace_back(name, e);
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    host_ms.push_back(ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6);
  }
  void finish() {
    ctx->timings.clear();
    cudaStreamSynchronize(ctx->stream);
    const bool trace = getenv("ETP_TRACE") != nullptr;
    for (size_t i = 1; i < marks.size(); i++) {
      float ms = 0;
      cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second);
      ctx->timings.emplace_back(marks[i].first, ms);
      if (trace) fprintf(stderr, "[etp trace] %-50s device %8.3f ms   host enqueue %8.3f ms\n", marks[i].first, ms, host_ms[i] - host_ms[i - 1]);
    }
  }
  ~PhaseTimer() { for (auto& m : marks) cudaEventDestroy(m.second); }
};

// ETP_TRACE=1: host clock between points of a host-heavy function
struct HostTrace {
  const char* what;
  bool on;
  double last;
  static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
  explicit HostTrace(const char* w) : what(w), on(getenv("ETP_TRACE") != nullptr), last(on ? now() : 0) {}
  void mark(const char* name) {
    if (!on) return;
    const double t = now();
    fprintf(stderr, "[etp trace]   %s: %-34s host %8.3f ms\n", what, name, t - last);
    last = t;
  }
};

unsigned blocks_for(size_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

// ctx->d_pow_result[1] doubles as the "a batch inverse met a zero" flag of the call in flight
unsigned long long* zero_flag(etp_ctx* ctx) { return (unsigned long long*)(ctx->d_pow_result + 1); }
int reset_zero_flag(etp_ctx* ctx) {
  ETP_CUDA(ctx, cudaMemsetAsync(zero_flag(ctx), 0, 8, ctx->stream));
  return ETP_OK;
}
// call after a synchronisation point of the stream
int check_zero_flag(etp_ctx* ctx, const char* what) {
  unsigned long long f = 0;
  ETP_CUDA(ctx, cudaMemcpyAsync(&f, zero_flag(ctx), 8, cudaMemcpyDeviceToHost, ctx->stream));
  ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (f) return etp_fail(ctx, ETP_ERR_PROOF, "Tried to invert zero (%s)", what);
  return ETP_OK;
}
int batch_inverse_dev(etp_ctx* ctx, const uint64_t* in, uint64_t* out, size_t n) {
  const int threads = 128;
  const size_t per_block = (size_t)threads * stark::INV_K;
  stark::batch_inverse<<<blocks_for(n, (int)per_block), threads, 0, ctx->stream>>>(in, out, n, zero_flag(ctx));
  ETP_LAUNCH_CHECK(ctx);
  return ETP_OK;
}

// exclusive prefix sum of n field elements
int exclusive_scan_dev(etp_ctx* ctx, const uint64_t* in, uint64_t* out, size_t n) {
  const size_t nb = (n + stark::SCAN_BLOCK - 1) / stark::SCAN_BLOCK;
  DevBuf<uint64_t> sums(ctx);
  ETP_TRY(sums.alloc(nb));
  stark::scan_block_sums<<<(unsigned)nb, stark::SCAN_THREADS, 0, ctx->stream>>>(in, n, sums.p);
  ETP_LAUNCH_CHECK(ctx);
  stark::scan_sums_serial<<<1, 1, 0, ctx->stream>>>(sums.p, nb);
  ETP_LAUNCH_CHECK(ctx);
  stark::scan_finish<<<(unsigned)nb, stark::SCAN_THREADS, 0, ctx->stream>>>(in, n, sums.p, out);
  ETP_LAUNCH_CHECK(ctx);
  return ETP_OK;
}

int spec_on_device(etp_ctx* ctx, const TableInfo& ti, const uint64_t** out) {
  if (!ti.reg) return etp_fail(ctx, ETP_ERR_STATE, "internal error: general auxiliary columns on a built-in table");
  if (!ti.reg->d_spec) {
    const auto& w = ti.aux->words;
    ETP_CUDA(ctx, cudaMalloc((void**)&ti.reg->d_spec, w.size() * 8));
    ETP_CUDA(ctx, cudaMemcpyAsync(ti.reg->d_spec, w.data(), w.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  *out = ti.reg->d_spec;
  return ETP_OK;
}

// All auxiliary polynomials of a table (values on the trace domain): lookup_helper_columns for every lookup and challenge,
// then the CTL helper columns, then the CTL Z columns (starky prover.rs: auxiliary_polys = lookup columns ++
// get_ctl_auxiliary_polys(ctl_data); cross_table_lookup.rs: ctl_helper_polys() ++ ctl_z_polys()).
int aux_columns(etp_ctx* ctx, const TableInfo& ti, int log_n, const uint64_t* trace, size_t stride, const uint64_t* lookup_ch, int n_ch,
                const uint64_t* ctl_ch, uint64_t* aux, const uint64_t* d_spec_override = nullptr) {
  if (!ti.lookup() && !ti.ctl()) return ETP_OK;
  const size_t n = (size_t)1 << log_n;
  size_t max_m = 1;
  for (auto& l : ti.aux->lookups) max_m = (size_t)l.n() > max_m ? (size_t)l.n() : max_m;
  for (auto& z : ti.aux->zs) max_m = z.sets.size() > max_m ? z.sets.size() : max_m;
  DevBuf<uint64_t> den(ctx), filt(ctx), term(ctx), freq(ctx), prefix(ctx);
  DevBuf<int> d_cols(ctx);
  ETP_TRY(den.alloc((max_m + 1) * n));
  ETP_TRY(term.alloc(n));
  ETP_TRY(d_cols.alloc(max_m + 1));
  bool general = ti.ctl();
  for (auto& l : ti.aux->lookups) general = general || !l.simple;
  const uint64_t* spec = nullptr;
  if (general) {
    if (d_spec_override) spec = d_spec_override;
    else ETP_TRY(spec_on_device(ctx, ti, &spec));
    ETP_TRY(filt.alloc(max_m * n));
    ETP_TRY(freq.alloc(n));
    ETP_TRY(prefix.alloc(n));
  }
  const unsigned gb = blocks_for(n, 256);
  // auxiliary column order (starky prover.rs): for each lookup, for each challenge: helpers..., Z
  uint64_t* out = aux;
  for (auto& l : ti.aux->lookups) {
    const int m = l.n(), nh = ti.helpers(l);
    if (l.simple) {
      std::vector<int> cols(l.simple_cols);
      cols.push_back(l.simple_table);
      ETP_CUDA(ctx, cudaMemcpyAsync(d_cols.p, cols.data(), cols.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
      ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `cols` dies at the end of this iteration
    }
    for (int k = 0; k < n_ch; k++) {
      const uint64_t ch = gl::canon(lookup_ch[k]);
      if (l.simple) {
        stark::lookup_denominators<<<blocks_for((size_t)(m + 1) * n, 256), 256, 0, ctx->stream>>>(trace, stride, d_cols.p, m + 1, ch, n, den.p);
        ETP_LAUNCH_CHECK(ctx);
        ETP_TRY(batch_inverse_dev(ctx, den.p, den.p, (size_t)(m + 1) * n));
        stark::lookup_terms<<<gb, 256, 0, ctx->stream>>>(den.p, m, ti.chunk(), trace + (size_t)l.simple_freq * stride, n, out, term.p);
        ETP_LAUNCH_CHECK(ctx);
      } else {
        // GrandProductChallenge { beta: 1, gamma: challenge } on one-column sets, each with its filter
        for (int j = 0; j < m; j++) {
          stark::aux_colset_eval<<<gb, 256, 0, ctx->stream>>>(trace, stride, (uint32_t)n, spec + l.col_off[j], 1, spec + l.filt_off[j], 1, ch,
                                                             den.p + (size_t)j * n, filt.p + (size_t)j * n);
          ETP_LAUNCH_CHECK(ctx);
        }
        stark::aux_column_eval<<<gb, 256, 0, ctx->stream>>>(trace, stride, (uint32_t)n, spec + l.table_off, ch, den.p + (size_t)m * n);
        ETP_LAUNCH_CHECK(ctx);
        stark::aux_column_eval<<<gb, 256, 0, ctx->stream>>>(trace, stride, (uint32_t)n, spec + l.freq_off, 0, freq.p);
        ETP_LAUNCH_CHECK(ctx);
        ETP_TRY(batch_inverse_dev(ctx, den.p, den.p, (size_t)(m + 1) * n));
        stark::aux_helper_terms<<<gb, 256, 0, ctx->stream>>>(den.p, filt.p, m, ti.chunk(), n, freq.p, den.p + (size_t)m * n, out, term.p);
        ETP_LAUNCH_CHECK(ctx);
      }
      ETP_TRY(exclusive_scan_dev(ctx, term.p, out + (size_t)nh * n, n));
      out += (size_t)(nh + 1) * n;
    }
  }
  // CTL: helper columns of every CtlZData, then every Z
  uint64_t* helpers = out;
  uint64_t* zs = out + (size_t)ti.n_ctl_helpers() * n;
  for (size_t zi = 0; zi < ti.aux->zs.size(); zi++) {
    const CtlZE& z = ti.aux->zs[zi];
    if (!ctl_ch) return etp_fail(ctx, ETP_ERR_INVALID, "the table requires CTLs but no CTL challenges were given");
    const uint64_t beta = gl::canon(ctl_ch[2 * z.challenge]), gamma = gl::canon(ctl_ch[2 * z.challenge + 1]);
    const int S = (int)z.sets.size(), nh = ti.ctl_helpers(z);
    for (int s = 0; s < S; s++) {
      stark::aux_colset_eval<<<gb, 256, 0, ctx->stream>>>(trace, stride, (uint32_t)n, spec + z.sets[s].cols_off, z.sets[s].n_cols,
                                                         spec + z.sets[s].filt_off, beta, gamma, den.p + (size_t)s * n, filt.p + (size_t)s * n);
      ETP_LAUNCH_CHECK(ctx);
    }
    ETP_TRY(batch_inverse_dev(ctx, den.p, den.p, (size_t)S * n));
    stark::aux_helper_terms<<<gb, 256, 0, ctx->stream>>>(den.p, filt.p, S, ti.chunk(), n, nullptr, nullptr, nh ? helpers : nullptr, term.p);
    ETP_LAUNCH_CHECK(ctx);
    helpers += (size_t)nh * n;
    // partial_sums: Z[i] = sum_{j >= i} term[j]
    ETP_TRY(exclusive_scan_dev(ctx, term.p, prefix.p, n));
    stark::suffix_from_prefix<<<gb, 256, 0, ctx->stream>>>(prefix.p, term.p, n, zs + zi * n);
    ETP_LAUNCH_CHECK(ctx);
  }
  return ETP_OK;
}

int ext_pow_table(etp_ctx* ctx, gl::Ext z, int bits, DevBuf<uint64_t>& lo, DevBuf<uint64_t>& hi, stark::ExtPowTable* out) {
  const int lo_bits = (bits + 1) / 2;
  const size_t n_lo = (size_t)1 << lo_bits, n_hi = (size_t)1 << (bits - lo_bits);
  std::vector<uint64_t> hl(2 * n_lo), hh(2 * n_hi);
  gl::Ext cur = gl::ext(1, 0);
  for (size_t i = 0; i < n_lo; i++) { cur = gl::ecanon(cur); hl[2 * i] = cur.c0; hl[2 * i + 1] = cur.c1; cur = gl::emul(cur, z); }
  const gl::Ext step = gl::ecanon(cur);
  cur = gl::ext(1, 0);
  for (size_t i = 0; i < n_hi; i++) { cur = gl::ecanon(cur); hh[2 * i] = cur.c0; hh[2 * i + 1] = cur.c1; cur = gl::emul(cur, step); }
  ETP_TRY(lo.alloc(2 * n_lo));
  ETP_TRY(hi.alloc(2 * n_hi));
  ETP_CUDA(ctx, cudaMemcpyAsync(lo.p, hl.data(), hl.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  ETP_CUDA(ctx, cudaMemcpyAsync(hi.p, hh.data(), hh.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  out->lo = lo.p; out->hi = hi.p; out->lo_bits = lo_bits;
  return ETP_OK;
}

// evaluate all polynomials of a batch at z0 and z1: out0/out1 get n_cols ext values
int eval_coeffs(etp_ctx* ctx, const uint64_t* coeffs, size_t n_len, size_t n_cols, const stark::ExtPowTable& t0, const stark::ExtPowTable& t1,
                std::vector<gl::Ext>& out0, std::vector<gl::Ext>& out1) {
  const uint32_t n = (uint32_t)n_len;
  const int np = (int)n_cols;
  out0.assign(np, gl::ext(0, 0));
  out1.assign(np, gl::ext(0, 0));
  if (np == 0) return ETP_OK;
  const unsigned gx = blocks_for(n, stark::OPEN_THREADS * stark::OPEN_CHUNK);
  const unsigned gy = (unsigned)np;
  DevBuf<uint64_t> partial(ctx);
  ETP_TRY(partial.alloc((size_t)gx * np * 4));
  stark::eval_polys_at_two_points<<<dim3(gx, gy), stark::OPEN_THREADS, 0, ctx->stream>>>(coeffs, n_len, np, n, t0, t1, partial.p);
  ETP_LAUNCH_CHECK(ctx);
  std::vector<uint64_t> host((size_t)gx * np * 4);
  ETP_CUDA(ctx, cudaMemcpyAsync(host.data(), partial.p, host.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (unsigned blk = 0; blk < gx; blk++)
    for (int q = 0; q < np; q++) {
      const uint64_t* v = &host[((size_t)blk * np + q) * 4];
      out0[q] = gl::eadd(out0[q], gl::ext(v[0], v[1]));
      out1[q] = gl::eadd(out1[q], gl::ext(v[2], v[3]));
    }
  for (auto& e : out0) e = gl::ecanon(e);
  for (auto& e : out1) e = gl::ecanon(e);
  return ETP_OK;
}

int eval_batch(etp_ctx* ctx, const etp_batch* b, const stark::ExtPowTable& t0, const stark::ExtPowTable& t1,
               std::vector<gl::Ext>& out0, std::vector<gl::Ext>& out1) {
  return eval_coeffs(ctx, b->coeffs, b->n(), b->n_cols, t0, t1, out0, out1);
}

// The trace LDE a quotient is computed over: one matrix of this context (a PolynomialBatch), or one device pointer per
// column (a column-split table: own columns and the peers', read over NVLink).
struct TraceView {
  const uint64_t* lde = nullptr;
  size_t stride = 0;
  const uint64_t* const* cols_dev = nullptr;  // device array of n_cols column bases, or nullptr
  size_t n_cols = 0;
  int log_n = 0, rate_bits = 0;
};

// n_scalars challenge scalars: lookup challenges [0, NUM_CHALLENGES), then the CTL (beta, gamma) pairs
int compute_quotient(etp_ctx* ctx, int table, const TraceView& trace, etp_batch* aux, const uint64_t* scalars, int n_scalars,
                     const uint64_t* pi, const uint64_t* alphas, int n_alphas, uint64_t* out_dev) {
  TableInfo ti;
  if (!table_info(ctx, table, &ti)) return etp_fail(ctx, ETP_ERR_INVALID, "unknown table %d", table);
  if ((int)trace.n_cols != ti.cols) return etp_fail(ctx, ETP_ERR_INVALID, "trace batch has the wrong number of columns");
  if (n_alphas < 1 || n_alphas > stark::MAX_CHALLENGES) return etp_fail(ctx, ETP_ERR_INVALID, "unsupported number of challenges");
  if (n_scalars < 0 || n_scalars > stark::MAX_CH_SCALARS) return etp_fail(ctx, ETP_ERR_INVALID, "too many challenge scalars");
  const int n_lookup_ch = n_scalars < NUM_CHALLENGES ? n_scalars : NUM_CHALLENGES;
  const int n_aux = ti.n_aux(n_lookup_ch);
  if ((ti.lookup() || ti.ctl()) && (!aux || (int)aux->n_cols != n_aux))
    return etp_fail(ctx, ETP_ERR_INVALID, "auxiliary batch does not match the table's lookups / CTLs");
  if (ti.reg && ((int)ti.reg->prog.n_aux != n_aux || (int)ti.reg->prog.n_ch > n_scalars))
    return etp_fail(ctx, ETP_ERR_INVALID, "constraint program expects %u auxiliary columns / %u challenge scalars, got %d / %d",
                    ti.reg->prog.n_aux, ti.reg->prog.n_ch, n_aux, n_scalars);
  const int log_n = trace.log_n, rate_bits = trace.rate_bits;
  const bool split = trace.cols_dev != nullptr;
  if (ti.reg && ti.reg->standalone && !split)
    return etp_fail(ctx, ETP_ERR_INVALID, "a standalone program (etp_program_register) is evaluated through etp_compute_quotient_polys_cols_dev");
  const int factor = quotient_factor(ti), qbits = log2_ceil(factor);
  if (qbits > rate_bits)
    return etp_fail(ctx, ETP_ERR_INVALID, "Having constraints of degree higher than the rate is not supported yet.");
  if (qbits > 3) return etp_fail(ctx, ETP_ERR_INVALID, "quotient degree factors above 8 are not supported");
  const int log_size = log_n + qbits, log_lde = log_n + rate_bits;
  const size_t size = (size_t)1 << log_size;
  stark::QuotientParams q{};
  q.trace = trace.lde; q.trace_stride = trace.stride; q.trace_cols = trace.cols_dev;
  q.aux = aux ? aux->lde : nullptr; q.aux_stride = aux ? aux->lde_n() : 0;
  q.log_lde = log_lde; q.log_size = log_size; q.step_log = rate_bits - qbits; q.next_step = 1 << qbits;
  ETP_TRY(get_pow_table(ctx, gl::root_of_unity(log_size), log_size, gl::GENERATOR, &q.coset));
  // ZeroPolyOnCoset: Z_H(x_i) = g^n * w_{2^qbits}^(i mod 2^qbits) - 1
  stark::ZhVals zh{};
  {
    uint64_t g_pow_n = gl::GENERATOR;
    for (int i = 0; i < log_n; i++) g_pow_n = gl::sqr(g_pow_n);
    const uint64_t w = gl::root_of_unity(qbits);
    uint64_t cur = 1;
    for (int i = 0; i < (1 << qbits); i++) {
      zh.v[i] = gl::canon(gl::sub(gl::mul(g_pow_n, cur), 1));
      q.zh_inv[i] = gl::canon(gl::inv(zh.v[i]));
      cur = gl::mul(cur, w);
    }
  }
  const uint64_t g = gl::root_of_unity(log_n);
  q.last = gl::canon(gl::inv(g));
  for (int j = 0; j < n_alphas; j++) q.alphas[j] = gl::canon(alphas[j]);
  q.n_alphas = n_alphas;
  for (int j = 0; j < n_scalars; j++) q.lookup_ch[j] = gl::canon(scalars[j]);
  q.n_lookup_ch = n_lookup_ch;
  for (int j = 0; j < ti.n_pi && j < stark::MAX_PUBLIC_INPUTS; j++) q.pi[j] = gl::canon(pi[j]);
  // Lagrange selectors at every point of the quotient coset
  DevBuf<uint64_t> lag(ctx), qvals(ctx), scratch(ctx);
  ETP_TRY(lag.alloc(2 * size));
  const unsigned gb = blocks_for(size, 128);
  stark::lagrange_denominators<<<gb, 128, 0, ctx->stream>>>(log_lde, log_size, q.step_log, q.coset, gl::canon((uint64_t)1 << log_n), g,
                                                           lag.p);
  ETP_LAUNCH_CHECK(ctx);
  ETP_TRY(batch_inverse_dev(ctx, lag.p, lag.p, 2 * size));
  stark::lagrange_finish<<<gb, 128, 0, ctx->stream>>>(log_lde, log_size, q.step_log, (1 << qbits) - 1, zh, lag.p);
  ETP_LAUNCH_CHECK(ctx);
  q.lag_first = lag.p; q.lag_last = lag.p + size;
  ETP_TRY(qvals.alloc((size_t)n_alphas * size));
  q.out = qvals.p;
  // powers of the alphas for the consumer's power-form fold: constraint i of N is weighted by alpha^(N-1-i)
  if (ti.reg) q.n_constraints = (int)ti.reg->prog.n_constraints;
  else if (table == ETP_TABLE_FIBONACCI) q.n_constraints = stark::Table<0>::CONSTRAINTS;
  else q.n_constraints = stark::Table<1>::CONSTRAINTS + stark::Table<1>::LOOKUP_CONSTRAINTS_PER_CHALLENGE * n_lookup_ch;
  DevBuf<uint64_t> apow(ctx);
  std::vector<uint64_t> apow_host((size_t)n_alphas * q.n_constraints);
  for (int j = 0; j < n_alphas; j++) {
    uint64_t cur = 1;
    for (int e = 0; e < q.n_constraints; e++) { apow_host[(size_t)j * q.n_constraints + e] = gl::canon(cur); cur = gl::mul(cur, q.alphas[j]); }
  }
  ETP_TRY(apow.alloc(apow_host.size() ? apow_host.size() : 1));
  ETP_CUDA(ctx, cudaMemcpyAsync(apow.p, apow_host.data(), apow_host.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  q.alpha_pows = apow.p;
  if (ti.reg) {
    if (split && !ti.reg->kernel_split.kernel) {
      std::vector<char> cubin;
      std::string log;
      ETP_TRY(jit_compile(ctx, cprog::generate_cuda(ti.reg->prog, true), &cubin, &log));
      ETP_TRY(jit_load(ctx, cubin, "etp_cprog_quotient", &ti.reg->kernel_split));
    }
    void* args[] = {(void*)&q};
    ETP_CUDA(ctx, cudaLaunchKernel((const void*)(split ? ti.reg->kernel_split : ti.reg->kernel).kernel, dim3(gb), dim3(128), args, 0, ctx->stream));
    ctx->launches++;
  } else if (table == ETP_TABLE_FIBONACCI) {
    if (split) stark::quotient_kernel<0, true><<<gb, 128, 0, ctx->stream>>>(q);
    else stark::quotient_kernel<0><<<gb, 128, 0, ctx->stream>>>(q);
  } else {
    if (split) stark::quotient_kernel<1, true><<<gb, 128, 0, ctx->stream>>>(q);
    else stark::quotient_kernel<1><<<gb, 128, 0, ctx->stream>>>(q);
  }
  ETP_LAUNCH_CHECK(ctx);
  // coset_ifft(7) of each challenge's values, then split into `factor` chunks of n coefficients.
  // size == factor * n whenever factor is a power of two; otherwise the tail must vanish (trim_to_len).
  DevBuf<uint64_t> coeffs(ctx);
  ETP_TRY(coeffs.alloc((size_t)n_alphas * size));
  ETP_TRY(scratch.alloc((size_t)n_alphas * size));
  NttArgs a;
  a.in = qvals.p; a.in_stride = size; a.n_in = (uint32_t)size; a.out = coeffs.p; a.out_stride = size;
  a.scratch = scratch.p; a.scratch_stride = size; a.log_n = log_size; a.n_cols = n_alphas;
  a.inverse = true; a.natural_out = true; a.coset_shift = gl::GENERATOR;
  ETP_TRY(ntt_run(ctx, a));
  const size_t n = (size_t)1 << log_n;
  for (int j = 0; j < n_alphas; j++)
    ETP_CUDA(ctx, cudaMemcpyAsync(out_dev + (size_t)j * factor * n, coeffs.p + (size_t)j * size, (size_t)factor * n * 8,
                                  cudaMemcpyDeviceToDevice, ctx->stream));
  return ETP_OK;
}

int compute_quotient(etp_ctx* ctx, int table, etp_batch* trace, etp_batch* aux, const uint64_t* scalars, int n_scalars,
                     const uint64_t* pi, const uint64_t* alphas, int n_alphas, uint64_t* out_dev) {
  if (!trace) return etp_fail(ctx, ETP_ERR_INVALID, "trace batch has the wrong number of columns");
  TraceView tv;
  tv.lde = trace->lde; tv.stride = trace->lde_n(); tv.n_cols = trace->n_cols; tv.log_n = trace->log_n; tv.rate_bits = trace->rate_bits;
  return compute_quotient(ctx, table, tv, aux, scalars, n_scalars, pi, alphas, n_alphas, out_dev);
}

int pow_grind(etp_ctx* ctx, const uint64_t state[12], int pos, int bits, uint64_t* witness) {
  if (pos < 0 || pos >= 8 || bits < 0 || bits > 40) return etp_fail(ctx, ETP_ERR_INVALID, "pow_grind: bad arguments");
  DevBuf<uint64_t> d_state(ctx);
  ETP_TRY(d_state.alloc(12));
  uint64_t st[12];
  for (int i = 0; i < 12; i++) st[i] = gl::canon(state[i]);
  ETP_CUDA(ctx, cudaMemcpyAsync(d_state.p, st, sizeof st, cudaMemcpyHostToDevice, ctx->stream));
  // candidates are scanned in increasing order, one batch per launch; the smallest hit of the first batch that has
  // one is the smallest witness overall.  2^bits candidates are needed on average: start with 2 x that, then grow.
  uint64_t batch = (uint64_t)1 << (bits + 1 < 12 ? 12 : (bits + 1 > 22 ? 22 : bits + 1));
  for (uint64_t base = 0;; base += batch) {
    if (base) batch = batch < ((uint64_t)1 << 22) ? batch << 1 : batch;
    unsigned long long init = ~0ull, found = ~0ull;
    ETP_CUDA(ctx, cudaMemcpyAsync(ctx->d_pow_result, &init, 8, cudaMemcpyHostToDevice, ctx->stream));
    stark::pow_grind<<<(unsigned)(batch / 128), 128, 0, ctx->stream>>>(d_state.p, pos, bits, base, (unsigned long long*)ctx->d_pow_result);
    ETP_LAUNCH_CHECK(ctx);
    ETP_CUDA(ctx, cudaMemcpyAsync(&found, ctx->d_pow_result, 8, cudaMemcpyDeviceToHost, ctx->stream));
    ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (found != ~0ull) { *witness = found; return ETP_OK; }
    if (base + batch >= ((uint64_t)1 << 44)) return etp_fail(ctx, ETP_ERR_PROOF, "Proof of work failed. This is highly unlikely!");
  }
}

size_t fri_proof_words(const size_t* oracle_cols, size_t n_oracles, const etp_fri_params& p) {
  const size_t cap = (size_t)4 << p.cap_height;
  const int log_lde = p.degree_bits + p.rate_bits;
  size_t per_query = 0;
  for (size_t o = 0; o < n_oracles; o++) per_query += oracle_cols[o] + 4 * (size_t)(log_lde - p.cap_height);
  int bits = log_lde;
  for (int l = 0; l < p.n_reductions; l++) {
    bits -= p.reduction_arity_bits[l];
    per_query += 2 * ((size_t)1 << p.reduction_arity_bits[l]) + 4 * (size_t)(bits - p.cap_height);
  }
  const size_t final_len = (size_t)1 << (p.degree_bits - fri_total_arities(p));
  return cap * p.n_reductions + (size_t)p.num_query_rounds * per_query + 2 * final_len + 1;
}

size_t proof_words(const TableInfo& ti, int log_n) {
  const etp_fri_params fp = standard_fast_params(log_n);
  const int n_aux = ti.n_aux(NUM_CHALLENGES), n_quot = quotient_factor(ti) * NUM_CHALLENGES;
  const size_t cap = (size_t)4 << CAP_HEIGHT;
  size_t w = HEADER_WORDS + cap * (2 + (n_aux ? 1 : 0)) + 2 * (size_t)(2 * ti.cols + 2 * n_aux + n_quot) + ti.n_ctl_zs();
  size_t oc[3];
  size_t no = 0;
  oc[no++] = ti.cols;
  if (n_aux) oc[no++] = n_aux;
  oc[no++] = n_quot;
  return w + fri_proof_words(oc, no, fp) + ti.n_pi;
}

int check_fri_params(etp_ctx* ctx, const etp_fri_params& p) {
  if (p.degree_bits < 0 || p.rate_bits < 0 || p.degree_bits + p.rate_bits > 30 || p.cap_height < 0 || p.n_reductions < 0 || p.n_reductions > 16 ||
      p.proof_of_work_bits < 0 || p.proof_of_work_bits > 40 || p.num_query_rounds < 0 || p.num_query_rounds > 4096)
    return etp_fail(ctx, ETP_ERR_INVALID, "bad FRI parameters");
  for (int i = 0; i < p.n_reductions; i++)
    if (p.reduction_arity_bits[i] != ARITY_BITS) return etp_fail(ctx, ETP_ERR_INVALID, "only arity-16 FRI reductions are supported (ConstantArityBits(4, 5))");
  if (fri_total_arities(p) > p.degree_bits + p.rate_bits - p.cap_height || p.degree_bits + p.rate_bits < p.cap_height)
    return etp_fail(ctx, ETP_ERR_INVALID, "FRI total reduction arity is too large.");
  if (fri_total_arities(p) > p.degree_bits) return etp_fail(ctx, ETP_ERR_INVALID, "FRI reductions exceed the degree");
  return ETP_OK;
}
}  // namespace

// =================================================================================================
// the FRI prover, step by step (plonky2/src/fri/prover.rs)
// =================================================================================================
struct FriLayer {
  uint64_t* values = nullptr;  // n ext, bit-reversed order (== leaves, 2^arity ext per row)
  uint64_t* levels = nullptr;
  size_t n_leaves = 0;
  int log_n = 0;
  std::vector<uint64_t> cap;
};
struct etp_fri_state {
  etp_ctx* ctx = nullptr;
  etp_fri_params p{};
  std::vector<FriLayer> layers;  // n_reductions + 1 entries; layers[l].values exists once layer l has been produced
  int committed = 0, folded = 0;  // layers with a tree / layers folded away
  uint64_t shift = gl::GENERATOR;  // coset shift of the current layer: 7^(16^l)
  ~etp_fri_state() {
    for (auto& x : layers) { dev_free(ctx, x.values); dev_free(ctx, x.levels); }
  }
};
namespace {
// takes ownership of `values` (a dev_alloc'ed block of 2 * 2^(degree_bits + rate_bits) words)
int fri_begin_owned(etp_ctx* ctx, uint64_t* values, const etp_fri_params& p, etp_fri_state** out) {
  auto* s = new etp_fri_state();
  s->ctx = ctx; s->p = p;
  s->layers.resize(p.n_reductions + 1);
  s->layers[0].values = values;
  s->layers[0].log_n = p.degree_bits + p.rate_bits;
  *out = s;
  return ETP_OK;
}
int fri_commit_layer(etp_fri_state* s, uint64_t* cap_out) {
  etp_ctx* ctx = s->ctx;
  if (s->committed >= s->p.n_reductions || s->committed != s->folded) return etp_fail(ctx, ETP_ERR_STATE, "fri_commit_layer: no layer to commit");
  FriLayer& L = s->layers[s->committed];
  const size_t cap_words = (size_t)4 << s->p.cap_height;
  L.n_leaves = ((size_t)1 << L.log_n) >> ARITY_BITS;
  L.cap.resize(cap_words);
  ETP_TRY(dev_alloc(ctx, levels_words(L.n_leaves, s->p.cap_height) * 8, (void**)&L.levels));
  ETP_TRY(launch_leaf_hash_rowmajor(ctx, L.values, 2 << ARITY_BITS, L.n_leaves, L.levels));
  ETP_TRY(merkle_build_levels(ctx, L.levels, L.n_leaves, s->p.cap_height, L.cap.data()));
  if (cap_out) memcpy(cap_out, L.cap.data(), cap_words * 8);
  s->committed++;
  return ETP_OK;
}
int fri_fold(etp_fri_state* s, gl::Ext beta) {
  etp_ctx* ctx = s->ctx;
  if (s->folded >= s->p.n_reductions || s->committed != s->folded + 1) return etp_fail(ctx, ETP_ERR_STATE, "fri_fold: commit the layer first");
  FriLayer& L = s->layers[s->folded];
  FriLayer& Nx = s->layers[s->folded + 1];
  const size_t cur_n = (size_t)1 << L.log_n;
  Nx.log_n = L.log_n - ARITY_BITS;
  ETP_TRY(dev_alloc(ctx, (2 * cur_n >> ARITY_BITS) * 8, (void**)&Nx.values));
  stark::FoldParams f{};
  f.in = L.values; f.out = Nx.values; f.log_n = L.log_n; f.beta = gl::ecanon(beta);
  ETP_TRY(get_pow_table(ctx, gl::canon(gl::inv(gl::root_of_unity(L.log_n))), L.log_n, gl::canon(gl::inv(s->shift)), &f.x0_inv));
  {
    const uint64_t wi = gl::canon(gl::inv(gl::root_of_unity(ARITY_BITS)));
    uint64_t cur = 1;
    for (int i = 0; i < 16; i++) { f.w16_inv_pows[i] = gl::canon(cur); cur = gl::mul(cur, wi); }
  }
  f.inv16 = gl::canon(gl::inv(16));
  stark::fri_fold16<<<blocks_for(L.n_leaves, 128), 128, 0, ctx->stream>>>(f);
  ETP_LAUNCH_CHECK(ctx);
  s->shift = gl::canon(gl::pow(s->shift, 16));
  s->folded++;
  return ETP_OK;
}
// final polynomial: coset iDFT of the last layer's values on the host (<= 2^8 points)
int fri_final_poly(etp_fri_state* s, std::vector<uint64_t>* final_coeffs) {
  etp_ctx* ctx = s->ctx;
  if (s->folded != s->p.n_reductions) return etp_fail(ctx, ETP_ERR_STATE, "fri_final_poly: layers remain to be folded");
  const FriLayer& F = s->layers[s->p.n_reductions];
  if (F.log_n > 16) return etp_fail(ctx, ETP_ERR_INVALID, "FRI final layer too large (2^%d values)", F.log_n);
  const size_t fin_n = (size_t)1 << F.log_n, final_len = fin_n >> s->p.rate_bits;
  std::vector<uint64_t> fin_host(2 * fin_n);
  ETP_CUDA(ctx, cudaMemcpyAsync(fin_host.data(), F.values, fin_host.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  std::vector<gl::Ext> fin(fin_n);
  for (size_t p = 0; p < fin_n; p++) {
    const size_t k = gl::bitrev32((uint32_t)p, F.log_n);
    fin[k] = gl::ext(fin_host[2 * p], fin_host[2 * p + 1]);
  }
  hostf::ext_fft(fin, F.log_n, true);
  {
    const uint64_t si = gl::canon(gl::inv(s->shift));
    uint64_t cur = 1;
    for (size_t k = 0; k < fin_n; k++) { fin[k] = gl::ecanon(gl::emul_base(fin[k], cur)); cur = gl::mul(cur, si); }
  }
  for (size_t k = final_len; k < fin_n; k++)
    if (fin[k].c0 || fin[k].c1)
      return etp_fail(ctx, ETP_ERR_PROOF, "FRI final polynomial has degree >= %zu: the committed function is not a low-degree polynomial (trace violates the constraints)", final_len);
  final_coeffs->resize(2 * final_len);
  for (size_t k = 0; k < final_len; k++) { (*final_coeffs)[2 * k] = fin[k].c0; (*final_coeffs)[2 * k + 1] = fin[k].c1; }
  return ETP_OK;
}
// fri_committed_trees: all layers through the challenger, then observe the final polynomial
int fri_commit_phase(etp_fri_state* s, hostf::Challenger& ch, uint64_t* caps_out, std::vector<uint64_t>* final_coeffs) {
  const size_t cap_words = (size_t)4 << s->p.cap_height;
  for (int l = 0; l < s->p.n_reductions; l++) {
    ETP_TRY(fri_commit_layer(s, nullptr));
    const FriLayer& L = s->layers[l];
    ch.observe(L.cap.data(), cap_words);
    if (caps_out) memcpy(caps_out + (size_t)l * cap_words, L.cap.data(), cap_words * 8);
    ETP_TRY(fri_fold(s, ch.get_ext()));
  }
  ETP_TRY(fri_final_poly(s, final_coeffs));
  ch.observe(final_coeffs->data(), final_coeffs->size());
  return ETP_OK;
}
// fri_prover_query_rounds for the given x indices: out receives, per index, the initial-tree rows + paths and the per-layer
// evaluations + paths
int fri_query_rounds(etp_fri_state* s, etp_batch* const* oracles, size_t n_oracles, const uint64_t* x_idx, size_t nq, uint64_t* out) {
  etp_ctx* ctx = s->ctx;
  const etp_fri_params& p = s->p;
  const int log_lde = p.degree_bits + p.rate_bits, n_layers = p.n_reductions;
  const size_t lde_n = (size_t)1 << log_lde;
  if (s->committed != n_layers) return etp_fail(ctx, ETP_ERR_STATE, "fri_query_rounds: the commit phase is not finished");
  if (nq == 0) return ETP_OK;
  const int init_path = log_lde - p.cap_height;
  for (size_t o = 0; o < n_oracles; o++)
    if (!oracles[o] || oracles[o]->ctx != ctx || oracles[o]->log_n != p.degree_bits || oracles[o]->rate_bits != p.rate_bits || oracles[o]->cap_height != p.cap_height)
      return etp_fail(ctx, ETP_ERR_INVALID, "FRI oracle %zu does not match the FRI parameters", o);
  // device staging: per oracle rows + paths, per layer rows + paths
  // (gather_paths stores digests as two 16-byte words: every path section starts on an even word)
  size_t stage_words = 0;
  for (size_t o = 0; o < n_oracles; o++) stage_words += nq * (oracles[o]->n_cols + 4 * (size_t)init_path) + 1;
  {
    int bits = log_lde;
    for (int l = 0; l < n_layers; l++) { bits -= ARITY_BITS; stage_words += nq * (32 + 4 * (size_t)(bits - p.cap_height)); }
  }
  DevBuf<uint64_t> stage(ctx), d_idx(ctx);
  ETP_TRY(stage.alloc(stage_words));
  ETP_TRY(d_idx.alloc(nq * (n_layers + 1)));
  std::vector<uint64_t> idx_all(nq * (n_layers + 1));
  for (size_t qn = 0; qn < nq; qn++) {
    uint64_t x = x_idx[qn];
    if (x >= lde_n) return etp_fail(ctx, ETP_ERR_INVALID, "FRI query index out of range");
    idx_all[qn] = x;
    for (int l = 0; l < n_layers; l++) { x >>= ARITY_BITS; idx_all[(size_t)(l + 1) * nq + qn] = x; }
  }
  ETP_CUDA(ctx, cudaMemcpyAsync(d_idx.p, idx_all.data(), idx_all.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  std::vector<size_t> off_rows(n_oracles), off_paths(n_oracles);
  size_t off = 0;
  for (size_t o = 0; o < n_oracles; o++) {
    const int nc = (int)oracles[o]->n_cols;
    off_rows[o] = off;
    if (nc) {
      merkle::gather_rows_colmajor<<<blocks_for(nq * nc, 256), 256, 0, ctx->stream>>>(oracles[o]->lde, lde_n, nc, d_idx.p, (int)nq, stage.p + off);
      ETP_LAUNCH_CHECK(ctx);
    }
    off += nq * nc;
    off += off & 1;
    off_paths[o] = off;
    if (init_path > 0) {
      stark::gather_paths<<<blocks_for(nq * init_path, 256), 256, 0, ctx->stream>>>(oracles[o]->levels, (uint32_t)lde_n, init_path, d_idx.p, (int)nq,
                                                                                   stage.p + off);
      ETP_LAUNCH_CHECK(ctx);
    }
    off += nq * 4 * init_path;
  }
  std::vector<size_t> loff_rows(n_layers), loff_paths(n_layers);
  for (int l = 0; l < n_layers; l++) {
    const FriLayer& L = s->layers[l];
    const int path = L.log_n - ARITY_BITS - p.cap_height;
    loff_rows[l] = off;
    stark::gather_rows_rowmajor<<<blocks_for(nq * 32, 256), 256, 0, ctx->stream>>>(L.values, 32, d_idx.p + (size_t)(l + 1) * nq, (int)nq, stage.p + off);
    ETP_LAUNCH_CHECK(ctx);
    off += nq * 32;
    loff_paths[l] = off;
    if (path > 0) {
      stark::gather_paths<<<blocks_for(nq * path, 256), 256, 0, ctx->stream>>>(L.levels, (uint32_t)L.n_leaves, path, d_idx.p + (size_t)(l + 1) * nq,
                                                                              (int)nq, stage.p + off);
      ETP_LAUNCH_CHECK(ctx);
    }
    off += nq * 4 * path;
  }
  std::vector<uint64_t> sh(stage_words ? stage_words : 1);
  ETP_CUDA(ctx, cudaMemcpyAsync(sh.data(), stage.p, stage_words * 8, cudaMemcpyDeviceToHost, ctx->stream));
  ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  uint64_t* w = out;
  for (size_t qn = 0; qn < nq; qn++) {
    for (size_t o = 0; o < n_oracles; o++) {
      const size_t nc = oracles[o]->n_cols;
      memcpy(w, &sh[off_rows[o] + qn * nc], nc * 8); w += nc;
      memcpy(w, &sh[off_paths[o] + qn * 4 * init_path], (size_t)4 * init_path * 8); w += 4 * init_path;
    }
    for (int l = 0; l < n_layers; l++) {
      const int path = s->layers[l].log_n - ARITY_BITS - p.cap_height;
      memcpy(w, &sh[loff_rows[l] + qn * 32], 32 * 8); w += 32;
      memcpy(w, &sh[loff_paths[l] + qn * 4 * path], (size_t)4 * path * 8); w += 4 * path;
    }
  }
  return ETP_OK;
}

// fri_proof_of_work: the smallest witness whose response has `bits` leading zeros; the challenger observes the witness
// and the response is drawn from it (what the verifier's challenger does).
int fri_proof_of_work(etp_ctx* ctx, hostf::Challenger& ch, int bits, uint64_t* witness) {
  uint64_t st[12];
  memcpy(st, ch.sponge_state, sizeof st);
  for (uint32_t i = 0; i < ch.input_len; i++) st[i] = ch.input_buffer[i];
  ETP_TRY(pow_grind(ctx, st, (int)ch.input_len, bits, witness));
  ch.observe(*witness);
  const uint64_t pow_response = ch.get();
  if (bits && (pow_response >> (64 - bits)) != 0) return etp_fail(ctx, ETP_ERR_PROOF, "proof of work response mismatch");
  return ETP_OK;
}

// fri_proof: commit phase, proof of work, query rounds.  out: flat FriProof (fri_proof_words)
int fri_proof(etp_fri_state* s, etp_batch* const* oracles, size_t n_oracles, hostf::Challenger& ch, uint64_t* out, PhaseTimer* timer) {
  etp_ctx* ctx = s->ctx;
  const etp_fri_params& p = s->p;
  const size_t cap_words = (size_t)4 << p.cap_height;
  const size_t lde_n = (size_t)1 << (p.degree_bits + p.rate_bits);
  uint64_t* w = out;
  std::vector<uint64_t> final_coeffs;
  ETP_TRY(fri_commit_phase(s, ch, w, &final_coeffs));
  w += cap_words * p.n_reductions;
  if (timer) timer->mark("fold codewords in the commitment phase");
  uint64_t pow_witness = 0;
  ETP_TRY(fri_proof_of_work(ctx, ch, p.proof_of_work_bits, &pow_witness));
  if (timer) timer->mark("find proof-of-work witness");
  // fri_prover_query_rounds
  std::vector<uint64_t> qidx(p.num_query_rounds);
  for (int qn = 0; qn < p.num_query_rounds; qn++) qidx[qn] = ch.get() % lde_n;
  ETP_TRY(fri_query_rounds(s, oracles, n_oracles, qidx.data(), qidx.size(), w));
  std::vector<size_t> oc(n_oracles);
  for (size_t o = 0; o < n_oracles; o++) oc[o] = oracles[o]->n_cols;
  w = out + fri_proof_words(oc.data(), n_oracles, p) - final_coeffs.size() - 1;
  if (timer) timer->mark("build FRI query rounds");
  memcpy(w, final_coeffs.data(), final_coeffs.size() * 8); w += final_coeffs.size();
  *w++ = pow_witness;
  return ETP_OK;
}

template <int B>
void launch_combine(etp_ctx* ctx, const stark::CombineParams& c, size_t lde_n, bool chunked) {
  if (chunked) stark::combine_accumulate<B><<<dim3(blocks_for(lde_n, 128), c.n_chunks), 128, 0, ctx->stream>>>(c);
  else stark::combine_values<B><<<blocks_for(lde_n, 128), 128, 0, ctx->stream>>>(c);
}

// The combination step of prove_openings on the LDE coset: values[p] = sum_b alpha^(shift_b) (sum_k alpha^k f_{b,k}(x_p) - y_b) / (x_p - z_b)
// (ext, interleaved, bit-reversed order).  `col(oracle, poly)` resolves a FriPolynomialInfo to the device pointer of that
// polynomial's LDE column (bit-reversed rows, 2^(log_n + rate_bits) words) or nullptr — a PolynomialBatch of this context,
// or a column of a peer GPU's shard mapped over NVLink (column-split tables).
typedef std::function<const uint64_t*(uint32_t, uint32_t)> ColResolver;
int combine_on_lde(etp_ctx* ctx, const etp_fri_batch* batches, size_t n_batches, const ColResolver& col, int log_n, int rate_bits, gl::Ext alpha,
                   const std::vector<std::vector<gl::Ext>>& ys, uint64_t* values) {
  if (n_batches < 1 || n_batches > (size_t)stark::MAX_FRI_BATCHES) return etp_fail(ctx, ETP_ERR_INVALID, "between 1 and %d FRI batches are supported", stark::MAX_FRI_BATCHES);
  const int log_lde = log_n + rate_bits;
  const size_t lde_n = (size_t)1 << log_lde;
  HostTrace ht("combine_on_lde");
  // unique columns in order of first appearance, with their alpha power per batch
  std::vector<stark::CombineCol> cols;
  std::map<std::pair<uint32_t, uint32_t>, int> where;
  size_t max_count = 0;
  for (size_t b = 0; b < n_batches; b++) {
    max_count = batches[b].n_polynomials > max_count ? batches[b].n_polynomials : max_count;
    for (size_t k = 0; k < batches[b].n_polynomials; k++) {
      const etp_fri_poly fpi = batches[b].polynomials[k];
      auto key = std::make_pair(fpi.oracle_index, fpi.polynomial_index);
      auto it = where.find(key);
      int u;
      if (it == where.end()) {
        stark::CombineCol cc{};
        cc.ptr = col(fpi.oracle_index, fpi.polynomial_index);
        if (!cc.ptr) return etp_fail(ctx, ETP_ERR_INVALID, "FRI batch %zu: polynomial (%u, %u) does not exist", b, fpi.oracle_index, fpi.polynomial_index);
        for (int j = 0; j < stark::MAX_FRI_BATCHES; j++) cc.idx[j] = stark::COMBINE_NONE;
        cols.push_back(cc);
        u = (int)cols.size() - 1;
        where.emplace(key, u);
      } else {
        u = it->second;
      }
      if (cols[u].idx[b] != stark::COMBINE_NONE) return etp_fail(ctx, ETP_ERR_INVALID, "FRI batch %zu lists a polynomial twice", b);
      cols[u].idx[b] = (uint32_t)k;
    }
  }
  ht.mark("unique columns");
  stark::CombineParams c{};
  c.n_cols = (int)cols.size(); c.n_batches = (int)n_batches; c.log_lde = log_lde;
  for (size_t b = 1; b < n_batches; b++) {  // prefix batches: unique column u carries power u in batch 0 and in batch b, for u < n_b
    bool prefix = batches[b].n_polynomials > 0 && batches[b].n_polynomials <= batches[0].n_polynomials;
    for (size_t u = 0; prefix && u < cols.size(); u++)
      prefix = u < batches[b].n_polynomials ? (cols[u].idx[b] == u && cols[u].idx[0] == u) : cols[u].idx[b] == stark::COMBINE_NONE;
    c.prefix_len[b] = prefix ? (int)batches[b].n_polynomials : 0;
  }
  // alpha powers, the reduced openings y_b = sum_k alpha^k f_k(z_b) and the batch shifts
  std::vector<uint64_t> apow(2 * (max_count + 1));
  std::vector<gl::Ext> apow_e(max_count + 1);
  {
    gl::Ext cur = gl::ext(1, 0);
    for (size_t k = 0; k <= max_count; k++) { cur = gl::ecanon(cur); apow_e[k] = cur; apow[2 * k] = cur.c0; apow[2 * k + 1] = cur.c1; cur = gl::emul(cur, alpha); }
  }
  {
    size_t later = 0;
    for (size_t b = n_batches; b-- > 0;) {
      c.z[b] = gl::ecanon(gl::ext(batches[b].point[0], batches[b].point[1]));
      c.seven_zc1_sq[b] = gl::canon(gl::mul(7, gl::mul(c.z[b].c1, c.z[b].c1)));
      gl::Ext y = gl::ext(0, 0);
      if (ys[b].size() != batches[b].n_polynomials) return etp_fail(ctx, ETP_ERR_INVALID, "opening count of batch %zu does not match its polynomial list", b);
      for (size_t k = 0; k < batches[b].n_polynomials; k++) y = gl::eadd(y, gl::emul(apow_e[k], ys[b][k]));
      c.y[b] = gl::ecanon(y);
      c.shift[b] = gl::ecanon(gl::epow(alpha, later));
      later += batches[b].n_polynomials;
    }
  }
  ht.mark("alpha powers, reduced openings");
  DevBuf<uint64_t> d_apow(ctx), den(ctx), partial(ctx);
  DevBuf<stark::CombineCol> d_cols(ctx);
  ETP_TRY(d_apow.alloc(apow.size()));
  ETP_TRY(den.alloc(n_batches * lde_n));
  ETP_TRY(d_cols.alloc(cols.size() ? cols.size() : 1));
  ETP_CUDA(ctx, cudaMemcpyAsync(d_apow.p, apow.data(), apow.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  ETP_CUDA(ctx, cudaMemcpyAsync(d_cols.p, cols.data(), cols.size() * sizeof(stark::CombineCol), cudaMemcpyHostToDevice, ctx->stream));
  ETP_TRY(get_pow_table(ctx, gl::root_of_unity(log_lde), log_lde, gl::GENERATOR, &c.coset));
  c.cols = d_cols.p; c.alpha_pows = d_apow.p; c.den = den.p; c.out = values;
  ht.mark("buffers, uploads, tables");
  ETP_TRY(reset_zero_flag(ctx));
  stark::combine_norms<<<blocks_for(lde_n, 256), 256, 0, ctx->stream>>>(c);
  ETP_LAUNCH_CHECK(ctx);
  ETP_TRY(batch_inverse_dev(ctx, den.p, den.p, n_batches * lde_n));
  // few rows, many columns: one thread per point cannot fill the machine, so the column sums are split into chunks
  const int B = n_batches <= 2 ? 2 : (n_batches == 3 ? 3 : 4);
  if (c.n_cols > 256 && lde_n * 4 <= (size_t)1 << 20) {
    c.col_chunk = 64;
    c.n_chunks = (c.n_cols + c.col_chunk - 1) / c.col_chunk;
    ETP_TRY(partial.alloc((size_t)c.n_chunks * lde_n * 2 * B));
    c.partial = partial.p;
    if (B == 2) launch_combine<2>(ctx, c, lde_n, true); else if (B == 3) launch_combine<3>(ctx, c, lde_n, true); else launch_combine<4>(ctx, c, lde_n, true);
    ETP_LAUNCH_CHECK(ctx);
  }
  if (B == 2) launch_combine<2>(ctx, c, lde_n, false); else if (B == 3) launch_combine<3>(ctx, c, lde_n, false); else launch_combine<4>(ctx, c, lde_n, false);
  ETP_LAUNCH_CHECK(ctx);
  ht.mark("launches");
  const int rc_zero = check_zero_flag(ctx, "an opening point lies on the LDE coset");
  ht.mark("kernels (sync)");
  return rc_zero;  // also: apow / cols (host) stay alive until here
}

// PolynomialBatch::prove_openings in evaluation form over the LDE coset, then fri_proof.  `ys`: the claimed value of every
// batch polynomial at the batch point when the caller already has them (the openings of a STARK), else nullptr (they are
// then evaluated here from the coefficients).
int prove_openings(etp_ctx* ctx, const etp_fri_batch* batches, size_t n_batches, etp_batch* const* oracles, size_t n_oracles,
                   hostf::Challenger& ch, const etp_fri_params& fp, const std::vector<std::vector<gl::Ext>>* ys, uint64_t* out, PhaseTimer* timer) {
  ETP_TRY(check_fri_params(ctx, fp));
  if (n_batches < 1 || n_batches > (size_t)stark::MAX_FRI_BATCHES) return etp_fail(ctx, ETP_ERR_INVALID, "between 1 and %d FRI batches are supported", stark::MAX_FRI_BATCHES);
  const int log_n = fp.degree_bits, log_lde = log_n + fp.rate_bits;
  const size_t lde_n = (size_t)1 << log_lde;
  for (size_t o = 0; o < n_oracles; o++)
    if (!oracles[o] || oracles[o]->ctx != ctx || oracles[o]->log_n != log_n || oracles[o]->rate_bits != fp.rate_bits || oracles[o]->cap_height != fp.cap_height)
      return etp_fail(ctx, ETP_ERR_INVALID, "FRI oracle %zu does not match the FRI parameters", o);
  for (size_t b = 0; b < n_batches; b++)
    for (size_t k = 0; k < batches[b].n_polynomials; k++) {
      const etp_fri_poly fpi = batches[b].polynomials[k];
      if (fpi.oracle_index >= n_oracles || fpi.polynomial_index >= oracles[fpi.oracle_index]->n_cols)
        return etp_fail(ctx, ETP_ERR_INVALID, "FRI batch %zu: polynomial (%u, %u) does not exist", b, fpi.oracle_index, fpi.polynomial_index);
    }
  const gl::Ext alpha = ch.get_ext();
  std::vector<std::vector<gl::Ext>> ys_local;
  if (!ys) {  // evaluate f_k(z_b) from the coefficients: one kernel per (oracle, pair of points)
    ys_local.resize(n_batches);
    for (size_t b = 0; b < n_batches; b += 2) {
      const gl::Ext z0 = gl::ecanon(gl::ext(batches[b].point[0], batches[b].point[1]));
      const gl::Ext z1 = b + 1 < n_batches ? gl::ecanon(gl::ext(batches[b + 1].point[0], batches[b + 1].point[1])) : z0;
      DevBuf<uint64_t> l0(ctx), h0(ctx), l1(ctx), h1(ctx);
      stark::ExtPowTable t0, t1;
      ETP_TRY(ext_pow_table(ctx, z0, log_n, l0, h0, &t0));
      ETP_TRY(ext_pow_table(ctx, z1, log_n, l1, h1, &t1));
      std::vector<std::vector<gl::Ext>> e0(n_oracles), e1(n_oracles);
      std::vector<bool> done(n_oracles, false);
      for (size_t bb = b; bb < b + 2 && bb < n_batches; bb++) {
        ys_local[bb].resize(batches[bb].n_polynomials);
        for (size_t k = 0; k < batches[bb].n_polynomials; k++) {
          const etp_fri_poly fpi = batches[bb].polynomials[k];
          if (!done[fpi.oracle_index]) { ETP_TRY(eval_batch(ctx, oracles[fpi.oracle_index], t0, t1, e0[fpi.oracle_index], e1[fpi.oracle_index])); done[fpi.oracle_index] = true; }
          ys_local[bb][k] = (bb == b ? e0 : e1)[fpi.oracle_index][fpi.polynomial_index];
        }
      }
    }
    ys = &ys_local;
  }
  uint64_t* values = nullptr;
  ETP_TRY(dev_alloc(ctx, 2 * lde_n * 8, (void**)&values));
  etp_fri_state* st = nullptr;
  ETP_TRY(fri_begin_owned(ctx, values, fp, &st));
  std::unique_ptr<etp_fri_state> guard(st);
  const ColResolver col = [&](uint32_t o, uint32_t k) -> const uint64_t* { return oracles[o]->lde + (size_t)k * lde_n; };
  ETP_TRY(combine_on_lde(ctx, batches, n_batches, col, log_n, fp.rate_bits, alpha, *ys, values));
  if (timer) timer->mark("compute openings proof: combine on the LDE domain");
  return fri_proof(st, oracles, n_oracles, ch, out, timer);
}

// starky::prover::prove_with_commitment.  trace_dev: the trace values (needed for the auxiliary columns).
int prove_with_commitment(etp_ctx* ctx, int table, const TableInfo& ti, etp_batch* trace, const uint64_t* trace_dev, size_t stride,
                          const uint64_t* ctl_ch, hostf::Challenger& ch, const uint64_t* pi_in, uint64_t* proof, PhaseTimer& timer) {
  const int log_n = trace->log_n;
  const etp_fri_params fp = standard_fast_params(log_n);
  ETP_TRY(check_fri_params(ctx, fp));
  if ((int)trace->n_cols != ti.cols || trace->rate_bits != RATE_BITS || trace->cap_height != CAP_HEIGHT || trace->ctx != ctx)
    return etp_fail(ctx, ETP_ERR_INVALID, "trace commitment does not match the table / standard_fast_config");
  if (ti.ctl() && !ctl_ch) return etp_fail(ctx, ETP_ERR_INVALID, "the table requires CTLs but no CTL challenges were given");
  const size_t n = (size_t)1 << log_n;
  const size_t cap_words = (size_t)4 << CAP_HEIGHT;
  const int n_lookup = ti.n_lookup_cols(NUM_CHALLENGES), n_helpers = ti.n_ctl_helpers(), n_zs = ti.n_ctl_zs();
  const int n_aux = n_lookup + n_helpers + n_zs, factor = quotient_factor(ti), n_quot = factor * NUM_CHALLENGES;
  uint64_t pi[stark::MAX_PUBLIC_INPUTS] = {};
  for (int i = 0; i < ti.n_pi; i++) pi[i] = gl::canon(pi_in[i]);

  uint64_t* w = proof;
  uint64_t* hdr = w; w += HEADER_WORDS;
  memset(hdr, 0, HEADER_WORDS * 8);
  hdr[0] = PROOF_MAGIC; hdr[1] = table; hdr[2] = log_n; hdr[3] = ti.cols; hdr[4] = n_aux; hdr[5] = n_quot; hdr[6] = CAP_HEIGHT;
  hdr[7] = fp.n_reductions; hdr[8] = ARITY_BITS; hdr[9] = (uint64_t)1 << (log_n - fri_total_arities(fp)); hdr[10] = NUM_QUERIES; hdr[11] = ti.n_pi;
  hdr[12] = RATE_BITS; hdr[13] = POW_BITS; hdr[14] = NUM_CHALLENGES; hdr[15] = proof_words(ti, log_n);
  hdr[16] = n_zs; hdr[17] = n_lookup; hdr[18] = n_helpers;
  memcpy(w, trace->cap.data(), cap_words * 8); w += cap_words;

  struct Batches {
    etp_batch *aux = nullptr, *quot = nullptr;
    ~Batches() { etp_batch_free(aux); etp_batch_free(quot); }
  } B;

  // ---- lookup challenges: the CTL betas when CTL challenges are given, else get_grand_product_challenge_set's betas
  uint64_t scalars[stark::MAX_CH_SCALARS] = {};
  if (ti.lookup()) {
    for (int k = 0; k < NUM_CHALLENGES; k++) {
      if (ctl_ch) scalars[k] = gl::canon(ctl_ch[2 * k]);
      else { scalars[k] = ch.get(); (void)ch.get(); }
    }
  }
  if (ctl_ch) for (int k = 0; k < 2 * NUM_CHALLENGES; k++) scalars[NUM_CHALLENGES + k] = gl::canon(ctl_ch[k]);
  const int n_scalars = ctl_ch ? 3 * NUM_CHALLENGES : (ti.lookup() ? NUM_CHALLENGES : 0);
  std::vector<uint64_t> zs_first(n_zs);
  if (n_aux) {
    DevBuf<uint64_t> aux_vals(ctx);
    ETP_TRY(aux_vals.alloc((size_t)n_aux * n));
    ETP_TRY(reset_zero_flag(ctx));
    ETP_TRY(aux_columns(ctx, ti, log_n, trace_dev, stride, scalars, NUM_CHALLENGES, ctl_ch ? scalars + NUM_CHALLENGES : nullptr, aux_vals.p));
    // ctl_zs_first: Z(1) = the first trace-domain value of every CTL Z
    if (n_zs)
      ETP_CUDA(ctx, cudaMemcpy2DAsync(zs_first.data(), 8, aux_vals.p + (size_t)(n_lookup + n_helpers) * n, n * 8, 8, n_zs, cudaMemcpyDeviceToHost,
                                      ctx->stream));
    timer.mark("compute lookup helper columns");
    ETP_TRY(batch_create(ctx, n_aux, log_n, RATE_BITS, 0, CAP_HEIGHT, &B.aux));
    ETP_TRY(batch_commit_from_values(B.aux, aux_vals.p, n));  // synchronises the stream
    ETP_TRY(check_zero_flag(ctx, "a lookup / CTL denominator vanishes on the trace"));
    timer.mark("auxiliary polys commit");
    ch.observe(B.aux->cap.data(), cap_words);
    memcpy(w, B.aux->cap.data(), cap_words * 8); w += cap_words;
  }
  uint64_t alphas[NUM_CHALLENGES];
  for (int j = 0; j < NUM_CHALLENGES; j++) alphas[j] = ch.get();

  // ---- quotient
  ETP_TRY(batch_create(ctx, n_quot, log_n, RATE_BITS, 0, CAP_HEIGHT, &B.quot));
  ETP_TRY(compute_quotient(ctx, table, trace, B.aux, scalars, n_scalars, pi, alphas, NUM_CHALLENGES, B.quot->coeffs));
  timer.mark("compute quotient polys");
  ETP_TRY(batch_commit_from_coeffs(B.quot));
  timer.mark("quotient polys commit");
  ch.observe(B.quot->cap.data(), cap_words);
  memcpy(w, B.quot->cap.data(), cap_words * 8); w += cap_words;

  // ---- openings
  const gl::Ext zeta = ch.get_ext();
  const uint64_t g = gl::root_of_unity(log_n);
  {
    gl::Ext zp = zeta;
    for (int i = 0; i < log_n; i++) zp = gl::emul(zp, zp);
    zp = gl::ecanon(zp);
    if (zp.c0 == 1 && zp.c1 == 0) return etp_fail(ctx, ETP_ERR_PROOF, "Opening point is in the subgroup.");
  }
  const gl::Ext zeta_next = gl::ecanon(gl::emul_base(zeta, g));
  std::vector<gl::Ext> tr0, tr1, ax0, ax1, qu0, qu1;
  {
    DevBuf<uint64_t> l0(ctx), h0(ctx), l1(ctx), h1(ctx);
    stark::ExtPowTable t0, t1;
    ETP_TRY(ext_pow_table(ctx, zeta, log_n, l0, h0, &t0));
    ETP_TRY(ext_pow_table(ctx, zeta_next, log_n, l1, h1, &t1));
    ETP_TRY(eval_batch(ctx, trace, t0, t1, tr0, tr1));
    if (B.aux) ETP_TRY(eval_batch(ctx, B.aux, t0, t1, ax0, ax1));
    ETP_TRY(eval_batch(ctx, B.quot, t0, t1, qu0, qu1));
  }
  timer.mark("compute openings proof: evaluate at zeta, g*zeta");
  // StarkOpeningSet order: local_values, next_values, auxiliary_polys, auxiliary_polys_next, ctl_zs_first, quotient_polys
  auto put = [&](const std::vector<gl::Ext>& v) { for (auto& e : v) { *w++ = e.c0; *w++ = e.c1; } };
  put(tr0); put(tr1); put(ax0); put(ax1);
  for (int k = 0; k < n_zs; k++) *w++ = gl::canon(zs_first[k]);
  put(qu0);
  // observe_openings(to_fri_openings): zeta batch = local ++ aux ++ quotient ; next batch = next ++ aux_next ; ctl_zs_first
  auto obs = [&](const std::vector<gl::Ext>& v) { for (auto& e : v) { ch.observe(e.c0); ch.observe(e.c1); } };
  obs(tr0); obs(ax0); obs(qu0); obs(tr1); obs(ax1);
  for (int k = 0; k < n_zs; k++) { ch.observe(zs_first[k]); ch.observe(0); }

  // ---- stark.fri_instance(zeta, g, num_ctl_helpers, num_ctl_zs, config) + prove_openings
  etp_batch* oracles[3];
  size_t n_oracles = 0;
  const uint32_t o_trace = (uint32_t)n_oracles; oracles[n_oracles++] = trace;
  const uint32_t o_aux = (uint32_t)n_oracles; if (B.aux) oracles[n_oracles++] = B.aux;
  const uint32_t o_quot = (uint32_t)n_oracles; oracles[n_oracles++] = B.quot;
  std::vector<etp_fri_poly> p0, p1, p2;
  std::vector<std::vector<gl::Ext>> ys(3);
  for (int cidx = 0; cidx < ti.cols; cidx++) { p0.push_back({o_trace, (uint32_t)cidx}); p1.push_back({o_trace, (uint32_t)cidx}); }
  for (int cidx = 0; cidx < n_aux; cidx++) { p0.push_back({o_aux, (uint32_t)cidx}); p1.push_back({o_aux, (uint32_t)cidx}); }
  for (int cidx = 0; cidx < n_quot; cidx++) p0.push_back({o_quot, (uint32_t)cidx});
  for (int k = 0; k < n_zs; k++) p2.push_back({o_aux, (uint32_t)(n_lookup + n_helpers + k)});
  ys[0] = tr0; ys[0].insert(ys[0].end(), ax0.begin(), ax0.end()); ys[0].insert(ys[0].end(), qu0.begin(), qu0.end());
  ys[1] = tr1; ys[1].insert(ys[1].end(), ax1.begin(), ax1.end());
  for (int k = 0; k < n_zs; k++) ys[2].push_back(gl::ext(gl::canon(zs_first[k]), 0));
  etp_fri_batch batches[3];
  batches[0] = {{zeta.c0, zeta.c1}, p0.data(), p0.size()};
  batches[1] = {{zeta_next.c0, zeta_next.c1}, p1.data(), p1.size()};
  batches[2] = {{1, 0}, p2.data(), p2.size()};
  const size_t n_batches = n_zs ? 3 : 2;
  ys.resize(n_batches);
  ETP_TRY(prove_openings(ctx, batches, n_batches, oracles, n_oracles, ch, fp, &ys, w, &timer));
  size_t oc[3];
  for (size_t o = 0; o < n_oracles; o++) oc[o] = oracles[o]->n_cols;
  w += fri_proof_words(oc, n_oracles, fp);
  for (int i = 0; i < ti.n_pi; i++) *w++ = pi[i];
  if ((size_t)(w - proof) != hdr[15]) return etp_fail(ctx, ETP_ERR_STATE, "internal error: proof size mismatch");
  return ETP_OK;
}

// starky::prover::prove.  trace_host != nullptr: the trace (n_cols x n, column-major, stride n) is still on the host;
// trace_dev is an empty device buffer of the same shape that the streamed trace commit fills column group by column group
// while it transforms and hashes the groups that have already arrived (etp_stark_prove_host).
int stark_prove_dev(etp_ctx* ctx, int table, int log_n, const uint64_t* trace_dev, size_t stride, const uint64_t* pi_in, uint64_t* proof,
                    const uint64_t* trace_host = nullptr) {
  TableInfo ti;
  if (!table_info(ctx, table, &ti)) return etp_fail(ctx, ETP_ERR_INVALID, "unknown table %d", table);
  if (log_n < 1 || log_n + RATE_BITS > 30) return etp_fail(ctx, ETP_ERR_INVALID, "unsupported degree_bits %d", log_n);
  if (ti.ctl()) return etp_fail(ctx, ETP_ERR_INVALID, "the table requires CTLs: use etp_prove_with_commitment with the CTL challenges");
  ETP_TRY(check_fri_params(ctx, standard_fast_params(log_n)));
  const size_t n = (size_t)1 << log_n;
  PhaseTimer timer(ctx);
  struct Guard { etp_batch* b = nullptr; ~Guard() { etp_batch_free(b); } } T;
  // ---- prove(): trace commitment; the challenger observes the public inputs, then the trace cap
  ETP_TRY(batch_create(ctx, ti.cols, log_n, RATE_BITS, 0, CAP_HEIGHT, &T.b));
  if (trace_host) {
    std::vector<const uint64_t*> cols(ti.cols);
    for (int c = 0; c < ti.cols; c++) cols[c] = trace_host + (size_t)c * n;
    ETP_TRY(batch_commit_from_host_streamed(T.b, cols.data(), true, const_cast<uint64_t*>(trace_dev)));
  } else {
    ETP_TRY(batch_commit_from_values(T.b, trace_dev, stride));
  }
  timer.mark("trace commit (IFFT + FFT + Merkle tree)");
  hostf::Challenger ch;
  uint64_t pi[stark::MAX_PUBLIC_INPUTS] = {};
  for (int i = 0; i < ti.n_pi; i++) pi[i] = gl::canon(pi_in[i]);
  ch.observe(pi, ti.n_pi);
  ch.observe(T.b->cap.data(), (size_t)4 << CAP_HEIGHT);
  ETP_TRY(prove_with_commitment(ctx, table, ti, T.b, trace_dev, stride, nullptr, ch, pi, proof, timer));
  timer.finish();
  return ETP_OK;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" void etp_challenger_init(etp_challenger* c) { if (c) memset(c, 0, sizeof *c); }
extern "C" void etp_challenger_observe(etp_challenger* c, const uint64_t* e, size_t n) {
  if (!c || (!e && n)) return;
  hostf::Challenger ch(*c);
  ch.observe(e, n);
  *c = ch;
}
extern "C" uint64_t etp_challenger_get_challenge(etp_challenger* c) {
  if (!c) return 0;
  hostf::Challenger ch(*c);
  const uint64_t v = ch.get();
  *c = ch;
  return v;
}
extern "C" void etp_challenger_get_n_challenges(etp_challenger* c, size_t n, uint64_t* out) {
  if (!c || (!out && n)) return;
  hostf::Challenger ch(*c);
  for (size_t i = 0; i < n; i++) out[i] = ch.get();
  *c = ch;
}
extern "C" void etp_challenger_compact(etp_challenger* c) {
  if (!c) return;
  hostf::Challenger ch(*c);
  ch.compact();
  *c = ch;
}
static bool challenger_ok(const etp_challenger* c) { return c && c->input_len < 8 && c->output_len <= 8; }

extern "C" int etp_fri_params_make(int degree_bits, int rate_bits, int cap_height, int pow_bits, int num_queries, etp_fri_params* out) {
  if (!out || degree_bits < 0 || rate_bits < 0 || degree_bits + rate_bits > 30 || cap_height < 0) return ETP_ERR_INVALID;
  fri_params_make(degree_bits, rate_bits, cap_height, pow_bits, num_queries, out);
  return ETP_OK;
}

extern "C" int etp_table_num_columns(const etp_ctx* c, int t) { TableInfo ti; return table_info(c, t, &ti) ? ti.cols : -1; }
extern "C" int etp_table_constraint_degree(const etp_ctx* c, int t) { TableInfo ti; return table_info(c, t, &ti) ? ti.degree : -1; }
extern "C" int etp_table_num_public_inputs(const etp_ctx* c, int t) { TableInfo ti; return table_info(c, t, &ti) ? ti.n_pi : -1; }
extern "C" int etp_table_num_aux_columns(const etp_ctx* c, int t, int nc) { TableInfo ti; return table_info(c, t, &ti) ? ti.n_aux(nc) : -1; }
extern "C" int etp_table_num_lookup_columns(const etp_ctx* c, int t, int nc) { TableInfo ti; return table_info(c, t, &ti) ? ti.n_lookup_cols(nc) : -1; }
extern "C" int etp_table_num_ctl_helper_columns(const etp_ctx* c, int t) { TableInfo ti; return table_info(c, t, &ti) ? ti.n_ctl_helpers() : -1; }
extern "C" int etp_table_num_ctl_zs(const etp_ctx* c, int t) { TableInfo ti; return table_info(c, t, &ti) ? ti.n_ctl_zs() : -1; }
extern "C" int etp_table_quotient_degree_factor(const etp_ctx* c, int t) { TableInfo ti; return table_info(c, t, &ti) ? quotient_factor(ti) : -1; }

static int register_table(etp_ctx* ctx, const uint64_t* program, size_t n_words, const std::vector<uint64_t>& spec_words, int* table_id_out,
                          bool standalone = false) {
  auto t = new RegisteredTable();
  struct Guard { RegisteredTable* t; ~Guard() { if (t) { jit_unload(&t->kernel); jit_unload(&t->kernel_split); delete t; } } } guard{t};
  const std::string why = cprog::parse(program, n_words, stark::MAX_PUBLIC_INPUTS, stark::MAX_CH_SCALARS, &t->prog);
  if (!why.empty()) return etp_fail(ctx, ETP_ERR_INVALID, "%s", why.c_str());
  TableInfo& ti = t->info;
  ti.cols = (int)t->prog.n_trace; ti.degree = (int)t->prog.degree; ti.n_pi = (int)t->prog.n_pi; ti.reg = t;
  ti.aux = std::make_shared<AuxSpec>();
  if (!spec_words.empty()) {
    const std::string bad = parse_aux_spec(spec_words.data(), spec_words.size(), ti.cols, NUM_CHALLENGES, ti.aux.get());
    if (!bad.empty()) return etp_fail(ctx, ETP_ERR_INVALID, "%s", bad.c_str());
  }
  if ((ti.lookup() || ti.ctl()) && ti.chunk() > 2)
    return etp_fail(ctx, ETP_ERR_INVALID, "lookups / CTLs need constraint degree 2 or 3 (eval_helper_columns: \"Allow other constraint degrees\")");
  if ((int)t->prog.n_aux != ti.n_aux(NUM_CHALLENGES))
    return etp_fail(ctx, ETP_ERR_INVALID, "table: the program reads %u auxiliary columns but the lookups / CTLs produce %d", t->prog.n_aux,
                    ti.n_aux(NUM_CHALLENGES));
  t->standalone = standalone;
  if (!standalone && (int)t->prog.n_ch > (ti.ctl() ? 3 * NUM_CHALLENGES : (ti.lookup() ? NUM_CHALLENGES : 0)))
    return etp_fail(ctx, ETP_ERR_INVALID, "table: the program reads %u challenge scalars, more than its lookups / CTLs provide", t->prog.n_ch);
  if (!standalone && log2_ceil(quotient_factor(ti)) > RATE_BITS)
    return etp_fail(ctx, ETP_ERR_INVALID, "Having constraints of degree higher than the rate is not supported yet.");
  if (standalone && quotient_factor(ti) > 8) return etp_fail(ctx, ETP_ERR_INVALID, "quotient degree factors above 8 are not supported");
  // compile now so that errors surface at registration, not in the middle of a proof (compiled programs are cached
  // per process: etp_jit.cu)
  std::vector<char> cubin;
  std::string log;
  ETP_TRY(jit_compile(ctx, cprog::generate_cuda(t->prog, standalone), &cubin, &log));
  ETP_TRY(jit_load(ctx, cubin, "etp_cprog_quotient", standalone ? &t->kernel_split : &t->kernel));
  ctx->tables.push_back(t);
  guard.t = nullptr;
  *table_id_out = ETP_TABLE_FIRST_REGISTERED + (int)ctx->tables.size() - 1;
  return ETP_OK;
}

extern "C" int etp_table_register(etp_ctx* ctx, const uint64_t* program, size_t n_words, const int32_t* lookups, size_t n_lookup_words,
                                  int* table_id_out) {
  etp_bind(ctx);
  if (!ctx || !program || !table_id_out || (!lookups && n_lookup_words)) return ETP_ERR_INVALID;
  // lookups: [n_lookups, then per lookup: table_col, freq_col, n_looking, looking columns...] -> general spec words
  std::vector<uint64_t> spec;
  if (n_lookup_words) {
    const int n_trace = n_words > 2 ? (int)program[2] : 0;
    std::vector<std::tuple<std::vector<int>, int, int>> ls;
    size_t pos = 0;
    const int nl = lookups[pos++];
    if (nl < 0 || nl > 64) return etp_fail(ctx, ETP_ERR_INVALID, "table: bad number of lookups");
    for (int i = 0; i < nl; i++) {
      if (pos + 3 > n_lookup_words) return etp_fail(ctx, ETP_ERR_INVALID, "table: truncated lookup description");
      const int table_col = lookups[pos++], freq_col = lookups[pos++], m = lookups[pos++];
      if (m < 1 || m > 4096 || pos + m > n_lookup_words) return etp_fail(ctx, ETP_ERR_INVALID, "table: bad looking-column count");
      std::vector<int> looking(lookups + pos, lookups + pos + m);
      pos += m;
      for (int c : looking)
        if (c < 0 || c >= n_trace) return etp_fail(ctx, ETP_ERR_INVALID, "table: lookup column out of range");
      if (table_col < 0 || table_col >= n_trace || freq_col < 0 || freq_col >= n_trace)
        return etp_fail(ctx, ETP_ERR_INVALID, "table: lookup column out of range");
      ls.emplace_back(looking, table_col, freq_col);
    }
    if (pos != n_lookup_words) return etp_fail(ctx, ETP_ERR_INVALID, "table: trailing words in the lookup description");
    spec = simple_spec_words(ls);
  }
  return register_table(ctx, program, n_words, spec, table_id_out);
}

extern "C" int etp_table_register_ex(etp_ctx* ctx, const uint64_t* program, size_t n_words, const uint64_t* aux_spec, size_t n_spec_words,
                                     int* table_id_out) {
  etp_bind(ctx);
  if (!ctx || !program || !table_id_out || (!aux_spec && n_spec_words)) return ETP_ERR_INVALID;
  return register_table(ctx, program, n_words, std::vector<uint64_t>(aux_spec, aux_spec + n_spec_words), table_id_out);
}

// A constraint program that is not a starky table: any constraint degree up to 9, up to MAX_CH_SCALARS challenge scalars, no
// auxiliary polynomials — the plonky2 circuit prover's vanishing polynomial (etp_compute_quotient_polys_cols_dev).
extern "C" int etp_program_register(etp_ctx* ctx, const uint64_t* program, size_t n_words, int* table_id_out) {
  etp_bind(ctx);
  if (!ctx || !program || !table_id_out) return ETP_ERR_INVALID;
  return register_table(ctx, program, n_words, std::vector<uint64_t>(), table_id_out, true);
}

extern "C" int etp_aux_columns_dev(etp_ctx* ctx, int table, int log_n, const uint64_t* trace_dev, size_t col_stride,
                                   const uint64_t* lookup_challenges, int n_challenges, const uint64_t* ctl_challenges, uint64_t* aux_dev) {
  etp_bind(ctx);
  if (!ctx || !trace_dev || !aux_dev) return ETP_ERR_INVALID;
  TableInfo ti;
  if (!table_info(ctx, table, &ti)) return etp_fail(ctx, ETP_ERR_INVALID, "unknown table %d", table);
  if (log_n < 0 || log_n > 30 || n_challenges < 0 || n_challenges > NUM_CHALLENGES || col_stride < ((size_t)1 << log_n) ||
      (ti.lookup() && (!lookup_challenges || n_challenges == 0)))
    return etp_fail(ctx, ETP_ERR_INVALID, "bad arguments");
  if (ti.ctl() && n_challenges != NUM_CHALLENGES) return etp_fail(ctx, ETP_ERR_INVALID, "CTL tables take %d challenges", NUM_CHALLENGES);
  ETP_TRY(reset_zero_flag(ctx));
  ETP_TRY(aux_columns(ctx, ti, log_n, trace_dev, col_stride, lookup_challenges, n_challenges, ctl_challenges, aux_dev));
  return check_zero_flag(ctx, "a lookup / CTL denominator vanishes on the trace");
}
extern "C" int etp_lookup_helper_columns_dev(etp_ctx* ctx, int table, int log_n, const uint64_t* trace_dev, size_t col_stride,
                                             const uint64_t* challenges, int n_challenges, uint64_t* aux_dev) {
  if (!challenges) return ETP_ERR_INVALID;
  return etp_aux_columns_dev(ctx, table, log_n, trace_dev, col_stride, challenges, n_challenges, nullptr, aux_dev);
}

extern "C" int etp_compute_quotient_polys_dev(etp_ctx* ctx, int table, etp_batch* trace, etp_batch* aux, const uint64_t* lookup_challenges,
                                              int n_lookup_challenges, const uint64_t* public_inputs, const uint64_t* alphas, int n_alphas,
                                              uint64_t* out_dev) {
  etp_bind(ctx);
  if (!ctx || !trace || !alphas || !out_dev) return ETP_ERR_INVALID;
  uint64_t zero[stark::MAX_PUBLIC_INPUTS] = {};
  ETP_TRY(reset_zero_flag(ctx));
  ETP_TRY(compute_quotient(ctx, table, trace, aux, lookup_challenges ? lookup_challenges : zero, n_lookup_challenges,
                           public_inputs ? public_inputs : zero, alphas, n_alphas, out_dev));
  ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ETP_OK;
}

// compute_quotient_polys over an arbitrary list of LDE columns (one device pointer per virtual trace column): the form the
// plonky2 circuit prover needs — its vanishing polynomial reads the constants / sigmas, wires and Z / partial-product
// oracles (three PolynomialBatches) and the point x itself (the LDE of the polynomial X) as ONE program's columns.
extern "C" int etp_compute_quotient_polys_cols_dev(etp_ctx* ctx, int table, const uint64_t* const* lde_cols, size_t n_cols, int log_n, int rate_bits,
                                                   const uint64_t* challenge_scalars, int n_scalars, const uint64_t* public_inputs, const uint64_t* alphas,
                                                   int n_alphas, uint64_t* out_dev) {
  etp_bind(ctx);
  if (!ctx || !lde_cols || !alphas || !out_dev || (n_scalars && !challenge_scalars)) return ETP_ERR_INVALID;
  TableInfo ti;
  if (!table_info(ctx, table, &ti)) return etp_fail(ctx, ETP_ERR_INVALID, "unknown table %d", table);
  if (ti.lookup() || ti.ctl()) return etp_fail(ctx, ETP_ERR_INVALID, "tables with auxiliary polynomials take etp_compute_quotient_polys_dev");
  if (ti.n_pi && !public_inputs) return ETP_ERR_INVALID;
  if (log_n < 1 || rate_bits < 0 || log_n + rate_bits > 30) return etp_fail(ctx, ETP_ERR_INVALID, "bad degree / rate");
  if ((size_t)ti.cols != n_cols) return etp_fail(ctx, ETP_ERR_INVALID, "the table has %d columns, %zu were given", ti.cols, n_cols);
  std::vector<uint64_t> cols(n_cols);
  for (size_t c = 0; c < n_cols; c++) {
    if (!lde_cols[c]) return etp_fail(ctx, ETP_ERR_INVALID, "null column %zu", c);
    cols[c] = (uint64_t)(uintptr_t)lde_cols[c];
  }
  DevBuf<uint64_t> d_cols(ctx);
  ETP_TRY(d_cols.alloc(n_cols));
  ETP_CUDA(ctx, cudaMemcpyAsync(d_cols.p, cols.data(), n_cols * 8, cudaMemcpyHostToDevice, ctx->stream));
  TraceView tv;
  tv.cols_dev = (const uint64_t* const*)d_cols.p; tv.n_cols = n_cols; tv.log_n = log_n; tv.rate_bits = rate_bits;
  uint64_t zero[stark::MAX_PUBLIC_INPUTS] = {};
  ETP_TRY(compute_quotient(ctx, table, tv, nullptr, n_scalars ? challenge_scalars : zero, n_scalars, public_inputs ? public_inputs : zero, alphas,
                           n_alphas, out_dev));
  ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // cols / d_cols are read until here
  return ETP_OK;
}

extern "C" int etp_pow_grind(etp_ctx* ctx, const uint64_t state[12], int pos, int bits, uint64_t* witness_out) {
  etp_bind(ctx);
  if (!ctx || !state || !witness_out) return ETP_ERR_INVALID;
  return pow_grind(ctx, state, pos, bits, witness_out);
}

extern "C" int etp_batch_eval_at_ext_point(etp_batch* b, const uint64_t z[2], uint64_t* out) {
  etp_bind(b ? b->ctx : nullptr);
  if (!b || !z || (!out && b->n_cols)) return ETP_ERR_INVALID;
  etp_ctx* ctx = b->ctx;
  const gl::Ext ze = gl::ecanon(gl::ext(z[0], z[1]));
  DevBuf<uint64_t> l0(ctx), h0(ctx);
  stark::ExtPowTable t0;
  ETP_TRY(ext_pow_table(ctx, ze, b->log_n, l0, h0, &t0));
  std::vector<gl::Ext> e0, e1;
  ETP_TRY(eval_batch(ctx, b, t0, t0, e0, e1));
  for (size_t c = 0; c < b->n_cols; c++) { out[2 * c] = e0[c].c0; out[2 * c + 1] = e0[c].c1; }
  return ETP_OK;
}

// plonky2::plonk::prover::all_wires_permutation_partial_products, laid out as the prover commits it: out = [Z per challenge]
// ++ [partial products of challenge 0] ++ [partial products of challenge 1] ...
static int plonk_partial_products_and_zs(etp_ctx* ctx, const uint64_t* wires_dev, size_t wires_stride, const uint64_t* sigmas_dev,
                                        size_t sigmas_stride, const uint64_t* k_is, int num_routed_wires, int degree_bits, int quotient_degree_factor,
                                        const uint64_t* betas, const uint64_t* gammas, int num_challenges, uint64_t* out_dev);
extern "C" int etp_plonk_partial_products_and_zs_dev(etp_ctx* ctx, const uint64_t* wires_dev, size_t wires_stride, const uint64_t* sigmas_dev,
                                                     size_t sigmas_stride, const uint64_t* k_is, int num_routed_wires, int degree_bits,
                                                     int quotient_degree_factor, const uint64_t* betas, const uint64_t* gammas, int num_challenges,
                                                     uint64_t* out_dev) {
  etp_bind(ctx);
  return plonk_partial_products_and_zs(ctx, wires_dev, wires_stride, sigmas_dev, sigmas_stride, k_is, num_routed_wires, degree_bits,
                                       quotient_degree_factor, betas, gammas, num_challenges, out_dev);
}
static int plonk_partial_products_and_zs(etp_ctx* ctx, const uint64_t* wires_dev, size_t wires_stride, const uint64_t* sigmas_dev,
                                        size_t sigmas_stride, const uint64_t* k_is, int num_routed_wires, int degree_bits, int quotient_degree_factor,
                                        const uint64_t* betas, const uint64_t* gammas, int num_challenges, uint64_t* out_dev) {
  if (!ctx || !wires_dev || !sigmas_dev || !k_is || !betas || !gammas || !out_dev) return ETP_ERR_INVALID;
  if (num_routed_wires < 1 || num_routed_wires > 4096 || degree_bits < 0 || degree_bits > 28 || quotient_degree_factor < 1 || num_challenges < 1 ||
      num_challenges > 8 || wires_stride < ((size_t)1 << degree_bits) || sigmas_stride < ((size_t)1 << degree_bits))
    return etp_fail(ctx, ETP_ERR_INVALID, "partial_products_and_zs: bad arguments");
  const size_t n = (size_t)1 << degree_bits;
  const int n_chunks = (num_routed_wires + quotient_degree_factor - 1) / quotient_degree_factor, n_pp = n_chunks - 1;
  DevBuf<uint64_t> den(ctx), chunk(ctx), rowprod(ctx), totals(ctx), d_k(ctx);
  ETP_TRY(den.alloc((size_t)num_routed_wires * n));
  ETP_TRY(chunk.alloc((size_t)n_chunks * n));
  ETP_TRY(rowprod.alloc(n));
  const size_t nb = (n + plonk::PSCAN_BLOCK - 1) / plonk::PSCAN_BLOCK;
  ETP_TRY(totals.alloc(nb));
  ETP_TRY(d_k.alloc(num_routed_wires));
  std::vector<uint64_t> kc(k_is, k_is + num_routed_wires);
  for (auto& k : kc) k = gl::canon(k);
  ETP_CUDA(ctx, cudaMemcpyAsync(d_k.p, kc.data(), kc.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  ntt::PowTable subgroup;
  ETP_TRY(get_pow_table(ctx, gl::root_of_unity(degree_bits), degree_bits, 1, &subgroup));
  ETP_TRY(reset_zero_flag(ctx));
  for (int c = 0; c < num_challenges; c++) {
    const uint64_t beta = gl::canon(betas[c]), gamma = gl::canon(gammas[c]);
    plonk::permutation_denominators<<<blocks_for((size_t)num_routed_wires * n, 256), 256, 0, ctx->stream>>>(
        wires_dev, wires_stride, sigmas_dev, sigmas_stride, num_routed_wires, (uint32_t)n, beta, gamma, den.p);
    ETP_LAUNCH_CHECK(ctx);
    ETP_TRY(batch_inverse_dev(ctx, den.p, den.p, (size_t)num_routed_wires * n));
    plonk::permutation_chunk_products<<<blocks_for(n, 128), 128, 0, ctx->stream>>>(wires_dev, wires_stride, den.p, d_k.p, num_routed_wires,
                                                                                 quotient_degree_factor, (uint32_t)n, subgroup, beta, gamma,
                                                                                 chunk.p, rowprod.p);
    ETP_LAUNCH_CHECK(ctx);
    uint64_t* z = out_dev + (size_t)c * n;
    plonk::prod_block_totals<<<(unsigned)nb, plonk::PSCAN_THREADS, 0, ctx->stream>>>(rowprod.p, n, totals.p);
    ETP_LAUNCH_CHECK(ctx);
    plonk::prod_totals_serial<<<1, 1, 0, ctx->stream>>>(totals.p, nb);
    ETP_LAUNCH_CHECK(ctx);
    plonk::prod_finish<<<(unsigned)nb, plonk::PSCAN_THREADS, 0, ctx->stream>>>(rowprod.p, n, totals.p, z);
    ETP_LAUNCH_CHECK(ctx);
    if (n_pp > 0) {
      plonk::permutation_partial_products<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(
          z, chunk.p, n_chunks, (uint32_t)n, out_dev + ((size_t)num_challenges + (size_t)c * n_pp) * n);
      ETP_LAUNCH_CHECK(ctx);
    }
  }
  return check_zero_flag(ctx, "a permutation-argument denominator vanishes");  // also keeps kc alive until the copy is done
}

// ---- plonky2 circuit prover (plonk/prover.rs prove) ------------------------------------------------------------------
// CommonCircuitData + ProverOnlyCircuitData as far as the device steps need them.  The vanishing polynomial of the circuit
// (plonk/vanishing_poly.rs eval_vanishing_poly) arrives as a constraint program over the virtual columns
// [constants | sigmas | wires | Zs | partial products | X] (include/etp_b200.h, etp_circuit_create).
constexpr uint64_t CIRCUIT_PROOF_MAGIC = 0x42323030504C4B31ULL;  // "B200PLK1"
struct etp_circuit {
  etp_ctx* ctx = nullptr;
  int degree_bits = 0, num_constants = 0, num_routed = 0, num_wires = 0, num_challenges = 0, qdf = 0, n_pp = 0, table = -1;
  etp_fri_params fp{};
  etp_batch* constants_sigmas = nullptr;
  etp_batch* x_poly = nullptr;     // the polynomial X: its LDE column is the evaluation point of every row
  uint64_t* d_sigmas = nullptr;    // sigma values on the subgroup (num_routed x n), for the permutation argument
  std::vector<uint64_t> k_is;
  uint64_t digest[4] = {};
  size_t n() const { return (size_t)1 << degree_bits; }
  int n_zs() const { return num_challenges * (1 + n_pp); }
  int n_quot() const { return num_challenges * qdf; }
  ~etp_circuit() {
    etp_batch_free(constants_sigmas);
    etp_batch_free(x_poly);
    if (d_sigmas) dev_free(ctx, d_sigmas);
  }
};
static void host_hash_no_pad(const std::vector<uint64_t>& in, uint64_t out[4]) {
  uint64_t st[12] = {};
  for (size_t off = 0; off < in.size(); off += 8) {
    for (size_t k = 0; k < 8 && off + k < in.size(); k++) st[k] = gl::canon(in[off + k]);
    etp_host_poseidon_permute(st);
  }
  memcpy(out, st, 32);
}
static size_t circuit_proof_words(const etp_circuit* c) {
  const size_t cap_words = (size_t)4 << c->fp.cap_height;
  const size_t oc[4] = {(size_t)(c->num_constants + c->num_routed), (size_t)c->num_wires, (size_t)c->n_zs(), (size_t)c->n_quot()};
  const size_t n_open = oc[0] + oc[1] + oc[2] + c->num_challenges + oc[3];
  return HEADER_WORDS + 3 * cap_words + 2 * n_open + fri_proof_words(oc, 4, c->fp) + 4;
}

extern "C" int etp_circuit_create(etp_ctx* ctx, const uint64_t* vanishing_program, size_t n_words, const uint64_t* constants, int num_constants,
                                  const uint64_t* sigmas, const uint64_t* k_is, int num_routed_wires, int num_wires, int degree_bits,
                                  int quotient_degree_factor, int num_challenges, const etp_fri_params* fri_params, const uint64_t* circuit_digest,
                                  etp_circuit** out) {
  etp_bind(ctx);
  if (!ctx || !vanishing_program || !constants || !sigmas || !k_is || !fri_params || !out) return ETP_ERR_INVALID;
  *out = nullptr;
  ETP_TRY(check_fri_params(ctx, *fri_params));
  if (fri_params->degree_bits != degree_bits || degree_bits < 1 || num_constants < 1 || num_routed_wires < 1 || num_wires < num_routed_wires ||
      num_challenges < 1 || num_challenges > NUM_CHALLENGES || quotient_degree_factor < 1 || quotient_degree_factor > 8 ||
      (1 << fri_params->rate_bits) < quotient_degree_factor)
    return etp_fail(ctx, ETP_ERR_INVALID, "circuit: bad shape (quotient_degree_factor must be <= min(8, 2^rate_bits), num_challenges <= %d)", NUM_CHALLENGES);
  std::unique_ptr<etp_circuit> c(new etp_circuit());
  c->ctx = ctx; c->degree_bits = degree_bits; c->num_constants = num_constants; c->num_routed = num_routed_wires; c->num_wires = num_wires;
  c->num_challenges = num_challenges; c->qdf = quotient_degree_factor; c->fp = *fri_params;
  c->n_pp = (num_routed_wires + quotient_degree_factor - 1) / quotient_degree_factor - 1;
  c->k_is.assign(k_is, k_is + num_routed_wires);
  ETP_TRY(register_table(ctx, vanishing_program, n_words, std::vector<uint64_t>(), &c->table, true));
  {
    TableInfo ti;
    table_info(ctx, c->table, &ti);
    const int want_cols = num_constants + num_routed_wires + num_wires + c->n_zs() + 1;
    if (ti.cols != want_cols || quotient_factor(ti) != quotient_degree_factor || ti.n_pi > 4)
      return etp_fail(ctx, ETP_ERR_INVALID, "circuit: the vanishing program must read %d virtual columns (has %d), have constraint degree %d and at most "
                      "4 public inputs (the public-input hash)", want_cols, ti.cols, quotient_degree_factor + 1);
    if ((int)ti.reg->prog.n_ch > 2 * num_challenges) return etp_fail(ctx, ETP_ERR_INVALID, "circuit: the vanishing program reads more than the betas and gammas");
  }
  const size_t n = c->n();
  const int ncs = num_constants + num_routed_wires;
  {
    DevBuf<uint64_t> vals(ctx);
    ETP_TRY(vals.alloc((size_t)ncs * n));
    ETP_CUDA(ctx, cudaMemcpyAsync(vals.p, constants, (size_t)num_constants * n * 8, cudaMemcpyHostToDevice, ctx->stream));
    ETP_CUDA(ctx, cudaMemcpyAsync(vals.p + (size_t)num_constants * n, sigmas, (size_t)num_routed_wires * n * 8, cudaMemcpyHostToDevice, ctx->stream));
    ETP_TRY(batch_create(ctx, ncs, degree_bits, c->fp.rate_bits, 0, c->fp.cap_height, &c->constants_sigmas));
    ETP_TRY(batch_commit_from_values(c->constants_sigmas, vals.p, n));
    ETP_TRY(dev_alloc(ctx, (size_t)num_routed_wires * n * 8, (void**)&c->d_sigmas));
    ETP_CUDA(ctx, cudaMemcpyAsync(c->d_sigmas, vals.p + (size_t)num_constants * n, (size_t)num_routed_wires * n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  {
    ETP_TRY(batch_create(ctx, 1, degree_bits, c->fp.rate_bits, 0, 0, &c->x_poly));
    ETP_CUDA(ctx, cudaMemsetAsync(c->x_poly->coeffs, 0, n * 8, ctx->stream));
    const uint64_t one = 1;
    ETP_CUDA(ctx, cudaMemcpyAsync(c->x_poly->coeffs + 1, &one, 8, cudaMemcpyHostToDevice, ctx->stream));
    ETP_TRY(batch_commit_from_coeffs(c->x_poly));
  }
  if (circuit_digest) {
    for (int i = 0; i < 4; i++) c->digest[i] = gl::canon(circuit_digest[i]);
  } else {  // stand-in for CircuitData::circuit_digest: hash_no_pad(constants_sigmas cap ++ degree_bits)
    std::vector<uint64_t> in(c->constants_sigmas->cap);
    in.push_back((uint64_t)degree_bits);
    host_hash_no_pad(in, c->digest);
  }
  *out = c.release();
  return ETP_OK;
}
extern "C" void etp_circuit_free(etp_circuit* c) {
  etp_bind(c ? c->ctx : nullptr);
  delete c;
}
extern "C" int etp_circuit_digest(const etp_circuit* c, uint64_t digest_out[4]) {
  if (!c || !digest_out) return ETP_ERR_INVALID;
  memcpy(digest_out, c->digest, 32);
  return ETP_OK;
}
extern "C" int etp_circuit_constants_sigmas_cap(const etp_circuit* c, uint64_t* cap_out) {
  if (!c || !cap_out) return ETP_ERR_INVALID;
  memcpy(cap_out, c->constants_sigmas->cap.data(), c->constants_sigmas->cap.size() * 8);
  return ETP_OK;
}
extern "C" size_t etp_circuit_proof_words(const etp_circuit* c) { return c ? circuit_proof_words(c) : 0; }

// plonk::prover::prove after witness generation.  wires_dev: num_wires x n values (column-major).
static int circuit_prove_dev(etp_circuit* c, const uint64_t* wires_dev, size_t stride, const uint64_t* pi_hash_in, uint64_t* proof) {
  etp_ctx* ctx = c->ctx;
  const size_t n = c->n(), cap_words = (size_t)4 << c->fp.cap_height;
  const int log_n = c->degree_bits, rate_bits = c->fp.rate_bits, K = c->num_challenges;
  PhaseTimer timer(ctx);
  uint64_t pi_hash[4];
  for (int i = 0; i < 4; i++) pi_hash[i] = gl::canon(pi_hash_in[i]);
  struct Batches {
    etp_batch *wires = nullptr, *zs = nullptr, *quot = nullptr;
    ~Batches() { etp_batch_free(wires); etp_batch_free(zs); etp_batch_free(quot); }
  } B;
  uint64_t* w = proof;
  uint64_t* hdr = w; w += HEADER_WORDS;
  memset(hdr, 0, HEADER_WORDS * 8);
  hdr[0] = CIRCUIT_PROOF_MAGIC; hdr[1] = log_n; hdr[2] = c->num_constants; hdr[3] = c->num_routed; hdr[4] = c->num_wires; hdr[5] = K;
  hdr[6] = c->n_pp; hdr[7] = c->qdf; hdr[8] = rate_bits; hdr[9] = c->fp.cap_height; hdr[10] = c->fp.n_reductions; hdr[11] = ARITY_BITS;
  hdr[12] = (uint64_t)1 << (log_n - fri_total_arities(c->fp)); hdr[13] = c->fp.num_query_rounds; hdr[14] = c->fp.proof_of_work_bits;
  hdr[15] = circuit_proof_words(c);
  hostf::Challenger ch;
  ch.observe(c->digest, 4);
  ch.observe(pi_hash, 4);
  // ---- wires commitment
  ETP_TRY(batch_create(ctx, c->num_wires, log_n, rate_bits, 0, c->fp.cap_height, &B.wires));
  ETP_TRY(batch_commit_from_values(B.wires, wires_dev, stride));
  timer.mark("wires commit");
  ch.observe(B.wires->cap.data(), cap_words);
  memcpy(w, B.wires->cap.data(), cap_words * 8); w += cap_words;
  uint64_t betas[NUM_CHALLENGES], gammas[NUM_CHALLENGES], alphas[NUM_CHALLENGES];
  for (int i = 0; i < K; i++) betas[i] = ch.get();
  for (int i = 0; i < K; i++) gammas[i] = ch.get();
  // ---- all_wires_permutation_partial_products + commitment
  {
    DevBuf<uint64_t> zs_vals(ctx);
    ETP_TRY(zs_vals.alloc((size_t)c->n_zs() * n));
    ETP_TRY(plonk_partial_products_and_zs(ctx, wires_dev, stride, c->d_sigmas, n, c->k_is.data(), c->num_routed, log_n, c->qdf, betas, gammas, K, zs_vals.p));
    timer.mark("partial products and Zs");
    ETP_TRY(batch_create(ctx, c->n_zs(), log_n, rate_bits, 0, c->fp.cap_height, &B.zs));
    ETP_TRY(batch_commit_from_values(B.zs, zs_vals.p, n));
    timer.mark("partial products and Zs commit");
  }
  ch.observe(B.zs->cap.data(), cap_words);
  memcpy(w, B.zs->cap.data(), cap_words * 8); w += cap_words;
  for (int i = 0; i < K; i++) alphas[i] = ch.get();
  // ---- compute_quotient_polys: the vanishing program over [constants | sigmas | wires | Zs | partial products | X]
  ETP_TRY(batch_create(ctx, c->n_quot(), log_n, rate_bits, 0, c->fp.cap_height, &B.quot));
  {
    const etp_batch* src[4] = {c->constants_sigmas, B.wires, B.zs, c->x_poly};
    std::vector<uint64_t> cols;
    for (const etp_batch* b : src)
      for (size_t k = 0; k < b->n_cols; k++) cols.push_back((uint64_t)(uintptr_t)(b->lde + k * b->lde_n()));
    DevBuf<uint64_t> d_cols(ctx);
    ETP_TRY(d_cols.alloc(cols.size()));
    ETP_CUDA(ctx, cudaMemcpyAsync(d_cols.p, cols.data(), cols.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    TraceView tv;
    tv.cols_dev = (const uint64_t* const*)d_cols.p; tv.n_cols = cols.size(); tv.log_n = log_n; tv.rate_bits = rate_bits;
    uint64_t scalars[stark::MAX_CH_SCALARS] = {};
    for (int i = 0; i < K; i++) { scalars[i] = betas[i]; scalars[K + i] = gammas[i]; }
    ETP_TRY(compute_quotient(ctx, c->table, tv, nullptr, scalars, 2 * K, pi_hash, alphas, K, B.quot->coeffs));
    ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // cols / d_cols are read until here
  }
  timer.mark("compute quotient polys");
  ETP_TRY(batch_commit_from_coeffs(B.quot));
  timer.mark("quotient polys commit");
  ch.observe(B.quot->cap.data(), cap_words);
  memcpy(w, B.quot->cap.data(), cap_words * 8); w += cap_words;
  // ---- openings
  const gl::Ext zeta = ch.get_ext();
  const uint64_t g = gl::root_of_unity(log_n);
  {
    gl::Ext zp = zeta;
    for (int i = 0; i < log_n; i++) zp = gl::emul(zp, zp);
    zp = gl::ecanon(zp);
    if (zp.c0 == 1 && zp.c1 == 0) return etp_fail(ctx, ETP_ERR_PROOF, "Opening point is in the subgroup.");
  }
  const gl::Ext zeta_next = gl::ecanon(gl::emul_base(zeta, g));
  etp_batch* oracles[4] = {c->constants_sigmas, B.wires, B.zs, B.quot};
  std::vector<gl::Ext> at[4], zs_next_all, unused;
  {
    DevBuf<uint64_t> l0(ctx), h0(ctx), l1(ctx), h1(ctx);
    stark::ExtPowTable t0, t1;
    ETP_TRY(ext_pow_table(ctx, zeta, log_n, l0, h0, &t0));
    ETP_TRY(ext_pow_table(ctx, zeta_next, log_n, l1, h1, &t1));
    for (int o = 0; o < 4; o++) ETP_TRY(eval_batch(ctx, oracles[o], t0, t1, at[o], o == 2 ? zs_next_all : unused));
  }
  timer.mark("compute openings proof: evaluate at zeta, g*zeta");
  std::vector<gl::Ext> zs_next(zs_next_all.begin(), zs_next_all.begin() + K);
  // OpeningSet: constants, plonk_sigmas, wires, plonk_zs, plonk_zs_next, partial_products, quotient_polys
  auto put = [&](const gl::Ext* v, size_t cnt) { for (size_t i = 0; i < cnt; i++) { *w++ = v[i].c0; *w++ = v[i].c1; } };
  put(at[0].data(), at[0].size()); put(at[1].data(), at[1].size()); put(at[2].data(), K); put(zs_next.data(), K);
  put(at[2].data() + K, at[2].size() - K); put(at[3].data(), at[3].size());
  // observe_openings(to_fri_openings): the zeta batch (constants, sigmas, wires, zs, partial products, quotient), then zs_next
  auto obs = [&](const std::vector<gl::Ext>& v) { for (auto& e : v) { ch.observe(e.c0); ch.observe(e.c1); } };
  for (int o = 0; o < 4; o++) obs(at[o]);
  obs(zs_next);
  // ---- get_fri_instance + prove_openings
  std::vector<etp_fri_poly> p0, p1;
  std::vector<std::vector<gl::Ext>> ys(2);
  for (uint32_t o = 0; o < 4; o++)
    for (uint32_t k = 0; k < oracles[o]->n_cols; k++) { p0.push_back({o, k}); ys[0].push_back(at[o][k]); }
  for (uint32_t k = 0; k < (uint32_t)K; k++) { p1.push_back({2, k}); ys[1].push_back(zs_next[k]); }
  etp_fri_batch batches[2];
  batches[0] = {{zeta.c0, zeta.c1}, p0.data(), p0.size()};
  batches[1] = {{zeta_next.c0, zeta_next.c1}, p1.data(), p1.size()};
  ETP_TRY(prove_openings(ctx, batches, 2, oracles, 4, ch, c->fp, &ys, w, &timer));
  const size_t oc[4] = {oracles[0]->n_cols, oracles[1]->n_cols, oracles[2]->n_cols, oracles[3]->n_cols};
  w += fri_proof_words(oc, 4, c->fp);
  for (int i = 0; i < 4; i++) *w++ = pi_hash[i];
  if ((size_t)(w - proof) != hdr[15]) return etp_fail(ctx, ETP_ERR_STATE, "internal error: circuit proof size mismatch");
  timer.finish();
  return ETP_OK;
}

extern "C" int etp_circuit_prove_dev(etp_circuit* c, const uint64_t* wires_dev, size_t col_stride, const uint64_t public_inputs_hash[4],
                                     uint64_t* proof_out) {
  etp_bind(c ? c->ctx : nullptr);
  if (!c || !wires_dev || !public_inputs_hash || !proof_out) return ETP_ERR_INVALID;
  if (col_stride < c->n()) return etp_fail(c->ctx, ETP_ERR_INVALID, "circuit: wires stride below n");
  return circuit_prove_dev(c, wires_dev, col_stride, public_inputs_hash, proof_out);
}
extern "C" int etp_circuit_prove_host(etp_circuit* c, const uint64_t* wires, const uint64_t public_inputs_hash[4], uint64_t* proof_out) {
  etp_bind(c ? c->ctx : nullptr);
  if (!c || !wires || !public_inputs_hash || !proof_out) return ETP_ERR_INVALID;
  etp_ctx* ctx = c->ctx;
  DevBuf<uint64_t> d(ctx);
  ETP_TRY(d.alloc((size_t)c->num_wires * c->n()));
  ETP_CUDA(ctx, cudaMemcpyAsync(d.p, wires, (size_t)c->num_wires * c->n() * 8, cudaMemcpyHostToDevice, ctx->stream));
  return circuit_prove_dev(c, d.p, c->n(), public_inputs_hash, proof_out);
}

extern "C" size_t etp_fri_proof_words(const size_t* oracle_num_cols, size_t n_oracles, const etp_fri_params* params) {
  if (!params || (!oracle_num_cols && n_oracles) || params->n_reductions < 0 || params->n_reductions > 16) return 0;
  return fri_proof_words(oracle_num_cols, n_oracles, *params);
}
extern "C" int etp_prove_openings(etp_ctx* ctx, const etp_fri_batch* batches, size_t n_batches, etp_batch* const* oracles, size_t n_oracles,
                                  etp_challenger* challenger, const etp_fri_params* params, uint64_t* fri_proof_out) {
  etp_bind(ctx);
  if (!ctx || !batches || !oracles || !n_oracles || !params || !fri_proof_out) return ETP_ERR_INVALID;
  if (!challenger_ok(challenger)) return etp_fail(ctx, ETP_ERR_INVALID, "bad challenger state");
  for (size_t b = 0; b < n_batches; b++)
    if (!batches[b].polynomials && batches[b].n_polynomials) return etp_fail(ctx, ETP_ERR_INVALID, "FRI batch %zu has no polynomial list", b);
  hostf::Challenger ch(*challenger);
  PhaseTimer timer(ctx);
  ETP_TRY(prove_openings(ctx, batches, n_batches, oracles, n_oracles, ch, *params, nullptr, fri_proof_out, &timer));
  timer.finish();
  *challenger = ch;
  return ETP_OK;
}

extern "C" int etp_fri_begin(etp_ctx* ctx, const uint64_t* values_dev, const etp_fri_params* params, etp_fri_state** out) {
  etp_bind(ctx);
  if (!ctx || !values_dev || !params || !out) return ETP_ERR_INVALID;
  *out = nullptr;
  ETP_TRY(check_fri_params(ctx, *params));
  const size_t words = (size_t)2 << (params->degree_bits + params->rate_bits);
  uint64_t* v = nullptr;
  ETP_TRY(dev_alloc(ctx, words * 8, (void**)&v));
  stark::canon_copy<<<(unsigned)((words + 255) / 256), 256, 0, ctx->stream>>>(values_dev, v, words);
  ctx->launches++;
  if (cudaGetLastError() != cudaSuccess) { dev_free(ctx, v); return etp_fail(ctx, ETP_ERR_CUDA, "copy kernel failed"); }
  return fri_begin_owned(ctx, v, *params, out);
}
extern "C" int etp_fri_commit_layer(etp_fri_state* s, uint64_t* cap_out) {
  etp_bind(s ? s->ctx : nullptr);
  if (!s || !cap_out) return ETP_ERR_INVALID;
  return fri_commit_layer(s, cap_out);
}
extern "C" int etp_fri_fold(etp_fri_state* s, const uint64_t beta[2]) {
  etp_bind(s ? s->ctx : nullptr);
  if (!s || !beta) return ETP_ERR_INVALID;
  return fri_fold(s, gl::ext(beta[0], beta[1]));
}
extern "C" int etp_fri_final_poly(etp_fri_state* s, uint64_t* coeffs_out) {
  etp_bind(s ? s->ctx : nullptr);
  if (!s || !coeffs_out) return ETP_ERR_INVALID;
  std::vector<uint64_t> fc;
  ETP_TRY(fri_final_poly(s, &fc));
  memcpy(coeffs_out, fc.data(), fc.size() * 8);
  return ETP_OK;
}
extern "C" int etp_fri_commit_phase(etp_fri_state* s, etp_challenger* challenger, uint64_t* caps_out, uint64_t* final_poly_out) {
  etp_bind(s ? s->ctx : nullptr);
  if (!s || !caps_out || !final_poly_out) return ETP_ERR_INVALID;
  if (!challenger_ok(challenger)) return etp_fail(s->ctx, ETP_ERR_INVALID, "bad challenger state");
  hostf::Challenger ch(*challenger);
  std::vector<uint64_t> fc;
  ETP_TRY(fri_commit_phase(s, ch, caps_out, &fc));
  memcpy(final_poly_out, fc.data(), fc.size() * 8);
  *challenger = ch;
  return ETP_OK;
}
extern "C" int etp_fri_proof_of_work(etp_ctx* ctx, etp_challenger* challenger, int proof_of_work_bits, uint64_t* witness_out) {
  etp_bind(ctx);
  if (!ctx || !witness_out) return ETP_ERR_INVALID;
  if (!challenger_ok(challenger)) return etp_fail(ctx, ETP_ERR_INVALID, "bad challenger state");
  if (proof_of_work_bits < 0 || proof_of_work_bits > 40) return etp_fail(ctx, ETP_ERR_INVALID, "unsupported proof_of_work_bits");
  hostf::Challenger ch(*challenger);
  ETP_TRY(fri_proof_of_work(ctx, ch, proof_of_work_bits, witness_out));
  *challenger = ch;
  return ETP_OK;
}
extern "C" int etp_fri_query_rounds(etp_fri_state* s, etp_batch* const* oracles, size_t n_oracles, const uint64_t* x_indices, size_t n_indices,
                                    uint64_t* out) {
  etp_bind(s ? s->ctx : nullptr);
  if (!s || (!oracles && n_oracles) || (!x_indices && n_indices) || (!out && n_indices)) return ETP_ERR_INVALID;
  return fri_query_rounds(s, oracles, n_oracles, x_indices, n_indices, out);
}
extern "C" void etp_fri_free(etp_fri_state* s) {
  etp_bind(s ? s->ctx : nullptr);
  delete s;
}

// ---- quotient / openings / FRI over a column-split table (etp_shard) ------------------------------------------
// The rank that calls these (the "leader" of a column-split proof) reads the trace columns where they live: its own HBM
// or, through the mappings of etp_shard_set_peer, a peer's over NVLink — the same fused access as the leaf hashing.
static int shard_column_table(etp_shard* s, DevBuf<uint64_t>& d_cols) {
  etp_ctx* ctx = s->ctx;
  std::vector<uint64_t> cols(s->n_cols_total);
  for (size_t c = 0; c < s->n_cols_total; c++) {
    const uint64_t* p = s->column(c);
    if (!p) return etp_fail(ctx, ETP_ERR_STATE, "shard: LDE of rank %zu not mapped (etp_shard_set_peer)", c / s->cps);
    cols[c] = (uint64_t)(uintptr_t)p;
  }
  ETP_TRY(d_cols.alloc(cols.size()));
  ETP_CUDA(ctx, cudaMemcpyAsync(d_cols.p, cols.data(), cols.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ETP_OK;
}

// out[k] = in[bitrev(k)]: an LDE column (bit-reversed rows, possibly in a peer's HBM) -> natural order, local
static __global__ void k_bitrev_gather(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, int bits) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >> bits) return;
  out[k] = in[gl::bitrev32((uint32_t)k, bits)];
}

// All auxiliary polynomials (values on the trace domain) of a column-split table, on the calling rank.  Only the few trace
// columns the lookups / CTLs read are needed as VALUES; they are recovered from the LDE the peers already map: bit-reversed
// gather over NVLink -> coset iFFT (size 2^(log_n + rate_bits)) -> FFT (size n).  ctl_zs_first_out: Z(1) of every CTL Z.
extern "C" int etp_shard_aux_columns_dev(etp_shard* s, int table, const uint64_t* lookup_challenges, int n_challenges, const uint64_t* ctl_challenges,
                                         uint64_t* aux_out_dev, uint64_t* ctl_zs_first_out) {
  etp_bind(s ? s->ctx : nullptr);
  if (!s || !aux_out_dev) return ETP_ERR_INVALID;
  etp_ctx* ctx = s->ctx;
  TableInfo ti;
  if (!table_info(ctx, table, &ti)) return etp_fail(ctx, ETP_ERR_INVALID, "unknown table %d", table);
  if ((size_t)ti.cols != s->n_cols_total) return etp_fail(ctx, ETP_ERR_INVALID, "the table does not have the shard's number of columns");
  if (n_challenges < 0 || n_challenges > NUM_CHALLENGES || (n_challenges && !lookup_challenges)) return etp_fail(ctx, ETP_ERR_INVALID, "bad lookup challenges");
  if (!ti.lookup() && !ti.ctl()) return ETP_OK;
  std::vector<int> used;
  auto compact = std::make_shared<AuxSpec>();
  const std::string why = compact_aux_spec(*ti.aux, ti.cols, NUM_CHALLENGES, &used, compact.get());
  if (!why.empty()) return etp_fail(ctx, ETP_ERR_STATE, "internal error: %s", why.c_str());
  const int log_n = s->log_n, log_lde = s->log_n + s->rate_bits;
  const size_t n = s->n(), lde_n = s->lde_n();
  DevBuf<uint64_t> vals(ctx), nat(ctx), co(ctx), scratch(ctx), d_spec(ctx);
  ETP_TRY(vals.alloc((used.empty() ? 1 : used.size()) * n));
  ETP_TRY(nat.alloc(lde_n));
  ETP_TRY(co.alloc(lde_n));
  ETP_TRY(scratch.alloc(lde_n));
  for (size_t j = 0; j < used.size(); j++) {
    const uint64_t* col = s->column((size_t)used[j]);
    if (!col) return etp_fail(ctx, ETP_ERR_STATE, "shard: LDE of rank %zu not mapped (etp_shard_set_peer)", (size_t)used[j] / s->cps);
    k_bitrev_gather<<<blocks_for(lde_n, 256), 256, 0, ctx->stream>>>(col, nat.p, log_lde);
    ETP_LAUNCH_CHECK(ctx);
    NttArgs a;
    a.in = nat.p; a.in_stride = lde_n; a.n_in = (uint32_t)lde_n; a.out = co.p; a.out_stride = lde_n; a.scratch = scratch.p; a.scratch_stride = lde_n;
    a.log_n = log_lde; a.n_cols = 1; a.inverse = true; a.natural_out = true; a.coset_shift = gl::GENERATOR;
    ETP_TRY(ntt_run(ctx, a));
    a = NttArgs();
    a.in = co.p; a.in_stride = lde_n; a.n_in = (uint32_t)n; a.out = vals.p + j * n; a.out_stride = n; a.scratch = scratch.p; a.scratch_stride = n;
    a.log_n = log_n; a.n_cols = 1; a.natural_out = true;
    ETP_TRY(ntt_run(ctx, a));
  }
  TableInfo tc = ti;
  tc.aux = compact;
  if (!compact->words.empty()) {
    ETP_TRY(d_spec.alloc(compact->words.size()));
    ETP_CUDA(ctx, cudaMemcpyAsync(d_spec.p, compact->words.data(), compact->words.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  }
  ETP_TRY(reset_zero_flag(ctx));
  ETP_TRY(aux_columns(ctx, tc, log_n, vals.p, n, lookup_challenges, n_challenges, ctl_challenges, aux_out_dev, d_spec.p));
  const int n_zs = ti.n_ctl_zs();
  if (n_zs && ctl_zs_first_out) {
    const size_t first = (size_t)(ti.n_lookup_cols(n_challenges) + ti.n_ctl_helpers()) * n;
    ETP_CUDA(ctx, cudaMemcpy2DAsync(ctl_zs_first_out, 8, aux_out_dev + first, n * 8, 8, n_zs, cudaMemcpyDeviceToHost, ctx->stream));
  }
  ETP_TRY(check_zero_flag(ctx, "a lookup / CTL denominator vanishes on the trace"));  // synchronises the stream
  for (int k = 0; k < n_zs && ctl_zs_first_out; k++) ctl_zs_first_out[k] = gl::canon(ctl_zs_first_out[k]);
  return ETP_OK;
}

// compute_quotient_polys over the split trace.  aux: the committed auxiliary polynomials (a batch of this rank) or NULL;
// challenge_scalars: what the constraint program reads (lookup challenges, then the CTL (beta, gamma) pairs).
extern "C" int etp_shard_compute_quotient_polys_dev(etp_shard* s, int table, etp_batch* aux, const uint64_t* challenge_scalars, int n_scalars,
                                                    const uint64_t* public_inputs, const uint64_t* alphas, int n_alphas, uint64_t* out_dev) {
  etp_bind(s ? s->ctx : nullptr);
  if (!s || !alphas || !out_dev || (n_scalars && !challenge_scalars)) return ETP_ERR_INVALID;
  etp_ctx* ctx = s->ctx;
  TableInfo ti;
  if (!table_info(ctx, table, &ti)) return etp_fail(ctx, ETP_ERR_INVALID, "unknown table %d", table);
  if (ti.n_pi && !public_inputs) return ETP_ERR_INVALID;
  if (aux && (aux->ctx != ctx || aux->log_n != s->log_n || aux->rate_bits != s->rate_bits))
    return etp_fail(ctx, ETP_ERR_INVALID, "auxiliary batch does not match the shard");
  DevBuf<uint64_t> d_cols(ctx);
  ETP_TRY(shard_column_table(s, d_cols));
  TraceView tv;
  tv.cols_dev = (const uint64_t* const*)d_cols.p; tv.n_cols = s->n_cols_total; tv.log_n = s->log_n; tv.rate_bits = s->rate_bits;
  uint64_t zero[stark::MAX_PUBLIC_INPUTS] = {};
  ETP_TRY(compute_quotient(ctx, table, tv, aux, n_scalars ? challenge_scalars : zero, n_scalars, public_inputs ? public_inputs : zero, alphas, n_alphas,
                           out_dev));
  ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // d_cols is read by the kernel
  return ETP_OK;
}

extern "C" int etp_shard_eval_at_ext_points(etp_shard* s, const uint64_t z0[2], const uint64_t z1[2], uint64_t* out0, uint64_t* out1) {
  etp_bind(s ? s->ctx : nullptr);
  if (!s || !z0 || !z1 || ((!out0 || !out1) && s->local_cols)) return ETP_ERR_INVALID;
  if (s->local_cols == 0) return ETP_OK;
  etp_ctx* ctx = s->ctx;
  DevBuf<uint64_t> l0(ctx), h0(ctx), l1(ctx), h1(ctx);
  stark::ExtPowTable t0, t1;
  ETP_TRY(ext_pow_table(ctx, gl::ecanon(gl::ext(z0[0], z0[1])), s->log_n, l0, h0, &t0));
  ETP_TRY(ext_pow_table(ctx, gl::ecanon(gl::ext(z1[0], z1[1])), s->log_n, l1, h1, &t1));
  std::vector<gl::Ext> e0, e1;
  ETP_TRY(eval_coeffs(ctx, s->coeffs, s->n(), s->local_cols, t0, t1, e0, e1));
  for (size_t c = 0; c < s->local_cols; c++) {
    out0[2 * c] = e0[c].c0; out0[2 * c + 1] = e0[c].c1;
    out1[2 * c] = e1[c].c0; out1[2 * c + 1] = e1[c].c1;
  }
  return ETP_OK;
}

extern "C" int etp_shard_fri_begin(etp_shard* s, etp_batch* const* extra_oracles, size_t n_extra, const etp_fri_batch* batches, size_t n_batches,
                                   const uint64_t* ys, const uint64_t alpha[2], const etp_fri_params* params, etp_fri_state** out) {
  etp_bind(s ? s->ctx : nullptr);
  if (!s || (!extra_oracles && n_extra) || !batches || !ys || !alpha || !params || !out) return ETP_ERR_INVALID;
  etp_ctx* ctx = s->ctx;
  *out = nullptr;
  ETP_TRY(check_fri_params(ctx, *params));
  if (params->degree_bits != s->log_n || params->rate_bits != s->rate_bits || params->cap_height != s->cap_height)
    return etp_fail(ctx, ETP_ERR_INVALID, "the FRI parameters do not match the shard");
  for (size_t o = 0; o < n_extra; o++)
    if (!extra_oracles[o] || extra_oracles[o]->ctx != ctx || extra_oracles[o]->log_n != s->log_n || extra_oracles[o]->rate_bits != s->rate_bits ||
        extra_oracles[o]->cap_height != s->cap_height)
      return etp_fail(ctx, ETP_ERR_INVALID, "FRI oracle %zu does not match the FRI parameters", o + 1);
  if (n_batches < 1 || n_batches > (size_t)stark::MAX_FRI_BATCHES) return etp_fail(ctx, ETP_ERR_INVALID, "between 1 and %d FRI batches are supported", stark::MAX_FRI_BATCHES);
  std::vector<std::vector<gl::Ext>> yv(n_batches);
  {
    const uint64_t* y = ys;
    for (size_t b = 0; b < n_batches; b++) {
      if (!batches[b].polynomials && batches[b].n_polynomials) return ETP_ERR_INVALID;
      yv[b].resize(batches[b].n_polynomials);
      for (size_t k = 0; k < batches[b].n_polynomials; k++, y += 2) yv[b][k] = gl::ecanon(gl::ext(y[0], y[1]));
    }
  }
  const size_t lde_n = s->lde_n();
  const ColResolver col = [&](uint32_t o, uint32_t k) -> const uint64_t* {
    if (o == 0) return s->column(k);
    if (o - 1 >= n_extra || k >= extra_oracles[o - 1]->n_cols) return nullptr;
    return extra_oracles[o - 1]->lde + (size_t)k * lde_n;
  };
  uint64_t* values = nullptr;
  ETP_TRY(dev_alloc(ctx, 2 * lde_n * 8, (void**)&values));
  etp_fri_state* st = nullptr;
  ETP_TRY(fri_begin_owned(ctx, values, *params, &st));
  std::unique_ptr<etp_fri_state> guard(st);
  ETP_TRY(combine_on_lde(ctx, batches, n_batches, col, s->log_n, s->rate_bits, gl::ecanon(gl::ext(alpha[0], alpha[1])), yv, values));
  *out = guard.release();
  return ETP_OK;
}

extern "C" size_t etp_stark_proof_words(const etp_ctx* ctx, int table, int log_n) {
  TableInfo ti;
  // 0 = no such proof: unknown table, or a degree for which cap_height > log2(LDE size) (upstream: MerkleTree::new asserts)
  if (!table_info(ctx, table, &ti) || log_n < 1 || log_n > 29 || log_n + RATE_BITS < CAP_HEIGHT) return 0;
  return proof_words(ti, log_n);
}

extern "C" int etp_prove_with_commitment(etp_ctx* ctx, int table, etp_batch* trace_commitment, const uint64_t* trace_dev, size_t col_stride,
                                         const uint64_t* ctl_challenges, etp_challenger* challenger, const uint64_t* public_inputs,
                                         uint64_t* proof_out) {
  etp_bind(ctx);
  if (!ctx || !trace_commitment || !trace_dev || !proof_out) return ETP_ERR_INVALID;
  if (!challenger_ok(challenger)) return etp_fail(ctx, ETP_ERR_INVALID, "bad challenger state");
  TableInfo ti;
  if (!table_info(ctx, table, &ti)) return etp_fail(ctx, ETP_ERR_INVALID, "unknown table %d", table);
  if (col_stride < trace_commitment->n()) return etp_fail(ctx, ETP_ERR_INVALID, "trace column stride is smaller than the trace length");
  uint64_t zero[stark::MAX_PUBLIC_INPUTS] = {};
  hostf::Challenger ch(*challenger);
  PhaseTimer timer(ctx);
  ETP_TRY(prove_with_commitment(ctx, table, ti, trace_commitment, trace_dev, col_stride, ctl_challenges, ch, public_inputs ? public_inputs : zero,
                                proof_out, timer));
  timer.finish();
  *challenger = ch;
  return ETP_OK;
}

extern "C" int etp_stark_prove_dev(etp_ctx* ctx, int table, int log_n, const uint64_t* trace_dev, size_t col_stride,
                                   const uint64_t* public_inputs, uint64_t* proof_out) {
  etp_bind(ctx);
  if (!ctx || !trace_dev || !proof_out) return ETP_ERR_INVALID;
  uint64_t zero[stark::MAX_PUBLIC_INPUTS] = {};
  return stark_prove_dev(ctx, table, log_n, trace_dev, col_stride, public_inputs ? public_inputs : zero, proof_out);
}

extern "C" int etp_stark_prove_host(etp_ctx* ctx, int table, int log_n, const uint64_t* trace, const uint64_t* public_inputs,
                                    uint64_t* proof_out) {
  etp_bind(ctx);
  if (!ctx || !trace || !proof_out) return ETP_ERR_INVALID;
  TableInfo ti;
  if (!table_info(ctx, table, &ti)) return etp_fail(ctx, ETP_ERR_INVALID, "unknown table %d", table);
  if (log_n < 1 || log_n > 29) return etp_fail(ctx, ETP_ERR_INVALID, "unsupported degree_bits %d", log_n);
  const size_t n = (size_t)1 << log_n;
  DevBuf<uint64_t> d(ctx);
  ETP_TRY(d.alloc((size_t)ti.cols * n));
  uint64_t zero[stark::MAX_PUBLIC_INPUTS] = {};
  return stark_prove_dev(ctx, table, log_n, d.p, n, public_inputs ? public_inputs : zero, proof_out, trace);
}

extern "C" int etp_last_prove_timings(const etp_ctx* ctx, const char** names, float* ms, int max) {
  if (!ctx) return 0;
  int k = 0;
  for (auto& t : ctx->timings) {
    if (k >= max) break;
    names[k] = t.first;
    ms[k] = t.second;
    k++;
  }
  return k;
}
