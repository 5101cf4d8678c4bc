// libetp_b200: the starky / FRI half of the C ABI — lookup helper columns, compute_quotient_polys,
// openings, FRI commit phase (evaluation-domain folding), proof of work, query rounds and the full
// single-table `starky::prover::prove` under StarkConfig::standard_fast_config()
// (/root/reference/common/src/prover_state/circuit.rs:204; reached from /root/reference/ops/src/lib.rs:52).
// See include/etp_b200.h for the upstream item behind each entry point and DESIGN.md for the proof
// wire format.  Product code: no oracle, no CPU fallback; the only host arithmetic is the transcript
// (Challenger) and the <= 2^8-coefficient FRI final polynomial.
#include <time.h>

#include <cstdlib>

#include "cprog.h"
#include "ctx.cuh"
#include "host_field.h"
#include "stark_kernels.cuh"

namespace {

// StarkConfig::standard_fast_config()
constexpr int NUM_CHALLENGES = 2, RATE_BITS = 1, CAP_HEIGHT = 4, POW_BITS = 16, ARITY_BITS = 4, FINAL_POLY_BITS = 5,
              NUM_QUERIES = 84;
constexpr uint64_t PROOF_MAGIC = 0x4232303053544B31ULL;  // "B200STK1"

// starky::lookup::Lookup without filters: `looking` columns are looked up in `table_col` with multiplicities
// `freq_col`.  Helper columns: one per chunk of (constraint_degree - 1) looking columns, then Z.
struct LookupInfo {
  std::vector<int> looking;
  int table_col = 0, freq_col = 0;
};
struct TableInfo {
  int cols = 0, degree = 0, n_pi = 0;
  std::vector<LookupInfo> lookups;
  RegisteredTable* reg = nullptr;  // program-defined table
  int chunk() const { return degree - 1 < 1 ? 1 : degree - 1; }
  int helpers(const LookupInfo& l) const { return ((int)l.looking.size() + chunk() - 1) / chunk(); }
  int aux_per_challenge() const { int a = 0; for (auto& l : lookups) a += helpers(l) + 1; return a; }
  int n_aux(int n_ch) const { return aux_per_challenge() * n_ch; }
  bool lookup() const { return !lookups.empty(); }
};
}  // namespace

// a table registered at run time: its description + constraint program + the kernel compiled from it
struct RegisteredTable {
  TableInfo info;
  cprog::Program prog;
  JitKernel kernel;
};
void free_registered_tables(etp_ctx* ctx) {
  for (auto* t : ctx->tables) { jit_unload(&t->kernel); delete t; }
  ctx->tables.clear();
}

namespace {
bool table_info(const etp_ctx* ctx, int t, TableInfo* o) {
  if (t == ETP_TABLE_FIBONACCI) { *o = TableInfo(); o->cols = 2; o->degree = 2; o->n_pi = 3; return true; }
  if (t == ETP_TABLE_MEMORY) {
    *o = TableInfo(); o->cols = 21; o->degree = 3; o->n_pi = 0;
    LookupInfo l; l.looking = {stark::M_RANGE_CHECK}; l.table_col = stark::M_COUNTER; l.freq_col = stark::M_FREQ;
    o->lookups.push_back(l);
    return true;
  }
  if (ctx && t >= ETP_TABLE_FIRST_REGISTERED && (size_t)(t - ETP_TABLE_FIRST_REGISTERED) < ctx->tables.size()) {
    *o = ctx->tables[t - ETP_TABLE_FIRST_REGISTERED]->info;
    return true;
  }
  return false;
}
int quotient_factor(const TableInfo& ti) { return ti.degree - 1 < 1 ? 1 : ti.degree - 1; }
int log2_ceil(int x) { int l = 0; while ((1 << l) < x) l++; return l; }
int fri_num_layers(int degree_bits) {  // FriReductionStrategy::ConstantArityBits(4, 5)
  int layers = 0;
  while (degree_bits > FINAL_POLY_BITS && degree_bits + RATE_BITS - ARITY_BITS >= CAP_HEIGHT) { layers++; degree_bits -= ARITY_BITS; }
  return layers;
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

unsigned blocks_for(size_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

int batch_inverse_dev(etp_ctx* ctx, const uint64_t* in, uint64_t* out, size_t n) {
  const int threads = 128;
  const size_t per_block = (size_t)threads * stark::INV_K;
  stark::batch_inverse<<<blocks_for(n, (int)per_block), threads, 0, ctx->stream>>>(in, out, n);
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

int lookup_helper_columns(etp_ctx* ctx, int table, int log_n, const uint64_t* trace, size_t stride, const uint64_t* ch, int n_ch,
                          uint64_t* aux) {
  TableInfo ti;
  if (!table_info(ctx, table, &ti)) return etp_fail(ctx, ETP_ERR_INVALID, "unknown table %d", table);
  if (!ti.lookup()) return ETP_OK;
  const size_t n = (size_t)1 << log_n;
  size_t max_m = 0;
  for (auto& l : ti.lookups) max_m = l.looking.size() > max_m ? l.looking.size() : max_m;
  DevBuf<uint64_t> den(ctx), term(ctx);
  DevBuf<int> d_cols(ctx);
  ETP_TRY(den.alloc((max_m + 1) * n));
  ETP_TRY(term.alloc(n));
  ETP_TRY(d_cols.alloc(max_m + 1));
  // auxiliary column order (starky prover.rs): for each lookup, for each challenge: helpers..., Z
  uint64_t* out = aux;
  for (auto& l : ti.lookups) {
    const int m = (int)l.looking.size(), nh = ti.helpers(l);
    std::vector<int> cols(l.looking);
    cols.push_back(l.table_col);
    ETP_CUDA(ctx, cudaMemcpyAsync(d_cols.p, cols.data(), cols.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `cols` dies at the end of this iteration
    for (int k = 0; k < n_ch; k++) {
      stark::lookup_denominators<<<blocks_for((size_t)(m + 1) * n, 256), 256, 0, ctx->stream>>>(trace, stride, d_cols.p, m + 1,
                                                                                                gl::canon(ch[k]), n, den.p);
      ETP_LAUNCH_CHECK(ctx);
      ETP_TRY(batch_inverse_dev(ctx, den.p, den.p, (size_t)(m + 1) * n));
      stark::lookup_terms<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(den.p, m, ti.chunk(), trace + (size_t)l.freq_col * stride, n, out,
                                                                      term.p);
      ETP_LAUNCH_CHECK(ctx);
      ETP_TRY(exclusive_scan_dev(ctx, term.p, out + (size_t)nh * n, n));
      out += (size_t)(nh + 1) * n;
    }
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
int eval_batch(etp_ctx* ctx, const etp_batch* b, const stark::ExtPowTable& t0, const stark::ExtPowTable& t1, gl::Ext z0, gl::Ext z1,
               std::vector<gl::Ext>& out0, std::vector<gl::Ext>& out1) {
  const uint32_t n = (uint32_t)b->n();
  const int np = (int)b->n_cols;
  out0.assign(np, gl::ext(0, 0));
  out1.assign(np, gl::ext(0, 0));
  if (np == 0) return ETP_OK;
  const unsigned gx = blocks_for(n, stark::OPEN_THREADS * stark::OPEN_CHUNK);
  const unsigned gy = (unsigned)np;
  DevBuf<uint64_t> partial(ctx);
  ETP_TRY(partial.alloc((size_t)gx * np * 4));
  stark::eval_polys_at_two_points<<<dim3(gx, gy), stark::OPEN_THREADS, 0, ctx->stream>>>(b->coeffs, b->n(), np, n, t0, t1, partial.p);
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

int compute_quotient(etp_ctx* ctx, int table, etp_batch* trace, etp_batch* aux, const uint64_t* lookup_ch, int n_lookup_ch,
                     const uint64_t* pi, const uint64_t* alphas, int n_alphas, uint64_t* out_dev) {
  TableInfo ti;
  if (!table_info(ctx, table, &ti)) return etp_fail(ctx, ETP_ERR_INVALID, "unknown table %d", table);
  if (!trace || (int)trace->n_cols != ti.cols) return etp_fail(ctx, ETP_ERR_INVALID, "trace batch has the wrong number of columns");
  if (n_alphas < 1 || n_alphas > stark::MAX_CHALLENGES) return etp_fail(ctx, ETP_ERR_INVALID, "unsupported number of challenges");
  const int n_aux = ti.n_aux(n_lookup_ch);
  if (ti.lookup() && (!aux || (int)aux->n_cols != n_aux || n_lookup_ch > stark::MAX_CHALLENGES))
    return etp_fail(ctx, ETP_ERR_INVALID, "auxiliary batch does not match the table's lookups");
  if (ti.reg && ((int)ti.reg->prog.n_aux != n_aux || (int)ti.reg->prog.n_ch > n_lookup_ch))
    return etp_fail(ctx, ETP_ERR_INVALID, "constraint program expects %u auxiliary columns / %u challenges, got %d / %d",
                    ti.reg->prog.n_aux, ti.reg->prog.n_ch, n_aux, n_lookup_ch);
  const int log_n = trace->log_n, rate_bits = trace->rate_bits;
  const int factor = quotient_factor(ti), qbits = log2_ceil(factor);
  if (qbits > rate_bits)
    return etp_fail(ctx, ETP_ERR_INVALID, "Having constraints of degree higher than the rate is not supported yet.");
  const int log_size = log_n + qbits, log_lde = log_n + rate_bits;
  const size_t size = (size_t)1 << log_size;
  stark::QuotientParams q{};
  q.trace = trace->lde; q.trace_stride = trace->lde_n();
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
  for (int j = 0; j < n_lookup_ch; j++) q.lookup_ch[j] = gl::canon(lookup_ch[j]);
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
    void* args[] = {(void*)&q};
    ETP_CUDA(ctx, cudaLaunchKernel((const void*)ti.reg->kernel.kernel, dim3(gb), dim3(128), args, 0, ctx->stream));
    ctx->launches++;
  } else if (table == ETP_TABLE_FIBONACCI) {
    stark::quotient_kernel<0><<<gb, 128, 0, ctx->stream>>>(q);
  } else {
    stark::quotient_kernel<1><<<gb, 128, 0, ctx->stream>>>(q);
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

size_t proof_words(const TableInfo& ti, int log_n) {
  const int n_aux = ti.n_aux(NUM_CHALLENGES), n_quot = quotient_factor(ti) * NUM_CHALLENGES;
  const int n_layers = fri_num_layers(log_n), log_lde = log_n + RATE_BITS;
  const size_t cap = (size_t)4 << CAP_HEIGHT;
  size_t w = 16 + cap * (2 + (n_aux ? 1 : 0)) + 2 * (size_t)(2 * ti.cols + 2 * n_aux + n_quot) + cap * n_layers;
  const int init_path = log_lde - CAP_HEIGHT;
  size_t per_query = ti.cols + 4 * init_path + (n_aux ? n_aux + 4 * init_path : 0) + n_quot + 4 * init_path;
  int bits = log_lde;
  for (int l = 0; l < n_layers; l++) { bits -= ARITY_BITS; per_query += 2 * (1 << ARITY_BITS) + 4 * (bits - CAP_HEIGHT); }
  w += NUM_QUERIES * per_query;
  w += 2 * ((size_t)1 << (log_n - ARITY_BITS * n_layers)) + 1 + ti.n_pi;
  return w;
}

struct FriLayer {
  uint64_t* values = nullptr;  // n ext, bit-reversed order (== leaves, 2^arity ext per row)
  uint64_t* levels = nullptr;
  size_t n_leaves = 0;
  int log_n = 0;
  std::vector<uint64_t> cap;
};

// trace_host != nullptr: the trace (n_cols x n, column-major, stride n) is still on the host; trace_dev is an empty device
// buffer of the same shape that the streamed trace commit fills column group by column group while it transforms and
// hashes the groups that have already arrived (etp_stark_prove_host).
int stark_prove_dev(etp_ctx* ctx, int table, int log_n, const uint64_t* trace_dev, size_t stride, const uint64_t* pi_in, uint64_t* proof,
                    const uint64_t* trace_host = nullptr) {
  TableInfo ti;
  if (!table_info(ctx, table, &ti)) return etp_fail(ctx, ETP_ERR_INVALID, "unknown table %d", table);
  if (log_n < 1 || log_n + RATE_BITS > 30) return etp_fail(ctx, ETP_ERR_INVALID, "unsupported degree_bits %d", log_n);
  const int n_layers = fri_num_layers(log_n);
  if (ARITY_BITS * n_layers > log_n + RATE_BITS - CAP_HEIGHT || log_n + RATE_BITS < CAP_HEIGHT)
    return etp_fail(ctx, ETP_ERR_INVALID, "FRI total reduction arity is too large.");
  const size_t n = (size_t)1 << log_n;
  const int log_lde = log_n + RATE_BITS;
  const size_t lde_n = (size_t)1 << log_lde, cap_words = (size_t)4 << CAP_HEIGHT;
  const int n_aux = ti.n_aux(NUM_CHALLENGES), factor = quotient_factor(ti), n_quot = factor * NUM_CHALLENGES;
  uint64_t pi[stark::MAX_PUBLIC_INPUTS] = {};
  for (int i = 0; i < ti.n_pi; i++) pi[i] = gl::canon(pi_in[i]);
  PhaseTimer timer(ctx);

  uint64_t* w = proof;
  uint64_t* hdr = w; w += 16;
  hdr[0] = PROOF_MAGIC; hdr[1] = table; hdr[2] = log_n; hdr[3] = ti.cols; hdr[4] = n_aux; hdr[5] = n_quot; hdr[6] = CAP_HEIGHT;
  hdr[7] = n_layers; hdr[8] = ARITY_BITS; hdr[9] = (uint64_t)1 << (log_n - ARITY_BITS * n_layers); hdr[10] = NUM_QUERIES; hdr[11] = ti.n_pi;
  hdr[12] = RATE_BITS; hdr[13] = POW_BITS; hdr[14] = NUM_CHALLENGES; hdr[15] = proof_words(ti, log_n);

  struct Batches {
    etp_batch *trace = nullptr, *aux = nullptr, *quot = nullptr;
    ~Batches() { etp_batch_free(trace); etp_batch_free(aux); etp_batch_free(quot); }
  } B;

  // ---- prove(): trace commitment; the challenger observes the public inputs, then the trace cap
  ETP_TRY(batch_create(ctx, ti.cols, log_n, RATE_BITS, 0, CAP_HEIGHT, &B.trace));
  if (trace_host) {
    std::vector<const uint64_t*> cols(ti.cols);
    for (int c = 0; c < ti.cols; c++) cols[c] = trace_host + (size_t)c * n;
    ETP_TRY(batch_commit_from_host_streamed(B.trace, cols.data(), true, const_cast<uint64_t*>(trace_dev)));
  } else {
    ETP_TRY(batch_commit_from_values(B.trace, trace_dev, stride));
  }
  timer.mark("trace commit (IFFT + FFT + Merkle tree)");
  hostf::Challenger ch;
  ch.observe(pi, ti.n_pi);
  ch.observe(B.trace->cap.data(), cap_words);
  memcpy(w, B.trace->cap.data(), cap_words * 8); w += cap_words;

  // ---- prove_with_commitment: lookup helper columns + auxiliary commitment
  uint64_t lookup_ch[NUM_CHALLENGES] = {0, 0};
  if (ti.lookup()) {
    // get_grand_product_challenge_set: (beta, gamma) per challenge, the lookup argument uses beta
    for (int k = 0; k < NUM_CHALLENGES; k++) { lookup_ch[k] = ch.get(); (void)ch.get(); }
    DevBuf<uint64_t> aux_vals(ctx);
    ETP_TRY(aux_vals.alloc((size_t)n_aux * n));
    ETP_TRY(lookup_helper_columns(ctx, table, log_n, trace_dev, stride, lookup_ch, NUM_CHALLENGES, aux_vals.p));
    timer.mark("compute lookup helper columns");
    ETP_TRY(batch_create(ctx, n_aux, log_n, RATE_BITS, 0, CAP_HEIGHT, &B.aux));
    ETP_TRY(batch_commit_from_values(B.aux, aux_vals.p, n));
    timer.mark("auxiliary polys commit");
    ch.observe(B.aux->cap.data(), cap_words);
    memcpy(w, B.aux->cap.data(), cap_words * 8); w += cap_words;
  }
  uint64_t alphas[NUM_CHALLENGES];
  for (int j = 0; j < NUM_CHALLENGES; j++) alphas[j] = ch.get();

  // ---- quotient
  ETP_TRY(batch_create(ctx, n_quot, log_n, RATE_BITS, 0, CAP_HEIGHT, &B.quot));
  ETP_TRY(compute_quotient(ctx, table, B.trace, B.aux, lookup_ch, ti.lookup() ? NUM_CHALLENGES : 0, pi, alphas, NUM_CHALLENGES, B.quot->coeffs));
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
    ETP_TRY(eval_batch(ctx, B.trace, t0, t1, zeta, zeta_next, tr0, tr1));
    if (B.aux) ETP_TRY(eval_batch(ctx, B.aux, t0, t1, zeta, zeta_next, ax0, ax1));
    ETP_TRY(eval_batch(ctx, B.quot, t0, t1, zeta, zeta_next, qu0, qu1));
  }
  timer.mark("compute openings proof: evaluate at zeta, g*zeta");
  auto put = [&](const std::vector<gl::Ext>& v) { for (auto& e : v) { *w++ = e.c0; *w++ = e.c1; } };
  put(tr0); put(tr1); put(ax0); put(ax1); put(qu0);
  // observe_openings(to_fri_openings): zeta batch = local ++ aux ++ quotient ; next batch = next ++ aux_next
  auto obs = [&](const std::vector<gl::Ext>& v) { for (auto& e : v) { ch.observe(e.c0); ch.observe(e.c1); } };
  obs(tr0); obs(ax0); obs(qu0); obs(tr1); obs(ax1);

  // ---- PolynomialBatch::prove_openings, in evaluation form over the LDE coset
  const gl::Ext alpha = ch.get_ext();
  const int n0 = ti.cols + n_aux + n_quot, n1 = ti.cols + n_aux;
  std::vector<uint64_t> apow(2 * (size_t)(n0 + 1));
  gl::Ext y0 = gl::ext(0, 0), y1 = gl::ext(0, 0), shift0;
  {
    gl::Ext cur = gl::ext(1, 0);
    std::vector<gl::Ext> all0 = tr0, all1 = tr1;
    all0.insert(all0.end(), ax0.begin(), ax0.end());
    all0.insert(all0.end(), qu0.begin(), qu0.end());
    all1.insert(all1.end(), ax1.begin(), ax1.end());
    for (int k = 0; k <= n0; k++) {
      cur = gl::ecanon(cur);
      apow[2 * k] = cur.c0; apow[2 * k + 1] = cur.c1;
      if (k < n0) y0 = gl::eadd(y0, gl::emul(cur, all0[k]));
      if (k < n1) y1 = gl::eadd(y1, gl::emul(cur, all1[k]));
      if (k == n1) shift0 = cur;
      cur = gl::emul(cur, alpha);
    }
  }
  std::vector<FriLayer> layers(n_layers + 1);
  struct LayerGuard {
    etp_ctx* ctx; std::vector<FriLayer>* l;
    ~LayerGuard() { for (auto& x : *l) { dev_free(ctx, x.values); dev_free(ctx, x.levels); } }
  } guard{ctx, &layers};
  layers[0].log_n = log_lde;
  ETP_TRY(dev_alloc(ctx, 2 * lde_n * 8, (void**)&layers[0].values));
  {
    DevBuf<uint64_t> d_apow(ctx), den(ctx);
    ETP_TRY(d_apow.alloc(apow.size()));
    ETP_TRY(den.alloc(2 * lde_n));
    ETP_CUDA(ctx, cudaMemcpyAsync(d_apow.p, apow.data(), apow.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    stark::CombineParams c{};
    c.cols[0] = B.trace->lde; c.strides[0] = lde_n; c.n_cols[0] = ti.cols;
    c.cols[1] = B.aux ? B.aux->lde : nullptr; c.strides[1] = lde_n; c.n_cols[1] = n_aux;
    c.cols[2] = B.quot->lde; c.strides[2] = lde_n; c.n_cols[2] = n_quot;
    c.n1 = n1; c.log_lde = log_lde;
    ETP_TRY(get_pow_table(ctx, gl::root_of_unity(log_lde), log_lde, gl::GENERATOR, &c.coset));
    c.alpha_pows = d_apow.p;
    c.y0 = gl::ecanon(y0); c.y1 = gl::ecanon(y1); c.z0 = zeta; c.z1 = zeta_next; c.shift0 = shift0;
    c.seven_z0c1_sq = gl::canon(gl::mul(7, gl::mul(zeta.c1, zeta.c1)));
    c.seven_z1c1_sq = gl::canon(gl::mul(7, gl::mul(zeta_next.c1, zeta_next.c1)));
    c.den = den.p; c.out = layers[0].values;
    stark::combine_norms<<<blocks_for(lde_n, 256), 256, 0, ctx->stream>>>(c);
    ETP_LAUNCH_CHECK(ctx);
    ETP_TRY(batch_inverse_dev(ctx, den.p, den.p, 2 * lde_n));
    // few rows, many columns: one thread per point cannot fill the machine, so the column sums are split into chunks
    DevBuf<uint64_t> partial(ctx);
    if (n0 > 256 && lde_n * 4 <= (size_t)1 << 20) {
      c.col_chunk = 64;
      c.n_chunks = (n0 + c.col_chunk - 1) / c.col_chunk;
      ETP_TRY(partial.alloc((size_t)c.n_chunks * lde_n * 4));
      c.partial = partial.p;
      stark::combine_accumulate<<<dim3(blocks_for(lde_n, 128), c.n_chunks), 128, 0, ctx->stream>>>(c);
      ETP_LAUNCH_CHECK(ctx);
    }
    stark::combine_values<<<blocks_for(lde_n, 128), 128, 0, ctx->stream>>>(c);
    ETP_LAUNCH_CHECK(ctx);
    ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // apow (host) and d_apow stay alive until here
  }
  timer.mark("compute openings proof: combine on the LDE domain");

  // ---- fri_committed_trees: evaluation-domain folding, shift_l = 7^(16^l)
  uint64_t shift = gl::GENERATOR;
  uint64_t w16_inv_pows[16];
  {
    const uint64_t wi = gl::canon(gl::inv(gl::root_of_unity(ARITY_BITS)));
    uint64_t cur = 1;
    for (int i = 0; i < 16; i++) { w16_inv_pows[i] = gl::canon(cur); cur = gl::mul(cur, wi); }
  }
  for (int l = 0; l < n_layers; l++) {
    FriLayer& L = layers[l];
    const size_t cur_n = (size_t)1 << L.log_n;
    L.n_leaves = cur_n >> ARITY_BITS;
    L.cap.resize(cap_words);
    ETP_TRY(dev_alloc(ctx, levels_words(L.n_leaves, CAP_HEIGHT) * 8, (void**)&L.levels));
    ETP_TRY(launch_leaf_hash_rowmajor(ctx, L.values, 2 << ARITY_BITS, L.n_leaves, L.levels));
    ETP_TRY(merkle_build_levels(ctx, L.levels, L.n_leaves, CAP_HEIGHT, L.cap.data()));
    ch.observe(L.cap.data(), cap_words);
    memcpy(w, L.cap.data(), cap_words * 8); w += cap_words;
    const gl::Ext beta = ch.get_ext();
    FriLayer& Nx = layers[l + 1];
    Nx.log_n = L.log_n - ARITY_BITS;
    ETP_TRY(dev_alloc(ctx, (2 * cur_n >> ARITY_BITS) * 8, (void**)&Nx.values));
    stark::FoldParams f{};
    f.in = L.values; f.out = Nx.values; f.log_n = L.log_n; f.beta = beta;
    ETP_TRY(get_pow_table(ctx, gl::canon(gl::inv(gl::root_of_unity(L.log_n))), L.log_n, gl::canon(gl::inv(shift)), &f.x0_inv));
    memcpy(f.w16_inv_pows, w16_inv_pows, sizeof w16_inv_pows);
    f.inv16 = gl::canon(gl::inv(16));
    stark::fri_fold16<<<blocks_for(L.n_leaves, 128), 128, 0, ctx->stream>>>(f);
    ETP_LAUNCH_CHECK(ctx);
    shift = gl::canon(gl::pow(shift, 16));
  }
  timer.mark("fold codewords in the commitment phase");

  // ---- final polynomial: coset iDFT of the last layer's values on the host (<= 2^8 points)
  const FriLayer& F = layers[n_layers];
  const size_t fin_n = (size_t)1 << F.log_n, final_len = fin_n >> RATE_BITS;
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
    const uint64_t si = gl::canon(gl::inv(shift));
    uint64_t cur = 1;
    for (size_t k = 0; k < fin_n; k++) { fin[k] = gl::ecanon(gl::emul_base(fin[k], cur)); cur = gl::mul(cur, si); }
  }
  for (size_t k = final_len; k < fin_n; k++)
    if (fin[k].c0 || fin[k].c1)
      return etp_fail(ctx, ETP_ERR_PROOF, "FRI final polynomial has degree >= %zu: the quotient is not a polynomial (trace violates the constraints)", final_len);
  std::vector<uint64_t> final_coeffs(2 * final_len);
  for (size_t k = 0; k < final_len; k++) { final_coeffs[2 * k] = fin[k].c0; final_coeffs[2 * k + 1] = fin[k].c1; }
  ch.observe(final_coeffs.data(), final_coeffs.size());

  // ---- fri_proof_of_work
  uint64_t st[12];
  memcpy(st, ch.state, sizeof st);
  for (int i = 0; i < ch.n_in; i++) st[i] = ch.in[i];
  uint64_t pow_witness = 0;
  ETP_TRY(pow_grind(ctx, st, ch.n_in, POW_BITS, &pow_witness));
  ch.observe(pow_witness);
  const uint64_t pow_response = ch.get();
  if (POW_BITS && (pow_response >> (64 - POW_BITS)) != 0) return etp_fail(ctx, ETP_ERR_PROOF, "proof of work response mismatch");
  timer.mark("find proof-of-work witness");

  // ---- fri_prover_query_rounds
  std::vector<uint64_t> qidx(NUM_QUERIES);
  for (int qn = 0; qn < NUM_QUERIES; qn++) qidx[qn] = ch.get() % lde_n;
  const int init_path = log_lde - CAP_HEIGHT;
  const etp_batch* init[3] = {B.trace, B.aux, B.quot};
  // device staging: per oracle rows + paths, per layer rows + paths
  size_t stage_words = 0;
  for (int o = 0; o < 3; o++)
    if (init[o]) stage_words += (size_t)NUM_QUERIES * (init[o]->n_cols + 4 * init_path);
  {
    int bits = log_lde;
    for (int l = 0; l < n_layers; l++) { bits -= ARITY_BITS; stage_words += (size_t)NUM_QUERIES * (32 + 4 * (bits - CAP_HEIGHT)); }
  }
  DevBuf<uint64_t> stage(ctx), d_idx(ctx);
  ETP_TRY(stage.alloc(stage_words));
  ETP_TRY(d_idx.alloc((size_t)NUM_QUERIES * (n_layers + 1)));
  std::vector<uint64_t> idx_all((size_t)NUM_QUERIES * (n_layers + 1));
  for (int qn = 0; qn < NUM_QUERIES; qn++) {
    uint64_t x = qidx[qn];
    idx_all[qn] = x;
    for (int l = 0; l < n_layers; l++) { x >>= ARITY_BITS; idx_all[(size_t)(l + 1) * NUM_QUERIES + qn] = x; }
  }
  ETP_CUDA(ctx, cudaMemcpyAsync(d_idx.p, idx_all.data(), idx_all.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  std::vector<size_t> off_rows, off_paths;
  size_t off = 0;
  for (int o = 0; o < 3; o++) {
    if (!init[o]) { off_rows.push_back(0); off_paths.push_back(0); continue; }
    const int nc = (int)init[o]->n_cols;
    off_rows.push_back(off);
    merkle::gather_rows_colmajor<<<blocks_for((size_t)NUM_QUERIES * nc, 256), 256, 0, ctx->stream>>>(init[o]->lde, lde_n, nc, d_idx.p,
                                                                                                    NUM_QUERIES, stage.p + off);
    ETP_LAUNCH_CHECK(ctx);
    off += (size_t)NUM_QUERIES * nc;
    off_paths.push_back(off);
    if (init_path > 0) {
      stark::gather_paths<<<blocks_for((size_t)NUM_QUERIES * init_path, 256), 256, 0, ctx->stream>>>(init[o]->levels, (uint32_t)lde_n, init_path,
                                                                                                   d_idx.p, NUM_QUERIES, stage.p + off);
      ETP_LAUNCH_CHECK(ctx);
    }
    off += (size_t)NUM_QUERIES * 4 * init_path;
  }
  std::vector<size_t> loff_rows(n_layers), loff_paths(n_layers);
  for (int l = 0; l < n_layers; l++) {
    const FriLayer& L = layers[l];
    const int path = L.log_n - ARITY_BITS - CAP_HEIGHT;
    loff_rows[l] = off;
    stark::gather_rows_rowmajor<<<blocks_for((size_t)NUM_QUERIES * 32, 256), 256, 0, ctx->stream>>>(
        L.values, 32, d_idx.p + (size_t)(l + 1) * NUM_QUERIES, NUM_QUERIES, stage.p + off);
    ETP_LAUNCH_CHECK(ctx);
    off += (size_t)NUM_QUERIES * 32;
    loff_paths[l] = off;
    if (path > 0) {
      stark::gather_paths<<<blocks_for((size_t)NUM_QUERIES * path, 256), 256, 0, ctx->stream>>>(
          L.levels, (uint32_t)L.n_leaves, path, d_idx.p + (size_t)(l + 1) * NUM_QUERIES, NUM_QUERIES, stage.p + off);
      ETP_LAUNCH_CHECK(ctx);
    }
    off += (size_t)NUM_QUERIES * 4 * path;
  }
  std::vector<uint64_t> sh(stage_words ? stage_words : 1);
  ETP_CUDA(ctx, cudaMemcpyAsync(sh.data(), stage.p, stage_words * 8, cudaMemcpyDeviceToHost, ctx->stream));
  ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int qn = 0; qn < NUM_QUERIES; qn++) {
    for (int o = 0; o < 3; o++) {
      if (!init[o]) continue;
      const size_t nc = init[o]->n_cols;
      memcpy(w, &sh[off_rows[o] + qn * nc], nc * 8); w += nc;
      memcpy(w, &sh[off_paths[o] + (size_t)qn * 4 * init_path], (size_t)4 * init_path * 8); w += 4 * init_path;
    }
    for (int l = 0; l < n_layers; l++) {
      const int path = layers[l].log_n - ARITY_BITS - CAP_HEIGHT;
      memcpy(w, &sh[loff_rows[l] + (size_t)qn * 32], 32 * 8); w += 32;
      memcpy(w, &sh[loff_paths[l] + (size_t)qn * 4 * path], (size_t)4 * path * 8); w += 4 * path;
    }
  }
  timer.mark("build FRI query rounds");
  memcpy(w, final_coeffs.data(), final_coeffs.size() * 8); w += final_coeffs.size();
  *w++ = pow_witness;
  for (int i = 0; i < ti.n_pi; i++) *w++ = pi[i];
  if ((size_t)(w - proof) != hdr[15]) return etp_fail(ctx, ETP_ERR_STATE, "internal error: proof size mismatch");
  timer.finish();
  return ETP_OK;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" int etp_table_num_columns(const etp_ctx* c, int t) { TableInfo ti; return table_info(c, t, &ti) ? ti.cols : -1; }
extern "C" int etp_table_constraint_degree(const etp_ctx* c, int t) { TableInfo ti; return table_info(c, t, &ti) ? ti.degree : -1; }
extern "C" int etp_table_num_public_inputs(const etp_ctx* c, int t) { TableInfo ti; return table_info(c, t, &ti) ? ti.n_pi : -1; }
extern "C" int etp_table_num_aux_columns(const etp_ctx* c, int t, int nc) { TableInfo ti; return table_info(c, t, &ti) ? ti.n_aux(nc) : -1; }
extern "C" int etp_table_quotient_degree_factor(const etp_ctx* c, int t) { TableInfo ti; return table_info(c, t, &ti) ? quotient_factor(ti) : -1; }

extern "C" int etp_table_register(etp_ctx* ctx, const uint64_t* program, size_t n_words, const int32_t* lookups, size_t n_lookup_words,
                                  int* table_id_out) {
  etp_bind(ctx);
  if (!ctx || !program || !table_id_out || (!lookups && n_lookup_words)) return ETP_ERR_INVALID;
  auto t = new RegisteredTable();
  struct Guard { RegisteredTable* t; ~Guard() { if (t) { jit_unload(&t->kernel); delete t; } } } guard{t};
  const std::string why = cprog::parse(program, n_words, stark::MAX_PUBLIC_INPUTS, stark::MAX_CHALLENGES, &t->prog);
  if (!why.empty()) return etp_fail(ctx, ETP_ERR_INVALID, "%s", why.c_str());
  TableInfo& ti = t->info;
  ti.cols = (int)t->prog.n_trace; ti.degree = (int)t->prog.degree; ti.n_pi = (int)t->prog.n_pi; ti.reg = t;
  // lookups: [n_lookups, then per lookup: table_col, freq_col, n_looking, looking columns...]
  if (n_lookup_words) {
    size_t pos = 0;
    const int nl = lookups[pos++];
    if (nl < 0 || nl > 64) return etp_fail(ctx, ETP_ERR_INVALID, "table: bad number of lookups");
    for (int i = 0; i < nl; i++) {
      if (pos + 3 > n_lookup_words) return etp_fail(ctx, ETP_ERR_INVALID, "table: truncated lookup description");
      LookupInfo l;
      l.table_col = lookups[pos++]; l.freq_col = lookups[pos++];
      const int m = lookups[pos++];
      if (m < 1 || m > 4096 || pos + m > n_lookup_words) return etp_fail(ctx, ETP_ERR_INVALID, "table: bad looking-column count");
      for (int j = 0; j < m; j++) l.looking.push_back(lookups[pos++]);
      for (int c : l.looking)
        if (c < 0 || c >= ti.cols) return etp_fail(ctx, ETP_ERR_INVALID, "table: lookup column out of range");
      if (l.table_col < 0 || l.table_col >= ti.cols || l.freq_col < 0 || l.freq_col >= ti.cols)
        return etp_fail(ctx, ETP_ERR_INVALID, "table: lookup column out of range");
      ti.lookups.push_back(l);
    }
    if (pos != n_lookup_words) return etp_fail(ctx, ETP_ERR_INVALID, "table: trailing words in the lookup description");
  }
  if ((int)t->prog.n_aux != ti.n_aux(NUM_CHALLENGES))
    return etp_fail(ctx, ETP_ERR_INVALID, "table: the program reads %u auxiliary columns but the lookups produce %d", t->prog.n_aux,
                    ti.n_aux(NUM_CHALLENGES));
  if (log2_ceil(quotient_factor(ti)) > RATE_BITS)
    return etp_fail(ctx, ETP_ERR_INVALID, "Having constraints of degree higher than the rate is not supported yet.");
  // same program registered before on this context: share the compiled kernel's source hash -> recompile is cheap
  // enough to skip a cache; compile now so that errors surface at registration, not in the middle of a proof
  std::vector<char> cubin;
  std::string log;
  ETP_TRY(jit_compile(ctx, cprog::generate_cuda(t->prog), &cubin, &log));
  ETP_TRY(jit_load(ctx, cubin, "etp_cprog_quotient", &t->kernel));
  ctx->tables.push_back(t);
  guard.t = nullptr;
  *table_id_out = ETP_TABLE_FIRST_REGISTERED + (int)ctx->tables.size() - 1;
  return ETP_OK;
}

extern "C" int etp_lookup_helper_columns_dev(etp_ctx* ctx, int table, int log_n, const uint64_t* trace_dev, size_t col_stride,
                                             const uint64_t* challenges, int n_challenges, uint64_t* aux_dev) {
  etp_bind(ctx);
  if (!ctx || !trace_dev || !challenges || !aux_dev) return ETP_ERR_INVALID;
  if (log_n < 0 || log_n > 30 || n_challenges < 0) return etp_fail(ctx, ETP_ERR_INVALID, "bad arguments");
  ETP_TRY(lookup_helper_columns(ctx, table, log_n, trace_dev, col_stride, challenges, n_challenges, aux_dev));
  ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ETP_OK;
}

extern "C" int etp_compute_quotient_polys_dev(etp_ctx* ctx, int table, etp_batch* trace, etp_batch* aux, const uint64_t* lookup_challenges,
                                              int n_lookup_challenges, const uint64_t* public_inputs, const uint64_t* alphas, int n_alphas,
                                              uint64_t* out_dev) {
  etp_bind(ctx);
  if (!ctx || !trace || !alphas || !out_dev) return ETP_ERR_INVALID;
  uint64_t zero[stark::MAX_PUBLIC_INPUTS] = {};
  ETP_TRY(compute_quotient(ctx, table, trace, aux, lookup_challenges ? lookup_challenges : zero, n_lookup_challenges,
                           public_inputs ? public_inputs : zero, alphas, n_alphas, out_dev));
  ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ETP_OK;
}

extern "C" int etp_pow_grind(etp_ctx* ctx, const uint64_t state[12], int pos, int bits, uint64_t* witness_out) {
  etp_bind(ctx);
  if (!ctx || !state || !witness_out) return ETP_ERR_INVALID;
  return pow_grind(ctx, state, pos, bits, witness_out);
}

extern "C" size_t etp_stark_proof_words(const etp_ctx* ctx, int table, int log_n) {
  TableInfo ti;
  if (!table_info(ctx, table, &ti) || log_n < 1 || log_n > 29) return 0;
  return proof_words(ti, log_n);
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
