/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Single-table STARK prover under StarkConfig::standard_fast_config()
 * (selected by the reference at /root/reference/common/src/prover_state/circuit.rs:204).
 * Restates (third-party, not on disk; pins in /root/reference/Cargo.lock:3441,4529,1675):
 *   starky 0.4.0   src/prover.rs            prove, prove_with_commitment, compute_quotient_polys
 *                  src/constraint_consumer.rs  ConstraintConsumer
 *                  src/vanishing_poly.rs    eval_vanishing_poly
 *                  src/lookup.rs            Lookup, lookup_helper_columns, eval_packed_lookups_generic
 *                  src/proof.rs             StarkOpeningSet::new / to_fri_openings
 *                  src/config.rs            StarkConfig::standard_fast_config, fri_params
 *                  src/fibonacci_stark.rs   FibonacciStark (upstream example table)
 *   plonky2 0.2.2  src/fri/oracle.rs        PolynomialBatch::prove_openings
 *                  src/fri/prover.rs        fri_proof, fri_committed_trees, fri_proof_of_work,
 *                                           fri_prover_query_rounds
 *                  src/fri/reduction_strategies.rs  ConstantArityBits(4, 5)
 *   evm_arithmetization 0.1.3  src/memory/{columns,memory_stark}.rs — the MEMORY table below is the
 *                  recalled sketch of SURVEY.md Appendix A (shape + constraint order), not a verified
 *                  copy: "memory-shaped", parity unpinned.
 *
 * Parity conventions fixed here: outputs canonical; PoW witness = the SMALLEST valid
 * one (upstream uses rayon find_any, which is only deterministic single-threaded).
 *
 * Flat proof layout (u64 words) is documented in DESIGN.md ("proof wire format") and mirrored by
 * eth_tx_proof_b200/csrc and tests/stark_verifier.py.
 */
#include "oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---------------- config ---------------- */
#define NUM_CHALLENGES 2 /* StarkConfig::standard_fast_config().num_challenges */
#define PROOF_MAGIC 0x4232303053544B32ULL /* "B200STK2" */
#define HEADER_WORDS 24
#define MEM_TRIE_DATA_SEGMENT 13ULL
#define MAX_CH_SCALARS 8 /* lookup challenges [0..2) then CTL (beta, gamma) pairs [2..6) */

/* StarkConfig::standard_fast_config().fri_config: rate_bits 1, cap_height 4, proof_of_work_bits 16,
 * ConstantArityBits(4, 5), num_query_rounds 84 (starky/src/config.rs) */
void orc_fri_params_standard_fast(int degree_bits, orc_fri_params *p) { orc_fri_params_make(degree_bits, 1, 4, 16, 84, p); }
/* FriConfig::fri_params with FriReductionStrategy::ConstantArityBits(4, 5) (plonky2/src/fri/reduction_strategies.rs) */
void orc_fri_params_make(int degree_bits, int rate_bits, int cap_height, int pow_bits, int num_queries, orc_fri_params *p) {
  memset(p, 0, sizeof *p);
  p->degree_bits = degree_bits; p->rate_bits = rate_bits; p->cap_height = cap_height; p->proof_of_work_bits = pow_bits;
  p->num_query_rounds = num_queries;
  int db = degree_bits;
  while (db > 5 && db + rate_bits - 4 >= cap_height) { p->reduction_arity_bits[p->n_reductions++] = 4; db -= 4; }
}
static int fri_total_arities(const orc_fri_params *p) { int t = 0; for (int i = 0; i < p->n_reductions; i++) t += p->reduction_arity_bits[i]; return t; }

/* ---------------- tables ---------------- */
/* Program-defined tables (constraint programs, format of eth_tx_proof_b200/csrc/cprog.h restated here): the oracle
 * INTERPRETS the program op by op; the product compiles it with NVRTC.  ids >= 16.
 * starky/src/lookup.rs Column / Filter / Lookup and starky/src/cross_table_lookup.rs CtlZData, as data: */
#define ORC_MAX_TABLES 8192
#define CPROG_MAGIC 0x3147525043505445ULL
#define AUXSPEC_MAGIC 0x3153585541505445ULL /* "ETPAUXS1" */
enum { OP_CONST = 0, OP_LV, OP_NV, OP_LA, OP_NA, OP_PI, OP_CH, OP_ADD, OP_SUB, OP_MUL, OP_EMIT, OP_EMIT_TRANSITION, OP_EMIT_FIRST,
       OP_EMIT_LAST };
typedef struct { int n_local, n_next; int *lcol, *ncol; uint64_t *lcoef, *ncoef; uint64_t constant; } column_t;
typedef struct { int n_prod, n_const; column_t *pa, *pb, *consts; } filter_t;
typedef struct { int n_cols; column_t *cols; filter_t *filters; column_t table, freq; } lookup_t;
typedef struct { int n_cols; column_t *cols; filter_t filter; } colset_t;
typedef struct { int challenge, n_sets; colset_t *sets; } ctlz_t;
typedef struct {
  int used, cols, degree, n_pi, n_lookups, n_zs;
  lookup_t *lookups;
  ctlz_t *zs;
  uint32_t n_ops, n_aux, n_ch;
  uint64_t *ops; /* 2 words per op; NULL for the built-in tables */
} table_t;
static table_t g_tables[ORC_MAX_TABLES];

static column_t column_single(int c) {
  column_t k; memset(&k, 0, sizeof k);
  k.n_local = 1; k.lcol = (int *)malloc(sizeof(int)); k.lcoef = (uint64_t *)malloc(8); k.lcol[0] = c; k.lcoef[0] = 1;
  return k;
}
static column_t column_constant(uint64_t v) { column_t k; memset(&k, 0, sizeof k); k.constant = v; return k; }
/* Filter::default(): no products, constants = [Column::constant(1)] — evaluates to 1 */
static filter_t filter_default(void) {
  filter_t f; memset(&f, 0, sizeof f);
  f.n_const = 1; f.consts = (column_t *)malloc(sizeof(column_t)); f.consts[0] = column_constant(1);
  return f;
}
static const table_t *get_table(int t) {
  if (!g_tables[ORC_TABLE_MEMORY].used) {
    table_t *f = &g_tables[ORC_TABLE_FIBONACCI], *m = &g_tables[ORC_TABLE_MEMORY];
    f->used = 1; f->cols = 2; f->degree = 2; f->n_pi = 3;
    m->used = 1; m->cols = 21; m->degree = 3; m->n_pi = 0; m->n_lookups = 1;
    m->lookups = (lookup_t *)calloc(1, sizeof(lookup_t));
    m->lookups[0].n_cols = 1; m->lookups[0].cols = (column_t *)malloc(sizeof(column_t)); m->lookups[0].cols[0] = column_single(18);
    m->lookups[0].filters = (filter_t *)malloc(sizeof(filter_t)); m->lookups[0].filters[0] = filter_default();
    m->lookups[0].table = column_single(19); m->lookups[0].freq = column_single(20);
  }
  if (t < 0 || t >= ORC_MAX_TABLES || !g_tables[t].used) { fprintf(stderr, "oracle: unknown table %d\n", t); abort(); }
  return &g_tables[t];
}
/* --- auxiliary-column spec parser (word format: include/etp_b200.h, etp_table_register_ex) --- */
typedef struct { const uint64_t *w; size_t n, pos; int bad; } rd_t;
static uint64_t rd(rd_t *r) { if (r->pos >= r->n) { r->bad = 1; return 0; } return r->w[r->pos++]; }
static column_t rd_column(rd_t *r, int n_trace) {
  column_t k; memset(&k, 0, sizeof k);
  k.n_local = (int)rd(r);
  if (k.n_local < 0 || k.n_local > 65536) { r->bad = 1; return k; }
  k.lcol = (int *)malloc(sizeof(int) * (k.n_local + 1)); k.lcoef = (uint64_t *)malloc(8 * (k.n_local + 1));
  for (int i = 0; i < k.n_local; i++) { k.lcol[i] = (int)rd(r); k.lcoef[i] = gl_canon(rd(r)); if (k.lcol[i] < 0 || k.lcol[i] >= n_trace) r->bad = 1; }
  k.n_next = (int)rd(r);
  if (k.n_next < 0 || k.n_next > 65536) { r->bad = 1; return k; }
  k.ncol = (int *)malloc(sizeof(int) * (k.n_next + 1)); k.ncoef = (uint64_t *)malloc(8 * (k.n_next + 1));
  for (int i = 0; i < k.n_next; i++) { k.ncol[i] = (int)rd(r); k.ncoef[i] = gl_canon(rd(r)); if (k.ncol[i] < 0 || k.ncol[i] >= n_trace) r->bad = 1; }
  k.constant = gl_canon(rd(r));
  return k;
}
static filter_t rd_filter(rd_t *r, int n_trace) {
  filter_t f; memset(&f, 0, sizeof f);
  f.n_prod = (int)rd(r);
  if (f.n_prod < 0 || f.n_prod > 4096) { r->bad = 1; return f; }
  f.pa = (column_t *)calloc(f.n_prod + 1, sizeof(column_t)); f.pb = (column_t *)calloc(f.n_prod + 1, sizeof(column_t));
  for (int i = 0; i < f.n_prod; i++) { f.pa[i] = rd_column(r, n_trace); f.pb[i] = rd_column(r, n_trace); }
  f.n_const = (int)rd(r);
  if (f.n_const < 0 || f.n_const > 4096) { r->bad = 1; return f; }
  f.consts = (column_t *)calloc(f.n_const + 1, sizeof(column_t));
  for (int i = 0; i < f.n_const; i++) f.consts[i] = rd_column(r, n_trace);
  return f;
}
static int parse_aux_spec(table_t *t, const uint64_t *w, size_t n) {
  rd_t r = {w, n, 0, 0};
  if (rd(&r) != AUXSPEC_MAGIC) return -1;
  t->n_lookups = (int)rd(&r); t->n_zs = (int)rd(&r);
  if (r.bad || t->n_lookups < 0 || t->n_lookups > 256 || t->n_zs < 0 || t->n_zs > 256) return -1;
  t->lookups = (lookup_t *)calloc(t->n_lookups + 1, sizeof(lookup_t));
  t->zs = (ctlz_t *)calloc(t->n_zs + 1, sizeof(ctlz_t));
  for (int i = 0; i < t->n_lookups && !r.bad; i++) {
    lookup_t *l = &t->lookups[i];
    l->n_cols = (int)rd(&r);
    if (l->n_cols < 1 || l->n_cols > 65536) return -1;
    l->cols = (column_t *)calloc(l->n_cols, sizeof(column_t)); l->filters = (filter_t *)calloc(l->n_cols, sizeof(filter_t));
    for (int j = 0; j < l->n_cols; j++) l->cols[j] = rd_column(&r, t->cols);
    for (int j = 0; j < l->n_cols; j++) l->filters[j] = rd_filter(&r, t->cols);
    l->table = rd_column(&r, t->cols); l->freq = rd_column(&r, t->cols);
  }
  for (int i = 0; i < t->n_zs && !r.bad; i++) {
    ctlz_t *z = &t->zs[i];
    z->challenge = (int)rd(&r); z->n_sets = (int)rd(&r);
    if (z->challenge < 0 || z->challenge >= NUM_CHALLENGES || z->n_sets < 1 || z->n_sets > 4096) return -1;
    z->sets = (colset_t *)calloc(z->n_sets, sizeof(colset_t));
    for (int s = 0; s < z->n_sets && !r.bad; s++) {
      colset_t *cs = &z->sets[s];
      cs->n_cols = (int)rd(&r);
      if (cs->n_cols < 1 || cs->n_cols > 65536) return -1;
      cs->cols = (column_t *)calloc(cs->n_cols, sizeof(column_t));
      for (int j = 0; j < cs->n_cols; j++) cs->cols[j] = rd_column(&r, t->cols);
      cs->filter = rd_filter(&r, t->cols);
    }
  }
  return (r.bad || r.pos != n) ? -1 : 0;
}
static int register_program(const uint64_t *program, size_t n_words) {
  get_table(0);
  if (n_words < 8 || program[0] != CPROG_MAGIC || n_words != 8 + 2 * program[1]) return -1;
  int id = -1;
  for (int i = 16; i < ORC_MAX_TABLES; i++) if (!g_tables[i].used) { id = i; break; }
  if (id < 0) return -1;
  table_t *t = &g_tables[id];
  memset(t, 0, sizeof *t);
  t->n_ops = (uint32_t)program[1]; t->cols = (int)program[2]; t->n_aux = (uint32_t)program[3]; t->n_pi = (int)program[4];
  t->n_ch = (uint32_t)program[5]; t->degree = (int)program[6];
  t->ops = (uint64_t *)malloc(2 * (size_t)t->n_ops * 8 + 8);
  memcpy(t->ops, program + 8, 2 * (size_t)t->n_ops * 8);
  return id;
}
/* lookups: [n_lookups, then per lookup: table_col, freq_col, n_looking, looking...] (plain columns, default filters) */
int orc_table_register(const uint64_t *program, size_t n_words, const int32_t *lookups, size_t n_lookup_words) {
  int id = register_program(program, n_words);
  if (id < 0) return -1;
  table_t *t = &g_tables[id];
  if (n_lookup_words) {
    size_t pos = 0;
    t->n_lookups = lookups[pos++];
    t->lookups = (lookup_t *)calloc(t->n_lookups + 1, sizeof(lookup_t));
    for (int i = 0; i < t->n_lookups; i++) {
      lookup_t *l = &t->lookups[i];
      l->table = column_single(lookups[pos++]); l->freq = column_single(lookups[pos++]); l->n_cols = lookups[pos++];
      l->cols = (column_t *)calloc(l->n_cols, sizeof(column_t)); l->filters = (filter_t *)calloc(l->n_cols, sizeof(filter_t));
      for (int j = 0; j < l->n_cols; j++) { l->cols[j] = column_single(lookups[pos++]); l->filters[j] = filter_default(); }
    }
  }
  t->used = 1;
  return id;
}
/* general form: lookups with linear-combination Columns and Filters, and the table's CTL Z descriptors */
int orc_table_register_ex(const uint64_t *program, size_t n_words, const uint64_t *aux_spec, size_t n_spec_words) {
  int id = register_program(program, n_words);
  if (id < 0) return -1;
  table_t *t = &g_tables[id];
  if (n_spec_words && parse_aux_spec(t, aux_spec, n_spec_words) != 0) return -1;
  t->used = 1;
  return id;
}
int orc_table_num_columns(int t) { return get_table(t)->cols; }
int orc_table_constraint_degree(int t) { return get_table(t)->degree; }
int orc_table_num_public_inputs(int t) { return get_table(t)->n_pi; }
int orc_table_uses_lookup(int t) { return get_table(t)->n_lookups > 0; }
int orc_table_requires_ctls(int t) { return get_table(t)->n_zs > 0; }
static int lookup_chunk(const table_t *t) { return t->degree - 1 < 1 ? 1 : t->degree - 1; }
static int lookup_helpers(const table_t *t, const lookup_t *l) { return (l->n_cols + lookup_chunk(t) - 1) / lookup_chunk(t); }
static int ctl_helpers(const table_t *t, const ctlz_t *z) { return z->n_sets > 1 ? (z->n_sets + lookup_chunk(t) - 1) / lookup_chunk(t) : 0; }
static int quotient_degree_factor(int t) {
  int d = orc_table_constraint_degree(t) - 1;
  return d < 1 ? 1 : d;
}
static int log2_ceil(int x) { int l = 0; while ((1 << l) < x) l++; return l; }
/* Lookup::num_helper_columns summed over lookups and challenges */
int orc_table_num_lookup_columns(int t, int n_challenges) {
  const table_t *tb = get_table(t);
  int a = 0;
  for (int i = 0; i < tb->n_lookups; i++) a += lookup_helpers(tb, &tb->lookups[i]) + 1;
  return a * n_challenges;
}
int orc_table_num_ctl_helper_columns(int t) {
  const table_t *tb = get_table(t);
  int a = 0;
  for (int i = 0; i < tb->n_zs; i++) a += ctl_helpers(tb, &tb->zs[i]);
  return a;
}
int orc_table_num_ctl_zs(int t) { return get_table(t)->n_zs; }
/* all auxiliary polynomials: lookup columns ++ CTL helper columns ++ CTL Z columns (starky prover.rs) */
int orc_table_num_aux_columns(int t, int n_challenges) {
  return orc_table_num_lookup_columns(t, n_challenges) + orc_table_num_ctl_helper_columns(t) + orc_table_num_ctl_zs(t);
}

enum { M_FILTER = 0, M_TIMESTAMP, M_IS_READ, M_CTX, M_SEG, M_VIRT, M_VALUE0, M_CFC = 14, M_SFC, M_VFC,
       M_INIT_AUX, M_RANGE_CHECK, M_COUNTER, M_FREQ };

typedef struct {
  uint64_t alphas[NUM_CHALLENGES], acc[NUM_CHALLENGES];
  int n;
  uint64_t z_last, lagrange_first, lagrange_last;
  long count;      /* running constraint index (check mode)            */
  long first_fail; /* check mode: first non-zero constraint, else -1   */
  int check;
} consumer_t;
static inline void c_constraint(consumer_t *c, uint64_t v) {
  if (c->check) { if (gl_canon(v) != 0 && c->first_fail < 0) c->first_fail = c->count; c->count++; return; }
  for (int j = 0; j < c->n; j++) c->acc[j] = gl_add(gl_mul(c->acc[j], c->alphas[j]), v);
}
static inline void c_transition(consumer_t *c, uint64_t v) { c_constraint(c, gl_mul(v, c->z_last)); }
static inline void c_first_row(consumer_t *c, uint64_t v) { c_constraint(c, gl_mul(v, c->lagrange_first)); }
static inline void c_last_row(consumer_t *c, uint64_t v) { c_constraint(c, gl_mul(v, c->lagrange_last)); }

static void eval_fibonacci(const uint64_t *lv, const uint64_t *nv, const uint64_t *pi, consumer_t *c) {
  c_first_row(c, gl_sub(lv[0], pi[0]));
  c_first_row(c, gl_sub(lv[1], pi[1]));
  c_last_row(c, gl_sub(lv[1], pi[2]));
  c_transition(c, gl_sub(nv[0], lv[1]));
  c_transition(c, gl_sub(gl_sub(nv[1], lv[0]), lv[1]));
}

static void eval_memory(const uint64_t *lv, const uint64_t *nv, const uint64_t *pi, consumer_t *c) {
  (void)pi;
  const uint64_t one = 1;
  uint64_t filter = lv[M_FILTER];
  c_constraint(c, gl_mul(filter, gl_sub(filter, one)));
  c_constraint(c, gl_mul(gl_sub(one, filter), gl_sub(one, lv[M_IS_READ])));
  uint64_t cfc = lv[M_CFC], sfc = lv[M_SFC], vfc = lv[M_VFC];
  uint64_t unchanged = gl_sub(gl_sub(gl_sub(one, cfc), sfc), vfc);
  c_constraint(c, gl_mul(cfc, gl_sub(one, cfc)));
  c_constraint(c, gl_mul(sfc, gl_sub(one, sfc)));
  c_constraint(c, gl_mul(vfc, gl_sub(one, vfc)));
  c_constraint(c, gl_mul(unchanged, gl_sub(one, unchanged)));
  uint64_t d_ctx = gl_sub(nv[M_CTX], lv[M_CTX]), d_seg = gl_sub(nv[M_SEG], lv[M_SEG]);
  uint64_t d_virt = gl_sub(nv[M_VIRT], lv[M_VIRT]), d_ts = gl_sub(nv[M_TIMESTAMP], lv[M_TIMESTAMP]);
  c_transition(c, gl_mul(sfc, d_ctx));
  c_transition(c, gl_mul(vfc, d_ctx));
  c_transition(c, gl_mul(vfc, d_seg));
  c_transition(c, gl_mul(unchanged, d_ctx));
  c_transition(c, gl_mul(unchanged, d_seg));
  c_transition(c, gl_mul(unchanged, d_virt));
  uint64_t computed = gl_add(gl_add(gl_mul(cfc, gl_sub(d_ctx, one)), gl_mul(sfc, gl_sub(d_seg, one))),
                             gl_add(gl_mul(vfc, gl_sub(d_virt, one)), gl_mul(unchanged, d_ts)));
  c_transition(c, gl_sub(lv[M_RANGE_CHECK], computed));
  uint64_t init_aux = lv[M_INIT_AUX];
  c_transition(c, gl_sub(init_aux, gl_mul(gl_mul(nv[M_SEG], gl_sub(one, unchanged)), nv[M_IS_READ])));
  for (int i = 0; i < 8; i++) {
    uint64_t v = lv[M_VALUE0 + i], nvv = nv[M_VALUE0 + i];
    c_transition(c, gl_mul(gl_mul(nv[M_IS_READ], unchanged), gl_sub(nvv, v)));
    c_transition(c, gl_mul(gl_mul(nv[M_CTX], init_aux), nvv));
    c_transition(c, gl_mul(gl_mul(gl_sub(nv[M_SEG], MEM_TRIE_DATA_SEGMENT), init_aux), nvv));
  }
  c_first_row(c, lv[M_COUNTER]);
  c_transition(c, gl_sub(gl_sub(nv[M_COUNTER], lv[M_COUNTER]), one));
}

/* constraint program interpreter: own constraints, lookup checks AND CTL checks are all in the program */
static void eval_program(const table_t *tb, const uint64_t *lv, const uint64_t *nv, const uint64_t *al, const uint64_t *an,
                         const uint64_t *pi, const uint64_t *ch, consumer_t *c) {
  uint64_t *v = (uint64_t *)malloc((size_t)tb->n_ops * 8 + 8);
  for (uint32_t k = 0; k < tb->n_ops; k++) {
    const uint64_t w0 = tb->ops[2 * k], imm = tb->ops[2 * k + 1];
    const int op = (int)(w0 & 0xFF);
    const uint32_t a = (uint32_t)((w0 >> 8) & 0xFFFFFFF), b = (uint32_t)((w0 >> 36) & 0xFFFFFFF);
    switch (op) {
      case OP_CONST: v[k] = imm; break;
      case OP_LV: v[k] = lv[a]; break;
      case OP_NV: v[k] = nv[a]; break;
      case OP_LA: v[k] = al ? al[a] : 0; break;
      case OP_NA: v[k] = an ? an[a] : 0; break;
      case OP_PI: v[k] = pi[a]; break;
      case OP_CH: v[k] = ch ? ch[a] : 0; break;
      case OP_ADD: v[k] = gl_add(v[a], v[b]); break;
      case OP_SUB: v[k] = gl_sub(v[a], v[b]); break;
      case OP_MUL: v[k] = gl_mul(v[a], v[b]); break;
      case OP_EMIT: c_constraint(c, v[a]); break;
      case OP_EMIT_TRANSITION: c_transition(c, v[a]); break;
      case OP_EMIT_FIRST: c_first_row(c, v[a]); break;
      case OP_EMIT_LAST: c_last_row(c, v[a]); break;
      default: break;
    }
  }
  free(v);
}

static void eval_table(int t, const uint64_t *lv, const uint64_t *nv, const uint64_t *pi, consumer_t *c) {
  if (t == ORC_TABLE_FIBONACCI) eval_fibonacci(lv, nv, pi, c);
  else eval_memory(lv, nv, pi, c);
}

/* Column::eval_with_next / Filter::eval_filter on one (local, next) row pair (starky/src/lookup.rs) */
static uint64_t column_eval(const column_t *k, const uint64_t *lv, const uint64_t *nv) {
  uint64_t acc = k->constant;
  for (int i = 0; i < k->n_local; i++) acc = gl_add(acc, gl_mul(lv[k->lcol[i]], k->lcoef[i]));
  for (int i = 0; i < k->n_next; i++) acc = gl_add(acc, gl_mul(nv[k->ncol[i]], k->ncoef[i]));
  return acc;
}
static uint64_t filter_eval(const filter_t *f, const uint64_t *lv, const uint64_t *nv) {
  uint64_t acc = 0;
  for (int i = 0; i < f->n_prod; i++) acc = gl_add(acc, gl_mul(column_eval(&f->pa[i], lv, nv), column_eval(&f->pb[i], lv, nv)));
  for (int i = 0; i < f->n_const; i++) acc = gl_add(acc, column_eval(&f->consts[i], lv, nv));
  return acc;
}
/* Column::eval_table / eval_all_rows: row i of a column-major trace, next row cyclic */
static uint64_t column_eval_table(const column_t *k, const uint64_t *trace, size_t n, size_t i) {
  uint64_t acc = k->constant;
  for (int j = 0; j < k->n_local; j++) acc = gl_add(acc, gl_mul(trace[(size_t)k->lcol[j] * n + i], k->lcoef[j]));
  for (int j = 0; j < k->n_next; j++) acc = gl_add(acc, gl_mul(trace[(size_t)k->ncol[j] * n + (i + 1) % n], k->ncoef[j]));
  return acc;
}
static uint64_t filter_eval_table(const filter_t *f, const uint64_t *trace, size_t n, size_t i) {
  uint64_t acc = 0;
  for (int j = 0; j < f->n_prod; j++) acc = gl_add(acc, gl_mul(column_eval_table(&f->pa[j], trace, n, i), column_eval_table(&f->pb[j], trace, n, i)));
  for (int j = 0; j < f->n_const; j++) acc = gl_add(acc, column_eval_table(&f->consts[j], trace, n, i));
  return acc;
}

/* eval_packed_lookups_generic for the BUILT-IN tables (programs carry these checks themselves): per lookup, per
 * challenge: helpers..., Z; eval_helper_columns with chunks of 1 or 2 and filters */
static void eval_lookups(int t, const uint64_t *lv, const uint64_t *nv, const uint64_t *aux_l, const uint64_t *aux_n,
                         const uint64_t *challenges, int n_ch, consumer_t *c) {
  const table_t *tb = get_table(t);
  const int chunk = lookup_chunk(tb);
  int start = 0;
  for (int li = 0; li < tb->n_lookups; li++) {
    const lookup_t *l = &tb->lookups[li];
    const int nh = lookup_helpers(tb, l);
    for (int k = 0; k < n_ch; k++) {
      uint64_t ch = challenges[k];
      uint64_t hsum = 0;
      for (int hc = 0; hc < nh; hc++) {
        int j0 = hc * chunk, j1 = j0 + chunk < l->n_cols ? j0 + chunk : l->n_cols;
        uint64_t h = aux_l[start + hc];
        if (j1 - j0 == 2) {
          uint64_t c0 = gl_add(column_eval(&l->cols[j0], lv, nv), ch), c1 = gl_add(column_eval(&l->cols[j0 + 1], lv, nv), ch);
          uint64_t f0 = filter_eval(&l->filters[j0], lv, nv), f1 = filter_eval(&l->filters[j0 + 1], lv, nv);
          /* combin1 * combin0 * h - f0 * combin1 - f1 * combin0 */
          c_constraint(c, gl_sub(gl_sub(gl_mul(gl_mul(c1, c0), h), gl_mul(f0, c1)), gl_mul(f1, c0)));
        } else {
          uint64_t c0 = gl_add(column_eval(&l->cols[j0], lv, nv), ch);
          c_constraint(c, gl_sub(gl_mul(c0, h), filter_eval(&l->filters[j0], lv, nv)));
        }
        hsum = gl_add(hsum, h);
      }
      uint64_t z = aux_l[start + nh], next_z = aux_n[start + nh];
      uint64_t table_with_challenge = gl_add(column_eval(&l->table, lv, nv), ch);
      uint64_t y = gl_sub(gl_mul(hsum, table_with_challenge), column_eval(&l->freq, lv, nv));
      c_first_row(c, z);
      c_constraint(c, gl_sub(gl_mul(gl_sub(next_z, z), table_with_challenge), y));
      start += nh + 1;
    }
  }
}

/* check_constraints analogue on the trace domain (no aux); returns -1 or row*1000 + constraint idx */
long orc_table_check_constraints(int t, int log_n, const uint64_t *trace, const uint64_t *pi) {
  size_t n = (size_t)1 << log_n;
  int nc = orc_table_num_columns(t);
  const table_t *tb = get_table(t);
  if (tb->ops) return -2; /* programs: use cprog.Program.check_trace (the lookup checks need the aux columns) */
  uint64_t *lv = (uint64_t *)malloc(2 * (size_t)nc * 8), *nv = lv + nc;
  long res = -1;
  for (size_t i = 0; i < n && res < 0; i++) {
    for (int c = 0; c < nc; c++) { lv[c] = trace[c * n + i]; nv[c] = trace[c * n + (i + 1) % n]; }
    consumer_t cs; memset(&cs, 0, sizeof cs);
    cs.check = 1; cs.first_fail = -1;
    cs.z_last = (i == n - 1) ? 0 : 1; cs.lagrange_first = (i == 0); cs.lagrange_last = (i == n - 1);
    eval_table(t, lv, nv, pi, &cs);
    if (cs.first_fail >= 0) res = (long)i * 1000 + cs.first_fail;
  }
  free(lv);
  return res;
}

/* ---------------- auxiliary columns ---------------- */
/* F::batch_multiplicative_inverse (Montgomery's trick); upstream panics on a zero ("Tried to invert zero"): returns 1 */
static int batch_inverse(uint64_t *x, size_t n) {
  uint64_t *pre = (uint64_t *)malloc(n * sizeof(uint64_t));
  uint64_t acc = 1;
  for (size_t i = 0; i < n; i++) { pre[i] = acc; acc = gl_mul(acc, x[i]); }
  if (gl_canon(acc) == 0) { free(pre); return 1; }
  uint64_t inv = gl_inv(acc);
  for (size_t i = n; i-- > 0;) { uint64_t xi = x[i]; x[i] = gl_mul(inv, pre[i]); inv = gl_mul(inv, xi); }
  free(pre);
  return 0;
}
/* starky/src/lookup.rs get_helper_cols on one chunk: sum over the chunk's (columns, filter) pairs of
 * filter(i) / combine(columns)(i), combine = reduce_with_powers(evals, beta) + gamma */
static int helper_column(const uint64_t *trace, size_t n, int n_sets, column_t *const *set_cols, const int *set_ncols,
                         const filter_t *const *filters, uint64_t beta, uint64_t gamma, uint64_t *out) {
  uint64_t *inv = (uint64_t *)malloc(n * 8);
  memset(out, 0, n * 8);
  int bad = 0;
  for (int s = 0; s < n_sets; s++) {
    for (size_t i = 0; i < n; i++) {
      uint64_t acc = 0;
      for (int j = set_ncols[s]; j-- > 0;) acc = gl_add(gl_mul(acc, beta), column_eval_table(&set_cols[s][j], trace, n, i));
      inv[i] = gl_add(acc, gamma);
    }
    bad |= batch_inverse(inv, n);
    for (size_t i = 0; i < n; i++) out[i] = gl_add(out[i], gl_mul(inv[i], filter_eval_table(filters[s], trace, n, i)));
  }
  free(inv);
  return bad;
}
/* All auxiliary polynomials of one table (values on the trace domain, column-major):
 *   lookup_helper_columns for every lookup and challenge (starky/src/lookup.rs), then the CTL helper columns of every
 *   CtlZData, then the CTL Z columns (starky/src/cross_table_lookup.rs partial_sums / get_ctl_auxiliary_polys).
 * lookup_ch: n_ch scalars; ctl_ch: n_ch (beta, gamma) pairs (may be NULL when the table has no CTL).  Returns 0, or 1 when a
 * denominator was zero (upstream panics). */
int orc_aux_columns(int t, int log_n, const uint64_t *trace, const uint64_t *lookup_ch, int n_ch, const uint64_t *ctl_ch,
                    uint64_t *aux) {
  const table_t *tb = get_table(t);
  size_t n = (size_t)1 << log_n;
  const int chunk = lookup_chunk(tb);
  int bad = 0;
  uint64_t *out = aux;
  uint64_t *tinv = (uint64_t *)malloc(n * 8);
  for (int li = 0; li < tb->n_lookups; li++) {
    const lookup_t *l = &tb->lookups[li];
    const int nh = lookup_helpers(tb, l);
    for (int k = 0; k < n_ch; k++) {
      uint64_t ch = lookup_ch[k];
      uint64_t *z = out + (size_t)nh * n;
      for (int hc = 0; hc < nh; hc++) {
        int j0 = hc * chunk, j1 = j0 + chunk < l->n_cols ? j0 + chunk : l->n_cols;
        column_t *sc[8]; int sn[8]; const filter_t *sf[8];
        for (int j = j0; j < j1; j++) { sc[j - j0] = &l->cols[j]; sn[j - j0] = 1; sf[j - j0] = &l->filters[j]; }
        /* GrandProductChallenge { beta: 1, gamma: challenge } */
        bad |= helper_column(trace, n, j1 - j0, sc, sn, sf, 1, ch, out + (size_t)hc * n);
        for (size_t i = 0; i < n; i++) out[(size_t)hc * n + i] = gl_canon(out[(size_t)hc * n + i]);
      }
      for (size_t i = 0; i < n; i++) tinv[i] = gl_add(column_eval_table(&l->table, trace, n, i), ch);
      bad |= batch_inverse(tinv, n);
      z[0] = 0;
      for (size_t i = 0; i + 1 < n; i++) {
        uint64_t tot = 0;
        for (int hc = 0; hc < nh; hc++) tot = gl_add(tot, out[(size_t)hc * n + i]);
        z[i + 1] = gl_canon(gl_add(z[i], gl_sub(tot, gl_mul(column_eval_table(&l->freq, trace, n, i), tinv[i]))));
      }
      out += (size_t)(nh + 1) * n;
    }
  }
  free(tinv);
  /* CTL: helper columns of every Z first, then the Zs */
  uint64_t *helpers = out, *zs = out + (size_t)orc_table_num_ctl_helper_columns(t) * n;
  uint64_t *hsum = (uint64_t *)malloc(n * 8), *tmp = (uint64_t *)malloc(n * 8);
  for (int zi = 0; zi < tb->n_zs; zi++) {
    const ctlz_t *z = &tb->zs[zi];
    const uint64_t beta = ctl_ch[2 * z->challenge], gamma = ctl_ch[2 * z->challenge + 1];
    const int nh = ctl_helpers(tb, z);
    memset(hsum, 0, n * 8);
    const int n_chunks = (z->n_sets + chunk - 1) / chunk;
    for (int hc = 0; hc < n_chunks; hc++) {
      int s0 = hc * chunk, s1 = s0 + chunk < z->n_sets ? s0 + chunk : z->n_sets;
      column_t *sc[8]; int sn[8]; const filter_t *sf[8];
      for (int s = s0; s < s1; s++) { sc[s - s0] = z->sets[s].cols; sn[s - s0] = z->sets[s].n_cols; sf[s - s0] = &z->sets[s].filter; }
      uint64_t *dst = nh ? helpers + (size_t)hc * n : tmp;
      bad |= helper_column(trace, n, s1 - s0, sc, sn, sf, beta, gamma, dst);
      for (size_t i = 0; i < n; i++) { dst[i] = gl_canon(dst[i]); hsum[i] = gl_add(hsum[i], dst[i]); }
    }
    helpers += (size_t)nh * n;
    /* partial_sums: z[n-1] = hsum[n-1]; z[i] = z[i+1] + hsum[i] */
    uint64_t *zc = zs + (size_t)zi * n;
    zc[n - 1] = gl_canon(hsum[n - 1]);
    for (size_t i = n - 1; i-- > 0;) zc[i] = gl_add(zc[i + 1], hsum[i]);
  }
  free(hsum); free(tmp);
  return bad;
}
/* round-1 name: lookups only */
void orc_lookup_helper_columns(int t, int log_n, const uint64_t *trace, const uint64_t *challenges, int n_ch, uint64_t *aux) {
  uint64_t zero[2 * NUM_CHALLENGES] = {0};
  orc_aux_columns(t, log_n, trace, challenges, n_ch, zero, aux);
}

/* ---------------- compute_quotient_polys ---------------- */
/* ch: challenge scalars: lookup challenges [0 .. NUM_CHALLENGES) then the CTL (beta, gamma) pairs */
void orc_compute_quotient_polys(int t, int log_n, const orc_batch *trace, const orc_batch *aux,
                                const uint64_t *ch, const uint64_t *pi,
                                const uint64_t *alphas, int n_alphas, uint64_t *quotient_chunks) {
  size_t degree = (size_t)1 << log_n;
  const int rate_bits = trace->rate_bits;
  int qbits = log2_ceil(quotient_degree_factor(t));
  int step = 1 << (rate_bits - qbits), next_step = 1 << qbits;
  size_t size = degree << qbits;
  int log_size = log_n + qbits, log_lde = log_n + rate_bits;
  int nc = orc_table_num_columns(t), na = aux ? (int)aux->n_cols : 0;
  /* Lagrange selectors on the coset: selector(degree, i).lde_onto_coset(qbits) */
  uint64_t *lfirst = (uint64_t *)calloc(size, 8), *llast = (uint64_t *)calloc(size, 8);
  uint64_t *tmp = (uint64_t *)calloc(degree, 8);
  tmp[0] = 1; orc_ifft(tmp, log_n); memcpy(lfirst, tmp, degree * 8); orc_coset_fft(lfirst, log_size, GL_GENERATOR);
  memset(tmp, 0, degree * 8);
  tmp[degree - 1] = 1; orc_ifft(tmp, log_n); memcpy(llast, tmp, degree * 8); orc_coset_fft(llast, log_size, GL_GENERATOR);
  free(tmp);
  /* ZeroPolyOnCoset::new(degree_bits, qbits): Z_H(x_i) = g^n * w_{2^qbits}^(i mod 2^qbits) - 1 */
  uint64_t zh_inv[16];
  {
    uint64_t g_pow_n = GL_GENERATOR;
    for (int i = 0; i < log_n; i++) g_pow_n = gl_sqr(g_pow_n);
    uint64_t w = gl_root_of_unity(qbits), cur = 1;
    for (int i = 0; i < (1 << qbits); i++) { zh_inv[i] = gl_inv(gl_sub(gl_mul(g_pow_n, cur), 1)); cur = gl_mul(cur, w); }
  }
  uint64_t last = gl_inv(gl_root_of_unity(log_n));
  uint64_t w_size = gl_root_of_unity(log_size);
  /* F::cyclic_subgroup_coset_known_order(w_size, coset_shift, size) */
  uint64_t *coset = (uint64_t *)malloc(size * 8);
  { uint64_t cur = GL_GENERATOR; for (size_t i = 0; i < size; i++) { coset[i] = cur; cur = gl_mul(cur, w_size); } }
  uint64_t *qvals = (uint64_t *)malloc((size_t)n_alphas * size * 8);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < size; i++) {
    size_t i_next = (i + next_step) % size;
    uint64_t x = coset[i];
    consumer_t cs; memset(&cs, 0, sizeof cs);
    cs.n = n_alphas;
    for (int j = 0; j < n_alphas; j++) cs.alphas[j] = alphas[j];
    cs.z_last = gl_sub(x, last); cs.lagrange_first = lfirst[i]; cs.lagrange_last = llast[i];
    const uint64_t *lv = trace->leaves + bitrev64(i * step, log_lde) * nc;
    const uint64_t *nv = trace->leaves + bitrev64(i_next * step, log_lde) * nc;
    const uint64_t *al = aux ? aux->leaves + bitrev64(i * step, log_lde) * na : NULL;
    const uint64_t *an = aux ? aux->leaves + bitrev64(i_next * step, log_lde) * na : NULL;
    if (get_table(t)->ops) {
      eval_program(get_table(t), lv, nv, al, an, pi, ch, &cs);
    } else {
      eval_table(t, lv, nv, pi, &cs);
      if (aux) eval_lookups(t, lv, nv, al, an, ch, NUM_CHALLENGES, &cs);
    }
    uint64_t dinv = zh_inv[i % (1 << qbits)];
    for (int j = 0; j < n_alphas; j++) qvals[(size_t)j * size + i] = gl_mul(cs.acc[j], dinv);
  }
  free(lfirst); free(llast); free(coset);
  /* coset_ifft(7) each; split into quotient_degree_factor chunks of `degree` coefficients */
  int factor = quotient_degree_factor(t);
  for (int j = 0; j < n_alphas; j++) {
    uint64_t *q = qvals + (size_t)j * size;
    orc_coset_ifft(q, log_size, GL_GENERATOR);
    /* trim_to_len(degree * factor): anything above must be zero */
    for (size_t k = degree * factor; k < size; k++)
      if (q[k] != 0) fprintf(stderr, "oracle: quotient not divisible by Z_H (table %d)\n", t);
    memcpy(quotient_chunks + (size_t)j * factor * degree, q, degree * factor * 8);
  }
  free(qvals);
}

/* ---------------- FRI ---------------- */
/* reduce_with_powers over chunks of 2^arity_bits ext coefficients (interleaved c0,c1) */
void orc_fri_fold_coeffs(const uint64_t *coeffs, size_t n, int arity_bits, const uint64_t beta[2], uint64_t *out) {
  size_t arity = (size_t)1 << arity_bits;
  gl2_t b = gl2(beta[0], beta[1]);
  for (size_t k = 0; k < n / arity; k++) {
    gl2_t acc = gl2(0, 0);
    for (size_t j = arity; j-- > 0;) {
      gl2_t c = gl2(coeffs[2 * (k * arity + j)], coeffs[2 * (k * arity + j) + 1]);
      acc = gl2_add(gl2_mul(acc, b), c);
    }
    out[2 * k] = acc.c0; out[2 * k + 1] = acc.c1;
  }
}
/* ext coset FFT: the DFT is F-linear with base-field twiddles, so transform c0 and c1 separately */
static void ext_coset_fft(uint64_t *v /* interleaved */, int log_n, uint64_t shift) {
  size_t n = (size_t)1 << log_n;
  uint64_t *a = (uint64_t *)malloc(n * 8), *b = (uint64_t *)malloc(n * 8);
  for (size_t i = 0; i < n; i++) { a[i] = v[2 * i]; b[i] = v[2 * i + 1]; }
  orc_coset_fft(a, log_n, shift); orc_coset_fft(b, log_n, shift);
  for (size_t i = 0; i < n; i++) { v[2 * i] = a[i]; v[2 * i + 1] = b[i]; }
  free(a); free(b);
}

/* fri_proof_of_work: smallest candidate whose response has >= bits leading zeros */
uint64_t orc_pow_grind(const uint64_t state[12], int pos, int bits) {
  for (uint64_t cand = 0;; cand++) {
    uint64_t s[12];
    memcpy(s, state, sizeof s);
    s[pos] = cand;
    orc_poseidon_permute(s);
    if (bits == 0 || (s[7] >> (64 - bits)) == 0) return cand;
  }
}

/* evaluate a base-coefficient polynomial at an ext point (to_extension().eval(z): Horner) */
static gl2_t eval_base_poly(const uint64_t *c, size_t n, gl2_t z) {
  gl2_t acc = gl2(0, 0);
  for (size_t i = n; i-- > 0;) acc = gl2_add(gl2_mul(acc, z), gl2_from_base(c[i]));
  return acc;
}
/* PolynomialBatch polynomials evaluated at an extension point: out has n_cols ext values (interleaved) */
void orc_batch_eval_at_ext_point(const orc_batch *b, const uint64_t z[2], uint64_t *out) {
  size_t n = (size_t)1 << b->log_n;
#pragma omp parallel for schedule(dynamic)
  for (size_t c = 0; c < b->n_cols; c++) {
    gl2_t v = eval_base_poly(b->coeffs + c * n, n, gl2(z[0], z[1]));
    out[2 * c] = v.c0; out[2 * c + 1] = v.c1;
  }
}

/* words of a flat FriProof: commit_phase_merkle_caps, query_round_proofs, final_poly, pow_witness */
size_t orc_fri_proof_words(const size_t *oracle_cols, size_t n_oracles, const orc_fri_params *p) {
  const size_t cap = (size_t)4 << p->cap_height;
  const int log_lde = p->degree_bits + p->rate_bits;
  size_t per_query = 0;
  for (size_t o = 0; o < n_oracles; o++) per_query += oracle_cols[o] + 4 * (size_t)(log_lde - p->cap_height);
  int bits = log_lde;
  for (int l = 0; l < p->n_reductions; l++) {
    bits -= p->reduction_arity_bits[l];
    per_query += 2 * ((size_t)1 << p->reduction_arity_bits[l]) + 4 * (size_t)(bits - p->cap_height);
  }
  const size_t final_len = (size_t)1 << (p->degree_bits - fri_total_arities(p));
  return cap * p->n_reductions + (size_t)p->num_query_rounds * per_query + 2 * final_len + 1;
}

/* PolynomialBatch::prove_openings (plonky2/src/fri/oracle.rs) + fri_proof (plonky2/src/fri/prover.rs) for a general
 * FriInstanceInfo: batches of (point, [(oracle_index, polynomial_index)]).  The challenger is updated in place.
 * Returns 0, 1 "FRI total reduction arity is too large", 3 PoW mismatch, 5 bad instance. */
int orc_prove_openings(const orc_fri_batch *batches, size_t n_batches, const orc_batch *const *oracles, size_t n_oracles,
                       orc_challenger *ch, const orc_fri_params *fp, uint64_t *out) {
  const int log_n = fp->degree_bits, log_lde = log_n + fp->rate_bits;
  const size_t n = (size_t)1 << log_n, lde_n = (size_t)1 << log_lde, cap_words = (size_t)4 << fp->cap_height;
  if (fri_total_arities(fp) > log_n + fp->rate_bits - fp->cap_height) return 1;
  for (size_t o = 0; o < n_oracles; o++)
    if (oracles[o]->log_n != log_n || oracles[o]->rate_bits != fp->rate_bits || oracles[o]->cap_height != fp->cap_height) return 5;
  uint64_t *w = out;
  uint64_t alpha_w[2];
  orc_challenger_get_n(ch, 2, alpha_w);
  gl2_t alpha = gl2(alpha_w[0], alpha_w[1]);
  gl2_t *final_poly = (gl2_t *)calloc(n, sizeof(gl2_t));
  gl2_t *comp = (gl2_t *)malloc(n * sizeof(gl2_t));
  for (size_t b = 0; b < n_batches; b++) {
    const size_t count = batches[b].n_polynomials;
    const gl2_t point = gl2(batches[b].point[0], batches[b].point[1]);
    const uint64_t **polys = (const uint64_t **)malloc((count + 1) * sizeof(*polys));
    for (size_t k = 0; k < count; k++) {
      const orc_fri_poly fpi = batches[b].polynomials[k];
      if (fpi.oracle_index >= n_oracles || fpi.polynomial_index >= oracles[fpi.oracle_index]->n_cols) { free(polys); free(comp); free(final_poly); return 5; }
      polys[k] = oracles[fpi.oracle_index]->coeffs + (size_t)fpi.polynomial_index * n;
    }
    /* reduce_polys_base: sum_k alpha^k * poly_k */
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
      gl2_t acc = gl2(0, 0);
      for (size_t k = count; k-- > 0;) acc = gl2_add(gl2_mul(acc, alpha), gl2_from_base(polys[k][i]));
      comp[i] = acc;
    }
    free(polys);
    /* divide_by_linear(point): synthetic division, result padded back to n with a zero */
    gl2_t *q = (gl2_t *)malloc(n * sizeof(gl2_t));
    gl2_t acc = gl2(0, 0);
    for (size_t i = n; i-- > 0;) {
      acc = gl2_add(gl2_mul(acc, point), comp[i]);
      if (i > 0) q[i - 1] = acc;
    }
    q[n - 1] = gl2(0, 0);
    /* alpha.shift_poly(final_poly); final_poly += quotient */
    gl2_t sh = gl2_pow(alpha, (uint64_t)count);
    for (size_t i = 0; i < n; i++) final_poly[i] = gl2_add(gl2_mul(final_poly[i], sh), q[i]);
    free(q);
  }
  free(comp);

  /* lde(rate_bits) + coset_fft(7) over the extension */
  uint64_t *coeffs = (uint64_t *)calloc(2 * lde_n, 8), *values = (uint64_t *)malloc(2 * lde_n * 8);
  for (size_t i = 0; i < n; i++) { coeffs[2 * i] = final_poly[i].c0; coeffs[2 * i + 1] = final_poly[i].c1; }
  free(final_poly);
  memcpy(values, coeffs, 2 * lde_n * 8);
  ext_coset_fft(values, log_lde, GL_GENERATOR);

  /* fri_committed_trees */
  uint64_t *layer_leaves[16], *layer_digests[16], *layer_caps[16];
  size_t layer_nleaves[16];
  size_t cur_n = lde_n;
  int cur_log = log_lde;
  uint64_t shift = GL_GENERATOR;
  for (int l = 0; l < fp->n_reductions; l++) {
    const int ab = fp->reduction_arity_bits[l];
    size_t arity = (size_t)1 << ab, nl = cur_n / arity;
    uint64_t *leaves = (uint64_t *)malloc(2 * cur_n * 8);
    for (size_t i = 0; i < cur_n; i++) { /* reverse_index_bits_in_place then chunk + flatten */
      size_t src = bitrev64(i, cur_log);
      leaves[2 * i] = values[2 * src]; leaves[2 * i + 1] = values[2 * src + 1];
    }
    size_t nd = 2 * (nl - ((size_t)1 << fp->cap_height));
    layer_leaves[l] = leaves; layer_nleaves[l] = nl;
    layer_digests[l] = (uint64_t *)malloc((nd ? nd : 1) * 32);
    layer_caps[l] = (uint64_t *)malloc(cap_words * 8);
    orc_merkle_new(leaves, nl, 2 * arity, fp->cap_height, layer_digests[l], layer_caps[l]);
    orc_challenger_observe(ch, layer_caps[l], cap_words);
    memcpy(w, layer_caps[l], cap_words * 8); w += cap_words;
    uint64_t beta[2];
    orc_challenger_get_n(ch, 2, beta);
    uint64_t *folded = (uint64_t *)malloc(2 * nl * 8);
    orc_fri_fold_coeffs(coeffs, cur_n, ab, beta, folded);
    free(coeffs); coeffs = folded;
    shift = gl_pow(shift, arity);
    cur_n = nl; cur_log -= ab;
    memcpy(values, coeffs, 2 * cur_n * 8);
    ext_coset_fft(values, cur_log, shift);
  }
  /* final poly: drop the (zero) top rate_bits part, observe */
  size_t final_len = cur_n >> fp->rate_bits;
  for (size_t i = final_len; i < cur_n; i++)
    if (coeffs[2 * i] || coeffs[2 * i + 1]) fprintf(stderr, "oracle: FRI final poly high part non-zero\n");
  orc_challenger_observe(ch, coeffs, 2 * final_len);

  /* fri_proof_of_work (input_buffer is overwritten into the sponge, candidate goes at n_in) */
  uint64_t st[12];
  memcpy(st, ch->state, sizeof st);
  for (int i = 0; i < ch->n_in; i++) st[i] = ch->in[i];
  uint64_t pow_witness = orc_pow_grind(st, ch->n_in, fp->proof_of_work_bits);
  orc_challenger_observe(ch, &pow_witness, 1);
  uint64_t pow_response = orc_challenger_get(ch);
  int rc = 0;
  if (fp->proof_of_work_bits && (pow_response >> (64 - fp->proof_of_work_bits)) != 0) rc = 3;

  /* fri_prover_query_rounds */
  uint64_t *qidx = (uint64_t *)malloc(8 * (size_t)fp->num_query_rounds);
  orc_challenger_get_n(ch, fp->num_query_rounds, qidx);
  for (int q = 0; q < fp->num_query_rounds; q++) {
    size_t x = (size_t)(qidx[q] % lde_n);
    for (size_t o = 0; o < n_oracles; o++) {
      const orc_batch *b = oracles[o];
      memcpy(w, b->leaves + x * b->n_cols, b->n_cols * 8); w += b->n_cols;
      orc_merkle_prove(b->digests, lde_n, fp->cap_height, x, w); w += 4 * (log_lde - fp->cap_height);
    }
    int bits = log_lde;
    for (int l = 0; l < fp->n_reductions; l++) {
      const int ab = fp->reduction_arity_bits[l];
      size_t arity = (size_t)1 << ab;
      x >>= ab; bits -= ab;
      memcpy(w, layer_leaves[l] + x * 2 * arity, 2 * arity * 8); w += 2 * arity;
      orc_merkle_prove(layer_digests[l], layer_nleaves[l], fp->cap_height, x, w); w += 4 * (bits - fp->cap_height);
    }
  }
  free(qidx);
  memcpy(w, coeffs, 2 * final_len * 8); w += 2 * final_len;
  *w++ = pow_witness;
  for (int l = 0; l < fp->n_reductions; l++) { free(layer_leaves[l]); free(layer_digests[l]); free(layer_caps[l]); }
  free(coeffs); free(values);
  return rc;
}

/* ---------------- proof sizes ---------------- */
typedef struct {
  int table, log_n, n_trace, n_aux, n_quot, n_layers, final_len, n_pi, n_lookup_cols, n_ctl_helpers, n_ctl_zs;
} shape_t;
static shape_t shape_of(int t, int log_n) {
  shape_t s;
  orc_fri_params fp; orc_fri_params_standard_fast(log_n, &fp);
  s.table = t; s.log_n = log_n; s.n_trace = orc_table_num_columns(t);
  s.n_lookup_cols = orc_table_num_lookup_columns(t, NUM_CHALLENGES);
  s.n_ctl_helpers = orc_table_num_ctl_helper_columns(t); s.n_ctl_zs = orc_table_num_ctl_zs(t);
  s.n_aux = s.n_lookup_cols + s.n_ctl_helpers + s.n_ctl_zs;
  s.n_quot = quotient_degree_factor(t) * NUM_CHALLENGES;
  s.n_layers = fp.n_reductions;
  s.final_len = 1 << (log_n - fri_total_arities(&fp));
  s.n_pi = orc_table_num_public_inputs(t);
  return s;
}
size_t orc_stark_proof_words(int t, int log_n) {
  shape_t s = shape_of(t, log_n);
  orc_fri_params fp; orc_fri_params_standard_fast(log_n, &fp);
  size_t cap = (size_t)4 << fp.cap_height, w = HEADER_WORDS;
  w += cap * (2 + (s.n_aux ? 1 : 0));
  w += 2 * (size_t)(2 * s.n_trace + 2 * s.n_aux + s.n_quot) + s.n_ctl_zs;
  size_t oc[3]; size_t no = 0;
  oc[no++] = s.n_trace; if (s.n_aux) oc[no++] = s.n_aux; oc[no++] = s.n_quot;
  w += orc_fri_proof_words(oc, no, &fp);
  w += s.n_pi;
  return w;
}

/* starky::prover::prove_with_commitment under standard_fast_config.  trace_vals: the trace (column-major), trace: its
 * commitment (from_values, rate_bits 1, cap_height 4).  ctl_ch: NULL (stand-alone prove: lookup challenges are drawn
 * from the challenger) or NUM_CHALLENGES (beta, gamma) pairs from get_grand_product_challenge_set of the multi-table
 * prover (then the lookup challenges are the betas).  The challenger is updated in place (state in / state out). */
int orc_prove_with_commitment(int t, int log_n, const uint64_t *trace_vals, const orc_batch *trace, const uint64_t *ctl_ch,
                              orc_challenger *ch, const uint64_t *pi, uint64_t *proof) {
  shape_t s = shape_of(t, log_n);
  orc_fri_params fp; orc_fri_params_standard_fast(log_n, &fp);
  size_t n = (size_t)1 << log_n;
  size_t cap_words = (size_t)4 << fp.cap_height;
  if (fri_total_arities(&fp) > log_n + fp.rate_bits - fp.cap_height) return 1; /* "FRI total reduction arity is too large." */
  if (s.n_ctl_zs && !ctl_ch) return 6; /* the table requires CTLs but no CTL challenges were given */
  uint64_t *w = proof;
  uint64_t *hdr = w; w += HEADER_WORDS;
  memset(hdr, 0, HEADER_WORDS * 8);
  hdr[0] = PROOF_MAGIC; hdr[1] = t; hdr[2] = log_n; hdr[3] = s.n_trace; hdr[4] = s.n_aux; hdr[5] = s.n_quot;
  hdr[6] = fp.cap_height; hdr[7] = s.n_layers; hdr[8] = 4; hdr[9] = s.final_len; hdr[10] = fp.num_query_rounds;
  hdr[11] = s.n_pi; hdr[12] = fp.rate_bits; hdr[13] = fp.proof_of_work_bits; hdr[14] = NUM_CHALLENGES;
  hdr[15] = orc_stark_proof_words(t, log_n); hdr[16] = s.n_ctl_zs; hdr[17] = s.n_lookup_cols; hdr[18] = s.n_ctl_helpers;
  memcpy(w, trace->cap, cap_words * 8); w += cap_words;

  /* lookup challenges: the CTL betas when there are CTL challenges, else get_grand_product_challenge_set's betas */
  uint64_t chs[MAX_CH_SCALARS] = {0};
  orc_batch *aux = NULL;
  if (orc_table_uses_lookup(t)) {
    if (ctl_ch) { for (int k = 0; k < NUM_CHALLENGES; k++) chs[k] = gl_canon(ctl_ch[2 * k]); }
    else {
      uint64_t raw[2 * NUM_CHALLENGES];
      orc_challenger_get_n(ch, 2 * NUM_CHALLENGES, raw);
      for (int k = 0; k < NUM_CHALLENGES; k++) chs[k] = raw[2 * k];
    }
  }
  if (ctl_ch) for (int k = 0; k < 2 * NUM_CHALLENGES; k++) chs[NUM_CHALLENGES + k] = gl_canon(ctl_ch[k]);
  uint64_t *aux_vals = NULL;
  if (s.n_aux) {
    aux_vals = (uint64_t *)malloc((size_t)s.n_aux * n * 8);
    if (orc_aux_columns(t, log_n, trace_vals, chs, NUM_CHALLENGES, chs + NUM_CHALLENGES, aux_vals)) { free(aux_vals); return 7; }
    aux = orc_batch_from_values(aux_vals, s.n_aux, log_n, fp.rate_bits, fp.cap_height);
    orc_challenger_observe(ch, aux->cap, cap_words);
    memcpy(w, aux->cap, cap_words * 8); w += cap_words;
  }
  uint64_t alphas[NUM_CHALLENGES];
  orc_challenger_get_n(ch, NUM_CHALLENGES, alphas);
  uint64_t *qchunks = (uint64_t *)malloc((size_t)s.n_quot * n * 8);
  orc_compute_quotient_polys(t, log_n, trace, aux, chs, pi, alphas, NUM_CHALLENGES, qchunks);
  orc_batch *quot = orc_batch_from_coeffs(qchunks, s.n_quot, log_n, fp.rate_bits, fp.cap_height);
  free(qchunks);
  orc_challenger_observe(ch, quot->cap, cap_words);
  memcpy(w, quot->cap, cap_words * 8); w += cap_words;

  uint64_t zeta_w[2];
  orc_challenger_get_n(ch, 2, zeta_w);
  gl2_t zeta = gl2(zeta_w[0], zeta_w[1]);
  uint64_t g = gl_root_of_unity(log_n);
  {
    gl2_t zp = zeta;
    for (int i = 0; i < log_n; i++) zp = gl2_mul(zp, zp);
    if (gl2_eq(zp, gl2(1, 0))) { free(aux_vals); orc_batch_free(aux); orc_batch_free(quot); return 2; } /* "Opening point is in the subgroup." */
  }
  gl2_t zeta_next = gl2_scalar_mul(zeta, g);

  /* StarkOpeningSet::new ; order: local, next, aux, aux_next, ctl_zs_first, quotient */
  const uint64_t z0[2] = {zeta.c0, zeta.c1}, z1[2] = {zeta_next.c0, zeta_next.c1};
  uint64_t *loc = w; w += 2 * s.n_trace;
  uint64_t *nxt = w; w += 2 * s.n_trace;
  uint64_t *au = w; w += 2 * s.n_aux;
  uint64_t *aun = w; w += 2 * s.n_aux;
  uint64_t *zs_first = w; w += s.n_ctl_zs;
  uint64_t *qu = w; w += 2 * s.n_quot;
  orc_batch_eval_at_ext_point(trace, z0, loc);
  orc_batch_eval_at_ext_point(trace, z1, nxt);
  if (aux) { orc_batch_eval_at_ext_point(aux, z0, au); orc_batch_eval_at_ext_point(aux, z1, aun); }
  orc_batch_eval_at_ext_point(quot, z0, qu);
  /* ctl_zs_first: the Z polynomials evaluated at 1 (= their first trace-domain value) */
  for (int k = 0; k < s.n_ctl_zs; k++) {
    const uint64_t one[2] = {1, 0};
    uint64_t *all = (uint64_t *)malloc(2 * (size_t)s.n_aux * 8);
    orc_batch_eval_at_ext_point(aux, one, all);
    zs_first[k] = all[2 * (s.n_lookup_cols + s.n_ctl_helpers + k)];
    if (all[2 * (s.n_lookup_cols + s.n_ctl_helpers + k) + 1] != 0 || zs_first[k] != gl_canon(aux_vals[(size_t)(s.n_lookup_cols + s.n_ctl_helpers + k) * n]))
      fprintf(stderr, "oracle: Z(1) differs from the first trace value\n");
    free(all);
  }
  free(aux_vals);
  /* observe_openings(to_fri_openings): zeta batch = local ++ aux ++ quotient ; next batch = next ++ aux_next ; ctl_zs_first (as ext) */
  orc_challenger_observe(ch, loc, 2 * s.n_trace);
  orc_challenger_observe(ch, au, 2 * s.n_aux);
  orc_challenger_observe(ch, qu, 2 * s.n_quot);
  orc_challenger_observe(ch, nxt, 2 * s.n_trace);
  orc_challenger_observe(ch, aun, 2 * s.n_aux);
  for (int k = 0; k < s.n_ctl_zs; k++) { uint64_t e[2] = {zs_first[k], 0}; orc_challenger_observe(ch, e, 2); }

  /* stark.fri_instance(zeta, g, num_ctl_helpers, num_ctl_zs, config) */
  const orc_batch *oracles[3]; size_t n_oracles = 0;
  const unsigned o_trace = (unsigned)n_oracles; oracles[n_oracles++] = trace;
  const unsigned o_aux = (unsigned)n_oracles; if (aux) oracles[n_oracles++] = aux;
  const unsigned o_quot = (unsigned)n_oracles; oracles[n_oracles++] = quot;
  orc_fri_poly *p0 = (orc_fri_poly *)malloc((size_t)(s.n_trace + s.n_aux + s.n_quot + 1) * sizeof *p0);
  orc_fri_poly *p1 = (orc_fri_poly *)malloc((size_t)(s.n_trace + s.n_aux + 1) * sizeof *p1);
  orc_fri_poly *p2 = (orc_fri_poly *)malloc((size_t)(s.n_ctl_zs + 1) * sizeof *p2);
  size_t k0 = 0, k1 = 0, k2 = 0;
  for (int c = 0; c < s.n_trace; c++) { p0[k0].oracle_index = o_trace; p0[k0++].polynomial_index = c; p1[k1].oracle_index = o_trace; p1[k1++].polynomial_index = c; }
  for (int c = 0; c < s.n_aux; c++) { p0[k0].oracle_index = o_aux; p0[k0++].polynomial_index = c; p1[k1].oracle_index = o_aux; p1[k1++].polynomial_index = c; }
  for (int c = 0; c < s.n_quot; c++) { p0[k0].oracle_index = o_quot; p0[k0++].polynomial_index = c; }
  for (int c = 0; c < s.n_ctl_zs; c++) { p2[k2].oracle_index = o_aux; p2[k2++].polynomial_index = s.n_lookup_cols + s.n_ctl_helpers + c; }
  orc_fri_batch batches[3];
  batches[0].point[0] = zeta.c0; batches[0].point[1] = zeta.c1; batches[0].polynomials = p0; batches[0].n_polynomials = k0;
  batches[1].point[0] = zeta_next.c0; batches[1].point[1] = zeta_next.c1; batches[1].polynomials = p1; batches[1].n_polynomials = k1;
  batches[2].point[0] = 1; batches[2].point[1] = 0; batches[2].polynomials = p2; batches[2].n_polynomials = k2;
  int rc = orc_prove_openings(batches, s.n_ctl_zs ? 3 : 2, oracles, n_oracles, ch, &fp, w);
  size_t oc[3]; for (size_t o = 0; o < n_oracles; o++) oc[o] = oracles[o]->n_cols;
  w += orc_fri_proof_words(oc, n_oracles, &fp);
  free(p0); free(p1); free(p2);
  for (int i = 0; i < s.n_pi; i++) *w++ = gl_canon(pi[i]);
  if (rc == 0 && (size_t)(w - proof) != hdr[15]) rc = 4;
  orc_batch_free(aux); orc_batch_free(quot);
  return rc;
}

/* starky::prover::prove: trace commitment; the challenger observes the public inputs, then the trace cap */
int orc_stark_prove(int t, int log_n, const uint64_t *trace_vals, const uint64_t *pi, uint64_t *proof) {
  orc_fri_params fp; orc_fri_params_standard_fast(log_n, &fp);
  orc_batch *trace = orc_batch_from_values(trace_vals, orc_table_num_columns(t), log_n, fp.rate_bits, fp.cap_height);
  orc_challenger ch; orc_challenger_init(&ch);
  orc_challenger_observe(&ch, pi, orc_table_num_public_inputs(t));
  orc_challenger_observe(&ch, trace->cap, (size_t)4 << fp.cap_height);
  int rc = orc_prove_with_commitment(t, log_n, trace_vals, trace, NULL, &ch, pi, proof);
  orc_batch_free(trace);
  return rc;
}
/* Challenger::compact (plonky2/src/iop/challenger.rs): flush pending inputs, drop buffered outputs, return the sponge state */
void orc_challenger_compact(orc_challenger *c, uint64_t state_out[12]) {
  if (c->n_in != 0) { uint64_t dummy = orc_challenger_get(c); (void)dummy; }
  c->n_out = 0;
  memcpy(state_out, c->state, 12 * 8);
}
