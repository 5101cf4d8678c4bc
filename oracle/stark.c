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
#define NUM_CHALLENGES 2
#define RATE_BITS 1
#define CAP_HEIGHT 4
#define POW_BITS 16
#define ARITY_BITS 4
#define FINAL_POLY_BITS 5
#define NUM_QUERIES 84
#define PROOF_MAGIC 0x4232303053544B31ULL /* "B200STK1" */
#define MEM_TRIE_DATA_SEGMENT 13ULL

static int fri_num_layers(int degree_bits) {
  /* ConstantArityBits(4,5).reduction_arity_bits(degree_bits, rate_bits, cap_height, _) */
  int layers = 0;
  while (degree_bits > FINAL_POLY_BITS && degree_bits + RATE_BITS - ARITY_BITS >= CAP_HEIGHT) {
    layers++;
    degree_bits -= ARITY_BITS;
  }
  return layers;
}

/* ---------------- tables ---------------- */
/* Program-defined tables (constraint programs, format of eth_tx_proof_b200/csrc/cprog.h restated here):
 * the oracle INTERPRETS the program op by op; the product compiles it with NVRTC.  ids >= 16. */
#define ORC_MAX_TABLES 64
#define ORC_MAX_LOOKUPS 64
#define CPROG_MAGIC 0x3147525043505445ULL
enum { OP_CONST = 0, OP_LV, OP_NV, OP_LA, OP_NA, OP_PI, OP_CH, OP_ADD, OP_SUB, OP_MUL, OP_EMIT, OP_EMIT_TRANSITION, OP_EMIT_FIRST,
       OP_EMIT_LAST };
typedef struct {
  int n_looking, *looking, table_col, freq_col;
} lookup_t;
typedef struct {
  int used, cols, degree, n_pi, n_lookups;
  lookup_t lookups[ORC_MAX_LOOKUPS];
  uint32_t n_ops, n_aux, n_ch;
  uint64_t *ops; /* 2 words per op; NULL for the built-in tables */
} table_t;
static table_t g_tables[ORC_MAX_TABLES];
static int g_mem_looking[1] = {18};
static const table_t *get_table(int t) {
  if (!g_tables[ORC_TABLE_MEMORY].used) {
    table_t *f = &g_tables[ORC_TABLE_FIBONACCI], *m = &g_tables[ORC_TABLE_MEMORY];
    f->used = 1; f->cols = 2; f->degree = 2; f->n_pi = 3;
    m->used = 1; m->cols = 21; m->degree = 3; m->n_pi = 0; m->n_lookups = 1;
    m->lookups[0].n_looking = 1; m->lookups[0].looking = g_mem_looking; m->lookups[0].table_col = 19; m->lookups[0].freq_col = 20;
  }
  if (t < 0 || t >= ORC_MAX_TABLES || !g_tables[t].used) { fprintf(stderr, "oracle: unknown table %d\n", t); abort(); }
  return &g_tables[t];
}
/* lookups: [n_lookups, then per lookup: table_col, freq_col, n_looking, looking...]; returns the id or -1 */
int orc_table_register(const uint64_t *program, size_t n_words, const int32_t *lookups, size_t n_lookup_words) {
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
  if (n_lookup_words) {
    size_t pos = 0;
    t->n_lookups = lookups[pos++];
    for (int i = 0; i < t->n_lookups; i++) {
      lookup_t *l = &t->lookups[i];
      l->table_col = lookups[pos++]; l->freq_col = lookups[pos++]; l->n_looking = lookups[pos++];
      l->looking = (int *)malloc(sizeof(int) * l->n_looking);
      for (int j = 0; j < l->n_looking; j++) l->looking[j] = lookups[pos++];
    }
  }
  t->used = 1;
  return id;
}
int orc_table_num_columns(int t) { return get_table(t)->cols; }
int orc_table_constraint_degree(int t) { return get_table(t)->degree; }
int orc_table_num_public_inputs(int t) { return get_table(t)->n_pi; }
int orc_table_uses_lookup(int t) { return get_table(t)->n_lookups > 0; }
static int lookup_chunk(const table_t *t) { return t->degree - 1 < 1 ? 1 : t->degree - 1; }
static int lookup_helpers(const table_t *t, const lookup_t *l) { return (l->n_looking + lookup_chunk(t) - 1) / lookup_chunk(t); }
static int quotient_degree_factor(int t) {
  int d = orc_table_constraint_degree(t) - 1;
  return d < 1 ? 1 : d;
}
static int log2_ceil(int x) { int l = 0; while ((1 << l) < x) l++; return l; }
/* per lookup and challenge: num_helper_columns = ceil(n_looking / (degree-1)) helpers + Z */
int orc_table_num_aux_columns(int t, int n_challenges) {
  const table_t *tb = get_table(t);
  int a = 0;
  for (int i = 0; i < tb->n_lookups; i++) a += lookup_helpers(tb, &tb->lookups[i]) + 1;
  return a * n_challenges;
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

/* constraint program interpreter: own constraints AND lookup checks are in the program */
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

/* eval_packed_lookups_generic (no filters): per lookup, per challenge: helpers..., Z */
static void eval_lookups(int t, const uint64_t *lv, const uint64_t *aux_l, const uint64_t *aux_n,
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
        /* eval_helper_columns: h * prod(col_j + ch) - sum_j prod_{i != j}(col_i + ch) */
        int j0 = hc * chunk, j1 = j0 + chunk < l->n_looking ? j0 + chunk : l->n_looking;
        uint64_t h = aux_l[start + hc], prod = 1, rhs = 0;
        for (int j = j0; j < j1; j++) prod = gl_mul(prod, gl_add(lv[l->looking[j]], ch));
        if (j1 - j0 == 1) rhs = 1;
        else
          for (int j = j0; j < j1; j++) {
            uint64_t tp = 1;
            for (int i = j0; i < j1; i++) if (i != j) tp = gl_mul(tp, gl_add(lv[l->looking[i]], ch));
            rhs = gl_add(rhs, tp);
          }
        c_constraint(c, gl_sub(gl_mul(h, prod), rhs));
        hsum = gl_add(hsum, h);
      }
      uint64_t z = aux_l[start + nh], next_z = aux_n[start + nh];
      uint64_t table_with_challenge = gl_add(lv[l->table_col], ch);
      uint64_t y = gl_sub(gl_mul(hsum, table_with_challenge), lv[l->freq_col]);
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

/* ---------------- lookup helper columns (starky lookup.rs: lookup_helper_columns) ---------------- */
static void batch_inverse(uint64_t *x, size_t n) {
  uint64_t *pre = (uint64_t *)malloc(n * sizeof(uint64_t));
  uint64_t acc = 1;
  for (size_t i = 0; i < n; i++) { pre[i] = acc; acc = gl_mul(acc, x[i]); }
  uint64_t inv = gl_inv(acc);
  for (size_t i = n; i-- > 0;) { uint64_t xi = x[i]; x[i] = gl_mul(inv, pre[i]); inv = gl_mul(inv, xi); }
  free(pre);
}
void orc_lookup_helper_columns(int t, int log_n, const uint64_t *trace, const uint64_t *challenges,
                               int n_ch, uint64_t *aux) {
  const table_t *tb = get_table(t);
  if (!tb->n_lookups) return;
  size_t n = (size_t)1 << log_n;
  const int chunk = lookup_chunk(tb);
  uint64_t *inv = (uint64_t *)malloc(n * sizeof(uint64_t)), *tinv = (uint64_t *)malloc(n * sizeof(uint64_t));
  uint64_t *out = aux;
  for (int li = 0; li < tb->n_lookups; li++) {
    const lookup_t *l = &tb->lookups[li];
    const int nh = lookup_helpers(tb, l);
    const uint64_t *table = trace + (size_t)l->table_col * n, *freq = trace + (size_t)l->freq_col * n;
    for (int k = 0; k < n_ch; k++) {
      uint64_t ch = challenges[k];
      uint64_t *z = out + (size_t)nh * n;
      for (int hc = 0; hc < nh; hc++) {
        uint64_t *h = out + (size_t)hc * n;
        memset(h, 0, n * 8);
        for (int j = hc * chunk; j < (hc + 1) * chunk && j < l->n_looking; j++) {
          const uint64_t *col = trace + (size_t)l->looking[j] * n;
          for (size_t i = 0; i < n; i++) inv[i] = gl_add(col[i], ch);
          batch_inverse(inv, n);
          for (size_t i = 0; i < n; i++) h[i] = gl_add(h[i], inv[i]);
        }
        for (size_t i = 0; i < n; i++) h[i] = gl_canon(h[i]);
      }
      for (size_t i = 0; i < n; i++) tinv[i] = gl_add(table[i], ch);
      batch_inverse(tinv, n);
      z[0] = 0;
      for (size_t i = 0; i + 1 < n; i++) {
        uint64_t tot = 0;
        for (int hc = 0; hc < nh; hc++) tot = gl_add(tot, out[(size_t)hc * n + i]);
        z[i + 1] = gl_canon(gl_add(z[i], gl_sub(tot, gl_mul(freq[i], tinv[i]))));
      }
      out += (size_t)(nh + 1) * n;
    }
  }
  free(inv); free(tinv);
}

/* ---------------- compute_quotient_polys ---------------- */
void orc_compute_quotient_polys(int t, int log_n, const orc_batch *trace, const orc_batch *aux,
                                const uint64_t *lookup_challenges, const uint64_t *pi,
                                const uint64_t *alphas, int n_alphas, uint64_t *quotient_chunks) {
  size_t degree = (size_t)1 << log_n;
  int qbits = log2_ceil(quotient_degree_factor(t));
  int step = 1 << (RATE_BITS - qbits), next_step = 1 << qbits;
  size_t size = degree << qbits;
  int log_size = log_n + qbits, log_lde = log_n + RATE_BITS;
  int nc = orc_table_num_columns(t), na = aux ? (int)aux->n_cols : 0;
  /* Lagrange selectors on the coset: selector(degree, i).lde_onto_coset(qbits) */
  uint64_t *lfirst = (uint64_t *)calloc(size, 8), *llast = (uint64_t *)calloc(size, 8);
  uint64_t *tmp = (uint64_t *)calloc(degree, 8);
  tmp[0] = 1; orc_ifft(tmp, log_n); memcpy(lfirst, tmp, degree * 8); orc_coset_fft(lfirst, log_size, GL_GENERATOR);
  memset(tmp, 0, degree * 8);
  tmp[degree - 1] = 1; orc_ifft(tmp, log_n); memcpy(llast, tmp, degree * 8); orc_coset_fft(llast, log_size, GL_GENERATOR);
  free(tmp);
  /* ZeroPolyOnCoset::new(degree_bits, qbits): Z_H(x_i) = g^n * w_{2^qbits}^(i mod 2^qbits) - 1 */
  uint64_t zh_inv[2];
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
      eval_program(get_table(t), lv, nv, al, an, pi, lookup_challenges, &cs);
    } else {
      eval_table(t, lv, nv, pi, &cs);
      if (aux) eval_lookups(t, lv, al, an, lookup_challenges, NUM_CHALLENGES, &cs);
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

static void observe_cap(orc_challenger *ch, const uint64_t *cap) { orc_challenger_observe(ch, cap, (size_t)4 << CAP_HEIGHT); }

/* ---------------- proof sizes ---------------- */
typedef struct {
  int table, log_n, n_trace, n_aux, n_quot, n_layers, final_len, n_pi;
} shape_t;
static shape_t shape_of(int t, int log_n) {
  shape_t s;
  s.table = t; s.log_n = log_n; s.n_trace = orc_table_num_columns(t);
  s.n_aux = orc_table_num_aux_columns(t, NUM_CHALLENGES);
  s.n_quot = quotient_degree_factor(t) * NUM_CHALLENGES;
  s.n_layers = fri_num_layers(log_n);
  s.final_len = 1 << (log_n - ARITY_BITS * s.n_layers);
  s.n_pi = orc_table_num_public_inputs(t);
  return s;
}
size_t orc_stark_proof_words(int t, int log_n) {
  shape_t s = shape_of(t, log_n);
  size_t cap = (size_t)4 << CAP_HEIGHT, w = 16;
  int log_lde = log_n + RATE_BITS;
  w += cap * (2 + (s.n_aux ? 1 : 0));
  w += 2 * (size_t)(2 * s.n_trace + 2 * s.n_aux + s.n_quot);
  w += cap * s.n_layers;
  size_t per_query = 0;
  int init_path = log_lde - CAP_HEIGHT;
  per_query += s.n_trace + 4 * init_path;
  if (s.n_aux) per_query += s.n_aux + 4 * init_path;
  per_query += s.n_quot + 4 * init_path;
  int bits = log_lde;
  for (int l = 0; l < s.n_layers; l++) {
    bits -= ARITY_BITS;
    per_query += 2 * (1 << ARITY_BITS) + 4 * (bits - CAP_HEIGHT);
  }
  w += NUM_QUERIES * per_query;
  w += 2 * (size_t)s.final_len + 1 + s.n_pi;
  return w;
}

/* evaluate a base-coefficient polynomial at an ext point (to_extension().eval(z): Horner) */
static gl2_t eval_base_poly(const uint64_t *c, size_t n, gl2_t z) {
  gl2_t acc = gl2(0, 0);
  for (size_t i = n; i-- > 0;) acc = gl2_add(gl2_mul(acc, z), gl2_from_base(c[i]));
  return acc;
}

int orc_stark_prove(int t, int log_n, const uint64_t *trace_vals, const uint64_t *pi, uint64_t *proof) {
  shape_t s = shape_of(t, log_n);
  size_t n = (size_t)1 << log_n;
  int log_lde = log_n + RATE_BITS;
  size_t lde_n = (size_t)1 << log_lde;
  size_t cap_words = (size_t)4 << CAP_HEIGHT;
  if (ARITY_BITS * s.n_layers > log_n + RATE_BITS - CAP_HEIGHT) return 1; /* "FRI total reduction arity is too large." */
  uint64_t *w = proof;
  uint64_t *hdr = w; w += 16;
  hdr[0] = PROOF_MAGIC; hdr[1] = t; hdr[2] = log_n; hdr[3] = s.n_trace; hdr[4] = s.n_aux; hdr[5] = s.n_quot;
  hdr[6] = CAP_HEIGHT; hdr[7] = s.n_layers; hdr[8] = ARITY_BITS; hdr[9] = s.final_len; hdr[10] = NUM_QUERIES;
  hdr[11] = s.n_pi; hdr[12] = RATE_BITS; hdr[13] = POW_BITS; hdr[14] = NUM_CHALLENGES;
  hdr[15] = orc_stark_proof_words(t, log_n);

  /* prove(): trace commitment, challenger observes public inputs then the trace cap */
  orc_batch *trace = orc_batch_from_values(trace_vals, s.n_trace, log_n, RATE_BITS, CAP_HEIGHT);
  orc_challenger ch; orc_challenger_init(&ch);
  orc_challenger_observe(&ch, pi, s.n_pi);
  observe_cap(&ch, trace->cap);
  memcpy(w, trace->cap, cap_words * 8); w += cap_words;

  /* prove_with_commitment */
  uint64_t lookup_ch[NUM_CHALLENGES] = {0};
  orc_batch *aux = NULL;
  if (orc_table_uses_lookup(t)) {
    /* get_grand_product_challenge_set(challenger, num_challenges): (beta, gamma) per challenge;
     * the lookup argument uses beta */
    uint64_t raw[2 * NUM_CHALLENGES];
    orc_challenger_get_n(&ch, 2 * NUM_CHALLENGES, raw);
    for (int k = 0; k < NUM_CHALLENGES; k++) lookup_ch[k] = raw[2 * k];
    uint64_t *aux_vals = (uint64_t *)malloc((size_t)s.n_aux * n * 8);
    orc_lookup_helper_columns(t, log_n, trace_vals, lookup_ch, NUM_CHALLENGES, aux_vals);
    aux = orc_batch_from_values(aux_vals, s.n_aux, log_n, RATE_BITS, CAP_HEIGHT);
    free(aux_vals);
    observe_cap(&ch, aux->cap);
    memcpy(w, aux->cap, cap_words * 8); w += cap_words;
  }
  uint64_t alphas[NUM_CHALLENGES];
  orc_challenger_get_n(&ch, NUM_CHALLENGES, alphas);
  uint64_t *qchunks = (uint64_t *)malloc((size_t)s.n_quot * n * 8);
  orc_compute_quotient_polys(t, log_n, trace, aux, lookup_ch, pi, alphas, NUM_CHALLENGES, qchunks);
  orc_batch *quot = orc_batch_from_coeffs(qchunks, s.n_quot, log_n, RATE_BITS, CAP_HEIGHT);
  free(qchunks);
  observe_cap(&ch, quot->cap);
  memcpy(w, quot->cap, cap_words * 8); w += cap_words;

  uint64_t zeta_w[2];
  orc_challenger_get_n(&ch, 2, zeta_w);
  gl2_t zeta = gl2(zeta_w[0], zeta_w[1]);
  uint64_t g = gl_root_of_unity(log_n);
  {
    gl2_t zp = zeta;
    for (int i = 0; i < log_n; i++) zp = gl2_mul(zp, zp);
    if (gl2_eq(zp, gl2(1, 0))) return 2; /* "Opening point is in the subgroup." */
  }
  gl2_t zeta_next = gl2_scalar_mul(zeta, g);

  /* StarkOpeningSet::new ; order: local, next, aux, aux_next, quotient */
  int n_all = s.n_trace + s.n_aux + s.n_quot;
  const uint64_t **polys = (const uint64_t **)malloc(n_all * sizeof(*polys));
  for (int c = 0; c < s.n_trace; c++) polys[c] = trace->coeffs + (size_t)c * n;
  for (int c = 0; c < s.n_aux; c++) polys[s.n_trace + c] = aux->coeffs + (size_t)c * n;
  for (int c = 0; c < s.n_quot; c++) polys[s.n_trace + s.n_aux + c] = quot->coeffs + (size_t)c * n;
  gl2_t *ev_zeta = (gl2_t *)malloc(n_all * sizeof(gl2_t)), *ev_next = (gl2_t *)malloc(n_all * sizeof(gl2_t));
#pragma omp parallel for schedule(dynamic)
  for (int c = 0; c < n_all; c++) {
    ev_zeta[c] = eval_base_poly(polys[c], n, zeta);
    if (c < s.n_trace + s.n_aux) ev_next[c] = eval_base_poly(polys[c], n, zeta_next);
  }
  uint64_t *op = w;
  for (int c = 0; c < s.n_trace; c++) { *w++ = ev_zeta[c].c0; *w++ = ev_zeta[c].c1; }
  for (int c = 0; c < s.n_trace; c++) { *w++ = ev_next[c].c0; *w++ = ev_next[c].c1; }
  for (int c = 0; c < s.n_aux; c++) { *w++ = ev_zeta[s.n_trace + c].c0; *w++ = ev_zeta[s.n_trace + c].c1; }
  for (int c = 0; c < s.n_aux; c++) { *w++ = ev_next[s.n_trace + c].c0; *w++ = ev_next[s.n_trace + c].c1; }
  for (int c = 0; c < s.n_quot; c++) { *w++ = ev_zeta[s.n_trace + s.n_aux + c].c0; *w++ = ev_zeta[s.n_trace + s.n_aux + c].c1; }
  /* observe_openings(to_fri_openings): zeta batch = local ++ aux ++ quotient ; next batch = next ++ aux_next */
  {
    const uint64_t *loc = op, *nxt = op + 2 * s.n_trace, *au = nxt + 2 * s.n_trace, *aun = au + 2 * s.n_aux, *qu = aun + 2 * s.n_aux;
    orc_challenger_observe(&ch, loc, 2 * s.n_trace);
    orc_challenger_observe(&ch, au, 2 * s.n_aux);
    orc_challenger_observe(&ch, qu, 2 * s.n_quot);
    orc_challenger_observe(&ch, nxt, 2 * s.n_trace);
    orc_challenger_observe(&ch, aun, 2 * s.n_aux);
  }

  /* PolynomialBatch::prove_openings */
  uint64_t alpha_w[2];
  orc_challenger_get_n(&ch, 2, alpha_w);
  gl2_t alpha = gl2(alpha_w[0], alpha_w[1]);
  gl2_t *final_poly = (gl2_t *)calloc(n, sizeof(gl2_t));
  gl2_t *comp = (gl2_t *)malloc(n * sizeof(gl2_t));
  for (int batch = 0; batch < 2; batch++) {
    int count = batch == 0 ? n_all : s.n_trace + s.n_aux;
    gl2_t point = batch == 0 ? zeta : zeta_next;
    /* reduce_polys_base: sum_k alpha^k * poly_k */
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
      gl2_t acc = gl2(0, 0);
      for (int k = count; k-- > 0;) acc = gl2_add(gl2_mul(acc, alpha), gl2_from_base(polys[k][i]));
      comp[i] = acc;
    }
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
  free(comp); free(ev_zeta); free(ev_next);

  /* lde(rate_bits) + coset_fft(7) over the extension */
  uint64_t *coeffs = (uint64_t *)calloc(2 * lde_n, 8), *values = (uint64_t *)malloc(2 * lde_n * 8);
  for (size_t i = 0; i < n; i++) { coeffs[2 * i] = final_poly[i].c0; coeffs[2 * i + 1] = final_poly[i].c1; }
  free(final_poly);
  memcpy(values, coeffs, 2 * lde_n * 8);
  ext_coset_fft(values, log_lde, GL_GENERATOR);

  /* fri_committed_trees */
  uint64_t *layer_leaves[8], *layer_digests[8], *layer_caps[8];
  size_t layer_nleaves[8];
  size_t cur_n = lde_n;
  int cur_log = log_lde;
  uint64_t shift = GL_GENERATOR;
  for (int l = 0; l < s.n_layers; l++) {
    size_t arity = (size_t)1 << ARITY_BITS, nl = cur_n / arity;
    uint64_t *leaves = (uint64_t *)malloc(2 * cur_n * 8);
    for (size_t i = 0; i < cur_n; i++) { /* reverse_index_bits_in_place then chunk + flatten */
      size_t src = bitrev64(i, cur_log);
      leaves[2 * i] = values[2 * src]; leaves[2 * i + 1] = values[2 * src + 1];
    }
    size_t nd = 2 * (nl - ((size_t)1 << CAP_HEIGHT));
    layer_leaves[l] = leaves; layer_nleaves[l] = nl;
    layer_digests[l] = (uint64_t *)malloc((nd ? nd : 1) * 32);
    layer_caps[l] = (uint64_t *)malloc(cap_words * 8);
    orc_merkle_new(leaves, nl, 2 * arity, CAP_HEIGHT, layer_digests[l], layer_caps[l]);
    observe_cap(&ch, layer_caps[l]);
    memcpy(w, layer_caps[l], cap_words * 8); w += cap_words;
    uint64_t beta[2];
    orc_challenger_get_n(&ch, 2, beta);
    uint64_t *folded = (uint64_t *)malloc(2 * nl * 8);
    orc_fri_fold_coeffs(coeffs, cur_n, ARITY_BITS, beta, folded);
    free(coeffs); coeffs = folded;
    shift = gl_pow(shift, arity);
    cur_n = nl; cur_log -= ARITY_BITS;
    memcpy(values, coeffs, 2 * cur_n * 8);
    ext_coset_fft(values, cur_log, shift);
  }
  /* final poly: drop the (zero) top rate_bits part, observe */
  size_t final_len = cur_n >> RATE_BITS;
  for (size_t i = final_len; i < cur_n; i++)
    if (coeffs[2 * i] || coeffs[2 * i + 1]) fprintf(stderr, "oracle: FRI final poly high part non-zero\n");
  orc_challenger_observe(&ch, coeffs, 2 * final_len);

  /* fri_proof_of_work (input_buffer is overwritten into the sponge, candidate goes at n_in) */
  uint64_t st[12];
  memcpy(st, ch.state, sizeof st);
  for (int i = 0; i < ch.n_in; i++) st[i] = ch.in[i];
  uint64_t pow_witness = orc_pow_grind(st, ch.n_in, POW_BITS);
  orc_challenger_observe(&ch, &pow_witness, 1);
  uint64_t pow_response = orc_challenger_get(&ch);
  if ((pow_response >> (64 - POW_BITS)) != 0) return 3;

  /* fri_prover_query_rounds */
  uint64_t qidx[NUM_QUERIES];
  orc_challenger_get_n(&ch, NUM_QUERIES, qidx);
  const orc_batch *init[3] = {trace, aux, quot};
  for (int q = 0; q < NUM_QUERIES; q++) {
    size_t x = (size_t)(qidx[q] % lde_n);
    for (int o = 0; o < 3; o++) {
      const orc_batch *b = init[o];
      if (!b) continue;
      memcpy(w, b->leaves + x * b->n_cols, b->n_cols * 8); w += b->n_cols;
      orc_merkle_prove(b->digests, lde_n, CAP_HEIGHT, x, w); w += 4 * (log_lde - CAP_HEIGHT);
    }
    int bits = log_lde;
    for (int l = 0; l < s.n_layers; l++) {
      size_t arity = (size_t)1 << ARITY_BITS;
      x >>= ARITY_BITS; bits -= ARITY_BITS;
      memcpy(w, layer_leaves[l] + x * 2 * arity, 2 * arity * 8); w += 2 * arity;
      orc_merkle_prove(layer_digests[l], layer_nleaves[l], CAP_HEIGHT, x, w); w += 4 * (bits - CAP_HEIGHT);
    }
  }
  memcpy(w, coeffs, 2 * final_len * 8); w += 2 * final_len;
  *w++ = pow_witness;
  for (int i = 0; i < s.n_pi; i++) *w++ = gl_canon(pi[i]);
  int rc = ((size_t)(w - proof) == hdr[15]) ? 0 : 4;

  for (int l = 0; l < s.n_layers; l++) { free(layer_leaves[l]); free(layer_digests[l]); free(layer_caps[l]); }
  free(coeffs); free(values); free(polys);
  orc_batch_free(trace); orc_batch_free(aux); orc_batch_free(quot);
  return rc;
}
