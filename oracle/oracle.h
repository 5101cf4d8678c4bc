/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C11 + OpenMP) of the plonky2 / starky
 * STARK proving hot path that eth-tx-proof's worker reaches through
 * /root/reference/ops/src/lib.rs:52 (generate_txn_proof), :72, :95.
 *
 * The algorithm lives in third-party crates that are NOT on disk:
 *   plonky2 0.2.2, plonky2_field 0.2.2, starky 0.4.0   (/root/reference/Cargo.lock:3441,3466,4529)
 *   evm_arithmetization 0.1.3 @ zk_evm 7e80405          (/root/reference/Cargo.lock:1675)
 * and there is no Rust toolchain, so oracle/_ref cannot be built.  Each function below names the
 * upstream file it restates.
 *
 * PARITY STATUS: pinned for the Goldilocks constants, the 360 Poseidon round constants (SHA-256) and
 * the Poseidon permutation (three upstream known-answer vectors).  Everything composed above the
 * permutation (sponge mode, FFT conventions, Merkle layout, FRI schedule, transcript order, the
 * memory-table constraint sketch) is restated from the published upstream design and is
 * **parity unpinned** against real plonky2 output: the reference's own tests hold no vector for it
 * (/root/reference/common/src/parsing.rs:57-105 is the whole test suite).
 *
 * Nothing under eth_tx_proof_b200/ may include, link or call this code.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include "goldilocks.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- poseidon.c ---- */
void orc_poseidon_constants(uint64_t out[360]);
void orc_poseidon_permute(uint64_t state[12]);       /* fast form (lazy reduction, split MDS) */
void orc_poseidon_permute_naive(uint64_t state[12]); /* the definition, round by round */
void orc_hash_no_pad(const uint64_t *in, size_t n, uint64_t out[4]);
void orc_hash_or_noop(const uint64_t *in, size_t n, uint64_t out[4]);
void orc_two_to_one(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]);

typedef struct {
  uint64_t state[12];
  uint64_t in[8];
  int n_in;
  uint64_t out[8];
  int n_out;
} orc_challenger;
void orc_challenger_init(orc_challenger *c);
void orc_challenger_observe(orc_challenger *c, const uint64_t *e, size_t n);
uint64_t orc_challenger_get(orc_challenger *c);
void orc_challenger_get_n(orc_challenger *c, size_t n, uint64_t *out);

/* ---- ntt.c : single-column transforms, natural order in and out ---- */
void orc_fft(uint64_t *a, int log_n);
void orc_ifft(uint64_t *a, int log_n);
void orc_coset_fft(uint64_t *a, int log_n, uint64_t shift);
void orc_coset_ifft(uint64_t *a, int log_n, uint64_t shift);
/* PolynomialCoeffs::lde(rate_bits).coset_fft(7): out has n << rate_bits values, natural order */
void orc_lde(const uint64_t *coeffs, int log_n, int rate_bits, uint64_t *out);

/* ---- merkle.c ---- */
/* leaves: n_leaves x leaf_len row-major. digests: 2*(n_leaves - 2^cap_height) x 4 in plonky2's
 * recursive layout; cap: 2^cap_height x 4. */
void orc_merkle_new(const uint64_t *leaves, size_t n_leaves, size_t leaf_len, int cap_height,
                    uint64_t *digests, uint64_t *cap);
/* siblings: (log2(n_leaves) - cap_height) x 4 */
void orc_merkle_prove(const uint64_t *digests, size_t n_leaves, int cap_height, size_t leaf_index,
                      uint64_t *siblings);
int orc_merkle_verify(const uint64_t *leaf, size_t leaf_len, size_t leaf_index, const uint64_t *siblings,
                      int n_siblings, const uint64_t *cap, int cap_height);

/* ---- batch.c : PolynomialBatch ---- */
typedef struct {
  size_t n_cols;
  int log_n, rate_bits, cap_height;
  uint64_t *coeffs;  /* n_cols x n, column-major (one Vec per polynomial upstream)      */
  uint64_t *leaves;  /* (n << rate_bits) x n_cols row-major, row i = point bitrev(i)     */
  uint64_t *digests; /* plonky2 layout                                                   */
  uint64_t *cap;     /* 2^cap_height x 4                                                 */
} orc_batch;
orc_batch *orc_batch_from_values(const uint64_t *values, size_t n_cols, int log_n, int rate_bits, int cap_height);
orc_batch *orc_batch_from_coeffs(const uint64_t *coeffs, size_t n_cols, int log_n, int rate_bits, int cap_height);
void orc_batch_free(orc_batch *b);
size_t orc_batch_num_digests(const orc_batch *b);
const uint64_t *orc_batch_coeffs(const orc_batch *b);
const uint64_t *orc_batch_leaves(const orc_batch *b);
const uint64_t *orc_batch_digests(const orc_batch *b);
const uint64_t *orc_batch_cap(const orc_batch *b);

/* ---- stark.c : tables, auxiliary columns, quotient, FRI, prove_with_commitment ---- */
#define ORC_TABLE_FIBONACCI 0
#define ORC_TABLE_MEMORY 1
/* FriConfig + FriParams (plonky2/src/fri/mod.rs) */
typedef struct {
  int rate_bits, cap_height, proof_of_work_bits, num_query_rounds, degree_bits, n_reductions;
  int reduction_arity_bits[16];
} orc_fri_params;
void orc_fri_params_make(int degree_bits, int rate_bits, int cap_height, int pow_bits, int num_queries, orc_fri_params *p);
void orc_fri_params_standard_fast(int degree_bits, orc_fri_params *p);
/* FriPolynomialInfo / FriBatchInfo (plonky2/src/fri/structure.rs) */
typedef struct { uint32_t oracle_index, polynomial_index; } orc_fri_poly;
typedef struct { uint64_t point[2]; const orc_fri_poly *polynomials; size_t n_polynomials; } orc_fri_batch;
/* program-defined table (constraint program + lookups, formats of the product's csrc/cprog.h / etp_table_register,
 * restated): interpreted op by op here.  Returns the table id (>= 16) or -1. */
int orc_table_register(const uint64_t *program, size_t n_words, const int32_t *lookups, size_t n_lookup_words);
/* general form (etp_table_register_ex): lookups with linear-combination Columns and Filters + the table's CTL Z descriptors */
int orc_table_register_ex(const uint64_t *program, size_t n_words, const uint64_t *aux_spec, size_t n_spec_words);
int orc_table_num_columns(int table);
int orc_table_constraint_degree(int table);
int orc_table_num_public_inputs(int table);
int orc_table_uses_lookup(int table);
int orc_table_requires_ctls(int table);
int orc_table_num_lookup_columns(int table, int n_challenges);
int orc_table_num_ctl_helper_columns(int table);
int orc_table_num_ctl_zs(int table);
int orc_table_num_aux_columns(int table, int n_challenges);
/* check_constraints analogue: -1 if the trace (column-major n_cols x n) satisfies the table's own
 * constraints on every row, else row*1000 + index of the first failing constraint */
long orc_table_check_constraints(int table, int log_n, const uint64_t *trace, const uint64_t *public_inputs);
/* number of u64 words of the flat proof for this table / size under standard_fast_config */
size_t orc_stark_proof_words(int table, int log_n);
/* starky::prover::prove, StarkConfig::standard_fast_config(); trace is column-major. Returns 0 on success. */
int orc_stark_prove(int table, int log_n, const uint64_t *trace, const uint64_t *public_inputs,
                    uint64_t *proof_out);
/* starky::prover::prove_with_commitment: pre-committed trace, optional CTL challenges, challenger state in/out */
int orc_prove_with_commitment(int table, int log_n, const uint64_t *trace_vals, const orc_batch *trace, const uint64_t *ctl_challenges,
                              orc_challenger *challenger, const uint64_t *public_inputs, uint64_t *proof_out);
void orc_challenger_compact(orc_challenger *c, uint64_t state_out[12]);
/* pieces exposed for piecewise parity tests */
void orc_lookup_helper_columns(int table, int log_n, const uint64_t *trace, const uint64_t *challenges,
                               int n_challenges, uint64_t *aux /* col-major */);
int orc_aux_columns(int table, int log_n, const uint64_t *trace, const uint64_t *lookup_challenges, int n_challenges,
                    const uint64_t *ctl_challenges, uint64_t *aux /* col-major */);
void orc_compute_quotient_polys(int table, int log_n, const orc_batch *trace, const orc_batch *aux,
                                const uint64_t *challenge_scalars, const uint64_t *public_inputs,
                                const uint64_t *alphas, int n_alphas, uint64_t *quotient_chunks);
uint64_t orc_pow_grind(const uint64_t state[12], int pos, int bits);
/* FRI commit phase on ext values (interleaved c0,c1; natural order) of size 2^log_size */
void orc_fri_fold_coeffs(const uint64_t *coeffs, size_t n, int arity_bits, const uint64_t beta[2], uint64_t *out);
void orc_batch_eval_at_ext_point(const orc_batch *b, const uint64_t z[2], uint64_t *out);
size_t orc_fri_proof_words(const size_t *oracle_cols, size_t n_oracles, const orc_fri_params *p);
int orc_prove_openings(const orc_fri_batch *batches, size_t n_batches, const orc_batch *const *oracles, size_t n_oracles,
                       orc_challenger *challenger, const orc_fri_params *params, uint64_t *fri_proof_out);

/* ---- plonk.c : first slice of plonky2's circuit prover ---- */
int orc_plonk_partial_products_and_zs(const uint64_t *wires, const uint64_t *sigmas, const uint64_t *k_is, int num_routed,
                                      int degree_bits, int quotient_degree_factor, const uint64_t *betas, const uint64_t *gammas,
                                      int num_challenges, uint64_t *out);

int orc_num_threads(void);
#ifdef __cplusplus
}
#endif
#endif
