/*
 * etp_b200 — C ABI of the B200-native STARK proving hot path (libetp_b200.so).
 *
 * The reference (0xPolygonZero/eth-tx-proof) has no FFI, plugin or backend trait on this path: its
 * worker calls `generate_txn_proof` (/root/reference/ops/src/lib.rs:52; also :72, :95), which runs
 * plonky2 0.2.2 / starky 0.4.0 (/root/reference/Cargo.lock:3441,4529; crates not on disk) on the CPU
 * under `StarkConfig::standard_fast_config()` (/root/reference/common/src/prover_state/circuit.rs:204).
 * The drop-in seam is Cargo source replacement of those crates; the entry points
 * below are what a patched plonky2/starky would bind (rust/etp_b200_sys/src/lib.rs).
 * Each one names the upstream Rust item it replaces.
 *
 * Conventions
 *  - all field elements are Goldilocks u64, little endian; inputs may be non-canonical (>= p),
 *    every output is canonical (< p);
 *  - "columns" are column-major: column c is `n` contiguous u64 (upstream: Vec<PolynomialValues<F>> /
 *    Vec<PolynomialCoeffs<F>>, one Vec per polynomial);
 *  - every function returns ETP_OK (0) or a negative error code; etp_last_error() gives the message
 *    for the calling context.  Upstream panics on misuse (assert!); the Rust glue turns a non-zero
 *    status into a panic / FatalError.  No exception or unwind crosses this boundary;
 *  - a context owns one CUDA device + stream + scratch pools.  Contexts are independent; one context
 *    must not be used from two threads at once (upstream: one tokio worker task per op);
 *  - *_host entry points take HOST pointers and copy; *_dev entry points take DEVICE pointers on the
 *    context's device (used by the benches and by callers that keep traces resident);
 *  - there is no CPU fallback: without a usable CUDA device etp_ctx_create fails.
 */
#ifndef ETP_B200_H
#define ETP_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ETP_OK 0
#define ETP_ERR_INVALID -1   /* bad argument (upstream: assert! panic) */
#define ETP_ERR_CUDA -2      /* CUDA runtime / launch failure, OOM        */
#define ETP_ERR_STATE -3     /* object used in the wrong state            */
#define ETP_ERR_PROOF -4     /* prover-side check failed (e.g. zeta in the subgroup, quotient not a polynomial) */

typedef struct etp_ctx etp_ctx;
typedef struct etp_batch etp_batch;   /* plonky2::fri::oracle::PolynomialBatch<GoldilocksField, PoseidonGoldilocksConfig, 2> */
typedef struct etp_tree etp_tree;     /* plonky2::hash::merkle_tree::MerkleTree<GoldilocksField, PoseidonHash>              */

/* ---- library / context ---------------------------------------------------------------------- */
const char *etp_version(void);
int etp_device_count(void);
int etp_ctx_create(int device, etp_ctx **out);
void etp_ctx_destroy(etp_ctx *ctx);
const char *etp_last_error(const etp_ctx *ctx);
int etp_ctx_synchronize(etp_ctx *ctx);
/* the context's cudaStream_t, for callers that enqueue their own work or time with CUDA events */
void *etp_ctx_stream(etp_ctx *ctx);
/* number of kernel launches issued by this context since creation */
uint64_t etp_ctx_launch_count(const etp_ctx *ctx);
/* Host-side Poseidon-12 permutation (plonky2/src/hash/poseidon.rs `Poseidon::poseidon`), in place: any u64 in, canonical
 * out.  Needs no GPU: it is what the library's Fiat-Shamir Challenger runs on the calling thread (the transcript is
 * strictly sequential); exported so that a caller-side Challenger can share it. */
void etp_host_poseidon_permute(uint64_t state[12]);
/* ALL_ROUND_CONSTANTS (30 x 12), MDS_MATRIX_CIRC, MDS_MATRIX_DIAG of plonky2/src/hash/poseidon_goldilocks.rs — for callers
 * that build circuits over the permutation (PoseidonGate); any pointer may be NULL */
void etp_poseidon_constants(uint64_t round_constants_out[360], uint64_t mds_circ_out[12], uint64_t mds_diag_out[12]);
/* Witness of one PoseidonGate row (plonky2/src/gates/poseidon.rs PoseidonGenerator::run_once): the 135 wires — inputs,
 * outputs, swap flag, deltas and the S-box inputs of every round after the first — for the given 12 inputs.  Host only;
 * the witness generation of the recursion layers' circuits (Merkle paths, sponges, in-circuit challenger) is mostly this. */
void etp_host_poseidon_gate_wires(const uint64_t inputs[12], int swap, uint64_t wires_out[135]);
/* Device scratch of a context comes from a per-context block cache (a released block is reused by the next
 * allocation of about the same size, so proofs / commits of a shape seen before allocate nothing).
 * etp_ctx_trim returns the cached blocks to the CUDA runtime; etp_ctx_cached_bytes reports how much is held. */
int etp_ctx_trim(etp_ctx *ctx);
size_t etp_ctx_cached_bytes(const etp_ctx *ctx);
/* Page-lock / release caller-owned host memory (cudaHostRegister): the *_host entry points then copy from it at full PCIe
 * speed, overlapped with the transforms.  Optional: pageable memory is accepted everywhere, its copies are just slower. */
int etp_host_pin(etp_ctx *ctx, void *ptr, size_t bytes);
int etp_host_unpin(etp_ctx *ctx, void *ptr);
/* device memory helpers (cudaMalloc / cudaFree / cudaMemcpyAsync on the context's stream + sync) */
int etp_dev_alloc(etp_ctx *ctx, size_t bytes, void **out);
int etp_dev_free(etp_ctx *ctx, void *ptr);
int etp_dev_upload(etp_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);
int etp_dev_download(etp_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);

/* Sustained issue rates of the integer / FP64 pipes of this device (thread-instructions per second, whole GPU), measured
 * with dependent-chain micro-kernels on the context's stream: [0] IMAD.WIDE.U32 with a zero addend, [1] mad.wide.u32
 * accumulating into a 64-bit addend (IMAD.WIDE + a 64-bit add), [2] DFMA.  bench.py uses them as the integer-pipe roofline denominator
 * of the Poseidon kernels ("an IMAD.WIDE.U32 micro-benchmark peak measured in the same run"). */
int etp_bench_pipe_rates(etp_ctx *ctx, double rates_out[3]);

/* ---- Fiat-Shamir transcript: plonky2/src/iop/challenger.rs Challenger<GoldilocksField, PoseidonHash> --------------
 * Plain data, owned by the caller: the "challenger state in / out" of every proving entry point below.  The fields are
 * upstream's: sponge_state, input_buffer (pending observations, < 8), output_buffer (challenges popped from the END).
 * Host-only functions (no GPU): duplex sponge in overwrite mode, rate 8. */
typedef struct etp_challenger {
  uint64_t sponge_state[12];
  uint64_t input_buffer[8];
  uint64_t output_buffer[8];
  uint32_t input_len, output_len;
} etp_challenger;
void etp_challenger_init(etp_challenger *c);                                        /* Challenger::new            */
void etp_challenger_observe(etp_challenger *c, const uint64_t *elements, size_t n); /* observe_elements / _cap    */
uint64_t etp_challenger_get_challenge(etp_challenger *c);                           /* get_challenge              */
void etp_challenger_get_n_challenges(etp_challenger *c, size_t n, uint64_t *out);   /* get_n_challenges           */
void etp_challenger_compact(etp_challenger *c);                                     /* compact (state = sponge_state) */

/* ---- FRI parameters: plonky2/src/fri/mod.rs FriConfig + FriParams (hiding = false) -------------------------------- */
typedef struct etp_fri_params {
  int rate_bits, cap_height, proof_of_work_bits, num_query_rounds, degree_bits, n_reductions;
  int reduction_arity_bits[16]; /* every entry must be 4 (FriReductionStrategy::ConstantArityBits(4, 5), the strategy of both
                                   StarkConfig::standard_fast_config and CircuitConfig::standard_recursion_config) */
} etp_fri_params;
/* FriConfig::fri_params(degree_bits, false) with ConstantArityBits(4, 5) */
int etp_fri_params_make(int degree_bits, int rate_bits, int cap_height, int proof_of_work_bits, int num_query_rounds,
                        etp_fri_params *out);

/* ---- hashing primitives: plonky2/src/hash/poseidon.rs, hashing.rs ---------------------------- */
/* PoseidonPermutation::permute on n independent 12-lane states (host pointers, n x 12) */
int etp_poseidon_permute_host(etp_ctx *ctx, uint64_t *states, size_t n);

/* ---- transforms: plonky2_field field/src/fft.rs, polynomial/mod.rs --------------------------- */
/* PolynomialValues::ifft on n_cols columns of 2^log_n values (host, in place, natural order) */
int etp_ifft_host(etp_ctx *ctx, uint64_t *cols, size_t n_cols, int log_n);
/* PolynomialCoeffs::fft */
int etp_fft_host(etp_ctx *ctx, uint64_t *cols, size_t n_cols, int log_n);
/* PolynomialCoeffs::lde(rate_bits).coset_fft(shift): out has n_cols x 2^(log_n+rate_bits), NATURAL order */
int etp_coset_lde_host(etp_ctx *ctx, const uint64_t *coeffs, size_t n_cols, int log_n, int rate_bits,
                       uint64_t shift, uint64_t *out);
/* PolynomialValues::coset_ifft(shift), in place */
int etp_coset_ifft_host(etp_ctx *ctx, uint64_t *cols, size_t n_cols, int log_n, uint64_t shift);

/* ---- MerkleTree::new(leaves, cap_height): plonky2/src/hash/merkle_tree.rs ---------------------- */
/* leaves: n_leaves x leaf_len row-major (upstream Vec<Vec<F>>), n_leaves a power of two,
 * cap_height <= log2(n_leaves) (else ETP_ERR_INVALID, upstream asserts). */
int etp_merkle_new_host(etp_ctx *ctx, const uint64_t *leaves, size_t n_leaves, size_t leaf_len, int cap_height,
                        etp_tree **out);
void etp_tree_free(etp_tree *t);
size_t etp_tree_num_digests(const etp_tree *t);           /* 2 * (n_leaves - 2^cap_height)                 */
int etp_tree_cap(etp_tree *t, uint64_t *cap_out);          /* MerkleTree.cap: 2^cap_height x 4              */
int etp_tree_digests(etp_tree *t, uint64_t *digests_out);  /* MerkleTree.digests in plonky2's layout        */
/* MerkleTree::prove(leaf_index): siblings_out has (log2(n_leaves) - cap_height) x 4 */
int etp_tree_prove(etp_tree *t, size_t leaf_index, uint64_t *siblings_out);

/* ---- PolynomialBatch: plonky2/src/fri/oracle.rs ------------------------------------------------ */
/* from_values(values, rate_bits, blinding, cap_height, timing, fft_root_table).
 * blinding must be 0 (the STARK path never blinds; ETP_ERR_INVALID otherwise).
 * cols: n_cols pointers, each to 2^log_n u64 (host). */
int etp_batch_from_values_host(etp_ctx *ctx, const uint64_t *const *cols, size_t n_cols, int log_n, int rate_bits,
                               int blinding, int cap_height, etp_batch **out);
/* from_coeffs(polynomials, ...) */
int etp_batch_from_coeffs_host(etp_ctx *ctx, const uint64_t *const *cols, size_t n_cols, int log_n, int rate_bits,
                               int blinding, int cap_height, etp_batch **out);
/* same with one device matrix (column c at values_dev + c * col_stride); the input is not modified */
int etp_batch_from_values_dev(etp_ctx *ctx, const uint64_t *values_dev, size_t col_stride, size_t n_cols, int log_n,
                              int rate_bits, int blinding, int cap_height, etp_batch **out);
int etp_batch_from_coeffs_dev(etp_ctx *ctx, const uint64_t *coeffs_dev, size_t col_stride, size_t n_cols, int log_n,
                              int rate_bits, int blinding, int cap_height, etp_batch **out);
/* re-run the commit into an existing batch of the same shape (no allocation; benches) */
int etp_batch_recommit_values_dev(etp_batch *b, const uint64_t *values_dev, size_t col_stride);
/* per-phase device times (ms, CUDA events on the context's stream) of the last commit into this batch,
 * named after plonky2's TimingTree scopes: [0] "IFFT", [1] "FFT + blinding" (coset LDE),
 * [2] "build Merkle tree": leaf hashing, [3] "build Merkle tree": inner levels + cap. */
int etp_batch_last_commit_timings(const etp_batch *b, float ms_out[4]);
void etp_batch_free(etp_batch *b);
size_t etp_batch_num_cols(const etp_batch *b);
int etp_batch_degree_log(const etp_batch *b);
size_t etp_batch_num_digests(const etp_batch *b);
/* merkle_tree.cap */
int etp_batch_cap(etp_batch *b, uint64_t *cap_out);
/* lazy host views of the public fields (plonky2 layouts): polynomials (n_cols x n, column-major),
 * merkle_tree.leaves ((n << rate_bits) x n_cols row-major, row i = point bitrev(i)), merkle_tree.digests */
int etp_batch_download_coeffs(etp_batch *b, uint64_t *out);
int etp_batch_download_leaves(etp_batch *b, uint64_t *out);
int etp_batch_download_digests(etp_batch *b, uint64_t *out);
/* merkle_tree.leaves[idx[q]] for q < n_idx: rows_out n_idx x n_cols */
int etp_batch_leaves_at(etp_batch *b, const uint64_t *idx, size_t n_idx, uint64_t *rows_out);
/* get_lde_values(index, step): leaves[bitrev(index * step)] */
int etp_batch_get_lde_values(etp_batch *b, size_t index, size_t step, uint64_t *row_out);
/* merkle_tree.prove(leaf_index) */
int etp_batch_prove(etp_batch *b, size_t leaf_index, uint64_t *siblings_out);
/* device views for fused callers: LDE matrix (column-major, column c at ptr + c*stride, bit-reversed
 * row order) and coefficients */
const uint64_t *etp_batch_lde_dev(const etp_batch *b, size_t *col_stride);
const uint64_t *etp_batch_coeffs_dev(const etp_batch *b, size_t *col_stride);

/* polynomials[c].to_extension().eval(z) for every polynomial of the batch (StarkOpeningSet::new's eval_commitment):
 * out has num_cols extension values (c0, c1 interleaved), canonical. */
int etp_batch_eval_at_ext_point(etp_batch *b, const uint64_t z[2], uint64_t *out);

/* ---- column-split commit of one oversized table across the GPUs of a box ------------------------
 * BASELINE.json north_star / SURVEY.md 8(e): PolynomialBatch::from_values of a table whose LDE does not
 * fit (or is too slow on) one GPU.  One process per GPU; rank g of `world` (a power of two <= 8 and
 * <= 2^cap_height) owns columns [g*cps, (g+1)*cps) with cps = etp_shard_cols_per_rank() and the leaf rows
 * [g*L/world, (g+1)*L/world), i.e. whole cap subtrees.  Protocol (every rank):
 *   etp_shard_create -> etp_shard_transform_values_{host,dev} (local iFFT + coset LDE)
 *   -> exchange etp_ipc_export(etp_shard_lde_dev) handles, etp_ipc_open + etp_shard_set_peer for each peer
 *   -> BARRIER -> etp_shard_commit_rows (leaf hashing reads the peers' columns over NVLink inside the
 *   kernel; own subtrees; returns the own 2^cap_height/world cap entries) -> all-gather the cap parts
 *   -> BARRIER before any LDE buffer is freed or overwritten.
 * The assembled cap, the Merkle paths and the rows equal those of the unsplit from_values. */
typedef struct etp_shard etp_shard;
#define ETP_IPC_HANDLE_BYTES 64
size_t etp_shard_cols_per_rank(size_t n_cols_total, int world);
int etp_shard_create(etp_ctx *ctx, size_t n_cols_total, int log_n, int rate_bits, int cap_height, int rank, int world,
                     etp_shard **out);
void etp_shard_free(etp_shard *s);
size_t etp_shard_first_col(const etp_shard *s);
size_t etp_shard_num_local_cols(const etp_shard *s);
size_t etp_shard_first_row(const etp_shard *s);
size_t etp_shard_num_rows(const etp_shard *s);
/* the local LDE matrix (local columns x (n << rate_bits), column-major, bit-reversed rows): what peers map */
const uint64_t *etp_shard_lde_dev(const etp_shard *s);
/* local_cols: etp_shard_num_local_cols() host pointers / one device matrix of the LOCAL columns */
int etp_shard_transform_values_host(etp_shard *s, const uint64_t *const *local_cols);
int etp_shard_transform_values_dev(etp_shard *s, const uint64_t *values_dev, size_t col_stride);
/* CUDA IPC plumbing for one-process-per-GPU peers (cudaIpcGetMemHandle / OpenMemHandle / CloseMemHandle) */
int etp_ipc_export(etp_ctx *ctx, const void *dev_ptr, unsigned char handle_out[ETP_IPC_HANDLE_BYTES]);
int etp_ipc_open(etp_ctx *ctx, const unsigned char handle[ETP_IPC_HANDLE_BYTES], void **dev_ptr_out);
int etp_ipc_close(etp_ctx *ctx, void *dev_ptr);
/* peer_lde: rank peer_rank's etp_shard_lde_dev as addressable from this context's device */
int etp_shard_set_peer(etp_shard *s, int peer_rank, const uint64_t *peer_lde);
/* cap_part_out: (2^cap_height / world) x 4 — entries [rank * 2^cap_height / world, ...) of merkle_tree.cap */
int etp_shard_commit_rows(etp_shard *s, uint64_t *cap_part_out);
/* merkle_tree.prove(leaf_index) for a leaf this rank owns: (log2(L) - cap_height) x 4 */
int etp_shard_prove(etp_shard *s, size_t leaf_index, uint64_t *siblings_out);
/* merkle_tree.leaves[idx[q]] (all n_cols_total columns, any owner): rows_out n_idx x n_cols_total */
int etp_shard_leaves_at(etp_shard *s, const uint64_t *idx, size_t n_idx, uint64_t *rows_out);
/* polynomials of the local columns (local x n, column-major) */
int etp_shard_download_coeffs(etp_shard *s, uint64_t *out);

/* ---- starky: tables, compute_quotient_polys, prove ------------------------------------------- */
#define ETP_TABLE_FIBONACCI 0 /* starky/src/fibonacci_stark.rs                                        */
#define ETP_TABLE_MEMORY 1    /* evm_arithmetization/src/memory/memory_stark.rs (shape) */
#define ETP_TABLE_FIRST_REGISTERED 16
/* ctx may be NULL for the built-in tables; registered tables belong to the context that registered them */
int etp_table_num_columns(const etp_ctx *ctx, int table);
int etp_table_constraint_degree(const etp_ctx *ctx, int table);
int etp_table_num_public_inputs(const etp_ctx *ctx, int table);
int etp_table_num_aux_columns(const etp_ctx *ctx, int table, int num_challenges);
int etp_table_quotient_degree_factor(const etp_ctx *ctx, int table);

/* Registers ANY starky table (the arithmetic, byte-packing, CPU, keccak, keccak-sponge, logic, memory STARKs of
 * evm_arithmetization, or a user's) from its constraint program: the table's `Stark::eval_packed_generic`
 * followed by `eval_packed_lookups_generic`, recorded once as straight-line code (format: csrc/cprog.h; the
 * patched starky records it with a symbolic PackedField, rust/etp_b200_sys).  The program is compiled for
 * sm_100a with NVRTC here, against the same field arithmetic as the built-in kernels, and the returned id
 * (>= ETP_TABLE_FIRST_REGISTERED, valid on this context) is accepted wherever a table id is.
 * lookups: starky::lookup::Lookup list without filters, flat int32:
 *   [n_lookups, then per lookup: table_column, frequencies_column, n_looking, looking columns...];
 * auxiliary columns are laid out as starky does: per lookup, per challenge: one helper column per chunk of
 * (constraint_degree - 1) looking columns, then Z. */
int etp_table_register(etp_ctx *ctx, const uint64_t *program, size_t n_words, const int32_t *lookups, size_t n_lookup_words,
                       int *table_id_out);
/* parse + NVRTC-compile a program without a device (CI / the Rust build): cubin size, or an error message */
int etp_cprog_compile_check(const uint64_t *program, size_t n_words, size_t *cubin_bytes_out, char *err, size_t err_len);
/* The CUDA source generated for a program (the text NVRTC compiles): one straight-line kernel, or — above 16384 ops, where ptxas
 * time on one function explodes — a chain of __noinline__ segment functions, each holding the dependencies of its own
 * constraints (ETP_CPROG_SEGMENT_OPS=<n> forces segments of n ops).  Returns the source length or a negative error; copies at
 * most cap - 1 bytes + terminator into `out` (may be NULL to query the length). */
int64_t etp_cprog_generate_cuda(const uint64_t *program, size_t n_words, char *out, size_t cap);

/* General registration.  `aux_spec` describes every auxiliary polynomial the prover must generate for this table
 * (starky/src/lookup.rs Lookup / Column / Filter, starky/src/cross_table_lookup.rs CtlZData) as u64 words:
 *   [0] magic "ETPAUXS1"  [1] n_lookups  [2] n_ctl_zs
 *   per lookup : n_columns, n_columns x Column, n_columns x Filter, table Column, frequencies Column
 *   per CTL Z  : challenge_index (< num_challenges), n_colsets, per colset: n_columns, Columns..., Filter
 *   Column     : n_local, (trace column, coefficient)*, n_next_row, (trace column, coefficient)*, constant
 *   Filter     : n_products, (Column, Column)*, n_constants, Column*          (Filter::default() = 0 products, [constant 1])
 * The CTL Z list is this table's `CtlData.zs_columns` in upstream's order (per cross-table lookup, per challenge; a table
 * that is looking k > 1 times in one CTL has k colsets in its entry, the looked table one).  Auxiliary polynomial order
 * (starky prover.rs): lookup columns [per lookup, per challenge: helpers, Z] ++ CTL helper columns ++ CTL Z columns.
 * The program must contain the table's constraints, its lookup checks and its CTL checks (eval_vanishing_poly order);
 * challenge scalars: CH 0..1 = lookup challenges, CH 2+2k / 3+2k = CTL (beta_k, gamma_k). */
int etp_table_register_ex(etp_ctx *ctx, const uint64_t *program, size_t n_words, const uint64_t *aux_spec, size_t n_spec_words,
                          int *table_id_out);
int etp_table_num_lookup_columns(const etp_ctx *ctx, int table, int num_challenges);
int etp_table_num_ctl_helper_columns(const etp_ctx *ctx, int table);
int etp_table_num_ctl_zs(const etp_ctx *ctx, int table);
/* All auxiliary polynomials of a table on the trace domain, on the device: lookup_helper_columns for every lookup and
 * challenge, then cross_table_lookup_data's helper columns and Z columns (partial_sums) for this table.
 * lookup_challenges: num_challenges scalars; ctl_challenges: num_challenges (beta, gamma) pairs or NULL (table without CTL).
 * aux_dev: etp_table_num_aux_columns columns of 2^log_n (column-major, stride 2^log_n).  A zero denominator
 * (upstream: "Tried to invert zero" panic) gives ETP_ERR_PROOF. */
int etp_aux_columns_dev(etp_ctx *ctx, int table, int log_n, const uint64_t *trace_dev, size_t col_stride,
                        const uint64_t *lookup_challenges, int num_challenges, const uint64_t *ctl_challenges, uint64_t *aux_dev);

/* starky::lookup::lookup_helper_columns for every lookup of the table and every challenge:
 * aux_dev gets etp_table_num_aux_columns columns of 2^log_n (column-major, stride 2^log_n). */
int etp_lookup_helper_columns_dev(etp_ctx *ctx, int table, int log_n, const uint64_t *trace_dev, size_t col_stride,
                                  const uint64_t *challenges, int n_challenges, uint64_t *aux_dev);
/* starky::prover::compute_quotient_polys: returns the quotient chunks
 * (quotient_degree_factor * n_alphas polynomials of 2^log_n coefficients, column-major) in out_dev. */
int etp_compute_quotient_polys_dev(etp_ctx *ctx, int table, etp_batch *trace, etp_batch *aux /* may be NULL */,
                                   const uint64_t *lookup_challenges /* challenge scalars: lookup challenges, then CTL (beta, gamma) pairs */,
                                   int n_lookup_challenges,
                                   const uint64_t *public_inputs, const uint64_t *alphas, int n_alphas,
                                   uint64_t *out_dev);

/* ---- plonky2 circuit prover: plonk::prover::prove after witness generation (plonky2 0.2.2 src/plonk/prover.rs; what every
 * shrink / root / aggregation / block proof of the reference runs: /root/reference/ops/src/lib.rs:52,72,95) ----------------
 * etp_circuit = CommonCircuitData + ProverOnlyCircuitData as far as the device steps need them, built once per circuit:
 *   vanishing_program : plonk/vanishing_poly.rs eval_vanishing_poly recorded as a constraint program (csrc/cprog.h) over the
 *                       virtual columns [constants | sigmas | wires | Zs | partial products (challenge-major) | X], terms emitted
 *                       in REVERSE order (reduce_with_powers weights term i with alpha^i), L_0(x)(Z - 1) as EMIT_FIRST_ROW,
 *                       Z(g x) as NV of the Z columns; CH 0..k-1 = betas, CH k..2k-1 = gammas; PI 0..3 = public_inputs_hash
 *   constants, sigmas : values on the subgroup, column-major (num_constants x n, num_routed_wires x n), host memory
 *   k_is              : the coset shifts of the routed columns
 *   fri_params        : e.g. etp_fri_params_make(degree_bits, 3, 4, 16, 28) for standard_recursion_config
 *   circuit_digest    : 4 words, or NULL for a stand-in (hash_no_pad(constants_sigmas cap ++ degree_bits))
 * etp_circuit_prove_*: wires = the witness, num_wires x n values column-major.  Transcript as upstream: circuit digest,
 * public_inputs_hash, wires cap -> betas, gammas; Z / partial-products cap -> alphas; quotient cap -> zeta; openings -> FRI.
 * Proof words "B200PLK1": header[24] = {magic, degree_bits, num_constants, num_routed_wires, num_wires, num_challenges,
 * num_partial_products, quotient_degree_factor, rate_bits, cap_height, n_fri_layers, arity_bits, final_poly_len, num_queries,
 * pow_bits, total_words}; wires_cap, plonk_zs_partial_products_cap, quotient_polys_cap; OpeningSet {constants, plonk_sigmas,
 * wires, plonk_zs, plonk_zs_next, partial_products, quotient_polys} (extension values); the flat FriProof over the oracles
 * [constants_sigmas, wires, zs_partial_products, quotient]; public_inputs_hash. */
typedef struct etp_circuit etp_circuit;
int etp_circuit_create(etp_ctx *ctx, const uint64_t *vanishing_program, size_t n_words, const uint64_t *constants,
                       int num_constants, const uint64_t *sigmas, const uint64_t *k_is, int num_routed_wires, int num_wires,
                       int degree_bits, int quotient_degree_factor, int num_challenges, const etp_fri_params *fri_params,
                       const uint64_t *circuit_digest, etp_circuit **out);
void etp_circuit_free(etp_circuit *c);
int etp_circuit_digest(const etp_circuit *c, uint64_t digest_out[4]);
int etp_circuit_constants_sigmas_cap(const etp_circuit *c, uint64_t *cap_out);
size_t etp_circuit_proof_words(const etp_circuit *c);
int etp_circuit_prove_host(etp_circuit *c, const uint64_t *wires, const uint64_t public_inputs_hash[4], uint64_t *proof_out);
int etp_circuit_prove_dev(etp_circuit *c, const uint64_t *wires_dev, size_t col_stride, const uint64_t public_inputs_hash[4],
                          uint64_t *proof_out);

/* Registers a constraint program that is NOT a starky table: constraint degree up to 9 (quotient degree factor up to 8), up
 * to 8 challenge scalars, no auxiliary polynomials, no StarkConfig limits — e.g. the vanishing polynomial of a plonky2
 * circuit (plonk/vanishing_poly.rs eval_vanishing_poly recorded over the virtual columns [constants | sigmas | wires | Zs |
 * partial products | X]).  The id is accepted by etp_compute_quotient_polys_cols_dev only. */
int etp_program_register(etp_ctx *ctx, const uint64_t *program, size_t n_words, int *table_id_out);
/* compute_quotient_polys over an arbitrary list of LDE columns: lde_cols[c] = device pointer of virtual trace column c
 * (bit-reversed rows, 2^(log_n + rate_bits) words), from any PolynomialBatch of this context.  The plonky2 circuit prover's
 * form (plonk/prover.rs compute_quotient_polys -> eval_vanishing_poly_base_batch): one program reads the constants / sigmas,
 * wires and Z / partial-product oracles and the LDE of X.  The table must have no auxiliary polynomials;
 * quotient_degree_factor <= 2^rate_bits and <= 8.  out_dev: n_alphas * quotient_degree_factor polynomials x n. */
int etp_compute_quotient_polys_cols_dev(etp_ctx *ctx, int table, const uint64_t *const *lde_cols, size_t n_cols, int log_n,
                                        int rate_bits, const uint64_t *challenge_scalars, int n_scalars,
                                        const uint64_t *public_inputs, const uint64_t *alphas, int n_alphas, uint64_t *out_dev);
/* plonky2::fri::prover::fri_proof_of_work: SMALLEST witness w such that the duplex response has
 * `bits` leading zeros; state = sponge state with the pending inputs already overwritten in,
 * pos = index the candidate is written to. */
int etp_pow_grind(etp_ctx *ctx, const uint64_t state[12], int pos, int bits, uint64_t *witness_out);

/* starky::prover::prove under StarkConfig::standard_fast_config(): trace commit, challenger observes the public inputs
 * and the trace cap, then prove_with_commitment (auxiliary columns, quotient, openings, FRI: commit phase, PoW, 84 query
 * rounds).  proof_out: etp_stark_proof_words() u64 in the flat wire format of DESIGN.md.  The trace (n_cols x 2^log_n column-major) is a host or a
 * device matrix. */
size_t etp_stark_proof_words(const etp_ctx *ctx, int table, int log_n);
int etp_stark_prove_host(etp_ctx *ctx, int table, int log_n, const uint64_t *trace, const uint64_t *public_inputs,
                         uint64_t *proof_out);
int etp_stark_prove_dev(etp_ctx *ctx, int table, int log_n, const uint64_t *trace_dev, size_t col_stride,
                        const uint64_t *public_inputs, uint64_t *proof_out);
/* per-phase device times (ms) of the last etp_stark_prove_* on this context, named after plonky2's
 * TimingTree scopes; returns the number of entries written (<= max). */
int etp_last_prove_timings(const etp_ctx *ctx, const char **names, float *ms, int max);

/* starky::prover::prove_with_commitment(stark, config, trace_poly_values, trace_commitment, ctl_data, ctl_challenges,
 * challenger, public_inputs, timing) under standard_fast_config — the call evm_arithmetization's prove_single_table makes
 * for each of the seven tables (/root/reference/ops/src/lib.rs:52 -> prove_with_traces).
 *   trace_commitment : PolynomialBatch::from_values(trace, rate_bits 1, blinding false, cap_height 4), already observed
 *                      by the caller's challenger (with every other table's cap);
 *   trace_dev        : the trace values (device, column-major), needed for the auxiliary columns;
 *   ctl_challenges   : num_challenges (beta, gamma) pairs from get_grand_product_challenge_set, or NULL for a table
 *                      outside any CTL (stand-alone prove: the lookup challenges are then drawn from the challenger).
 *                      The table's ctl_data (Z and helper columns) is computed here, on the device, from its registered
 *                      CtlZData descriptors;
 *   challenger       : transcript state, updated in place (state in -> state out).
 * proof_out: etp_stark_proof_words() u64, layout "B200STK2" (DESIGN.md): StarkProofWithPublicInputs field by field. */
int etp_prove_with_commitment(etp_ctx *ctx, int table, etp_batch *trace_commitment, const uint64_t *trace_dev, size_t col_stride,
                              const uint64_t *ctl_challenges, etp_challenger *challenger, const uint64_t *public_inputs,
                              uint64_t *proof_out);

/* ---- PolynomialBatch::prove_openings / plonky2::fri::prover::fri_proof for a general FriInstanceInfo -------------------
 * (plonky2/src/fri/oracle.rs, fri/prover.rs, fri/structure.rs).  This is what starky's prove_with_commitment and
 * plonky2's circuit prover (plonk/prover.rs, the recursion layers of /root/reference/ops/src/lib.rs:72,95) both end in. */
typedef struct etp_fri_poly { uint32_t oracle_index, polynomial_index; } etp_fri_poly;                 /* FriPolynomialInfo */
typedef struct etp_fri_batch { uint64_t point[2]; const etp_fri_poly *polynomials; size_t n_polynomials; } etp_fri_batch; /* FriBatchInfo */
/* words of a flat FriProof: commit_phase_merkle_caps (n_reductions x 2^cap x 4), query_round_proofs (num_query_rounds x
 * { per oracle: leaf row (n_cols), MerkleProof siblings x 4 ; per reduction: FriQueryStep.evals (2^arity ext — the
 * uncompressed FriProof carries all of them; only CompressedFriProof drops the queried one), siblings x 4 }),
 * final_poly (ext coefficients), pow_witness */
/* The quotients (F_b(x) - y_b) / (x - z_b) are formed point-wise on the LDE coset 7*H: an opening point that lies ON that coset
 * (never the case for a Fiat-Shamir zeta; upstream's coefficient-form division would accept it) gives ETP_ERR_PROOF. */
size_t etp_fri_proof_words(const size_t *oracle_num_cols, size_t n_oracles, const etp_fri_params *params);
int etp_prove_openings(etp_ctx *ctx, const etp_fri_batch *batches, size_t n_batches, etp_batch *const *oracles, size_t n_oracles,
                       etp_challenger *challenger, const etp_fri_params *params, uint64_t *fri_proof_out);

/* ---- plonky2's circuit prover, first slice (the recursion layers: /root/reference/ops/src/lib.rs:52,72,95) -----------------
 * plonky2::plonk::prover::all_wires_permutation_partial_products (wires_permutation_partial_products_and_zs per challenge,
 * util/partial_products.rs): the permutation argument's Z and partial-product polynomials, values on the subgroup (row i =
 * point g^i), computed on the device from the routed wires and the sigma values:
 *   quotient_j(i) = (wire_j(i) + beta k_j g^i + gamma) / (wire_j(i) + beta sigma_j(i) + gamma),
 *   chunk products over quotient_degree_factor wires, Z(g x) = Z(x) * (product of the row's chunks), Z(1) = 1.
 * wires_dev / sigmas_dev: num_routed_wires columns of 2^degree_bits (column-major, given strides); k_is: the coset shifts.
 * out_dev: num_challenges * (1 + num_partial_products) columns of 2^degree_bits, stride 2^degree_bits, in the order the
 * prover commits them: [Z of every challenge] ++ [partial products of challenge 0] ++ [of challenge 1] ...;
 * num_partial_products = ceil(num_routed_wires / quotient_degree_factor) - 1.  A zero denominator gives ETP_ERR_PROOF. */
int etp_plonk_partial_products_and_zs_dev(etp_ctx *ctx, const uint64_t *wires_dev, size_t wires_stride, const uint64_t *sigmas_dev,
                                          size_t sigmas_stride, const uint64_t *k_is, int num_routed_wires, int degree_bits,
                                          int quotient_degree_factor, const uint64_t *betas, const uint64_t *gammas, int num_challenges,
                                          uint64_t *out_dev);

/* ---- the FRI prover, step by step (fri_committed_trees is sequential through the challenger: beta_l depends on cap_l) --
 * values_dev: 2^(degree_bits + rate_bits) extension values (c0, c1 interleaved) of the polynomial on the coset 7*H in
 * BIT-REVERSED order (what reverse_index_bits_in_place gives upstream); the state takes a private copy. */
typedef struct etp_fri_state etp_fri_state;
int etp_fri_begin(etp_ctx *ctx, const uint64_t *values_dev, const etp_fri_params *params, etp_fri_state **out);
/* MerkleTree::new(chunked values of the current layer, cap_height) -> its cap (2^cap_height x 4) */
int etp_fri_commit_layer(etp_fri_state *s, uint64_t *cap_out);
/* fold the current layer with beta (arity 16) into the next one */
int etp_fri_fold(etp_fri_state *s, const uint64_t beta[2]);
/* after the last fold: the final polynomial's 2^(degree_bits - total arity bits) extension coefficients */
int etp_fri_final_poly(etp_fri_state *s, uint64_t *coeffs_out);
/* fused fri_committed_trees: every layer with the challenger (observe cap, draw beta), then observes the final polynomial.
 * caps_out: n_reductions caps; final_poly_out as above. */
int etp_fri_commit_phase(etp_fri_state *s, etp_challenger *challenger, uint64_t *caps_out, uint64_t *final_poly_out);
/* fri_proof_of_work: grinds the smallest witness for the challenger's current state, lets the challenger observe it and
 * draws the response (checked to have proof_of_work_bits leading zeros); the FRI query indices are the next
 * num_query_rounds challenges mod the LDE size. */
int etp_fri_proof_of_work(etp_ctx *ctx, etp_challenger *challenger, int proof_of_work_bits, uint64_t *witness_out);
/* fri_prover_query_rounds for the given x indices (< 2^(degree_bits + rate_bits)): out gets n_indices query rounds in the
 * layout of etp_fri_proof_words */
int etp_fri_query_rounds(etp_fri_state *s, etp_batch *const *oracles, size_t n_oracles, const uint64_t *x_indices, size_t n_indices,
                         uint64_t *out);
void etp_fri_free(etp_fri_state *s);

/* ---- proving a column-split table: quotient, openings and FRI on one rank (the "leader"), trace columns read
 * where they live (own HBM or a peer's over NVLink, through the etp_shard_set_peer mappings).  Protocol
 * (eth_tx_proof_b200/parallel.py prove_column_split; starky prover.rs prove_with_commitment): every rank commits the
 * shard -> leader: challenger, [etp_shard_aux_columns_dev, commits the auxiliary batch,] alphas, etp_shard_compute_quotient_polys_dev,
 * commits the quotient batch (a PolynomialBatch of its own), zeta -> every rank: etp_shard_eval_at_ext_points of its
 * columns at zeta and g*zeta, gathered -> leader: etp_shard_fri_begin (the combination step of prove_openings over
 * oracle 0 = the split table and oracles 1.. = its own batches), then the FRI prover step by step (etp_fri_*);
 * rows of the split table at the query indices: etp_shard_leaves_at; their Merkle paths: etp_shard_prove on the
 * ranks that own the leaves. */
/* All auxiliary polynomials of the table (values on the trace domain, order of etp_aux_columns_dev) computed on the
 * calling rank: the few trace columns the lookups / CTLs read are recovered from the mapped LDE (bit-reversed gather over
 * NVLink, coset iFFT, FFT).  aux_out_dev: num_aux_columns x n; ctl_zs_first_out: Z(1) per CTL Z (may be NULL).
 * A vanishing lookup / CTL denominator -> ETP_ERR_PROOF. */
int etp_shard_aux_columns_dev(etp_shard *s, int table, const uint64_t *lookup_challenges, int n_challenges,
                              const uint64_t *ctl_challenges, uint64_t *aux_out_dev, uint64_t *ctl_zs_first_out);
/* compute_quotient_polys over the split trace; aux: the committed auxiliary batch (NULL if the table has none);
 * challenge_scalars: lookup challenges, then the CTL (beta, gamma) pairs, as etp_compute_quotient_polys_dev;
 * out_dev: num_challenges * quotient_degree_factor polynomials x n. */
int etp_shard_compute_quotient_polys_dev(etp_shard *s, int table, etp_batch *aux, const uint64_t *challenge_scalars,
                                         int n_scalars, const uint64_t *public_inputs, const uint64_t *alphas, int n_alphas,
                                         uint64_t *out_dev);
/* polynomials of the LOCAL columns at z0 and z1 (eval_commitment): out0 / out1 get num_local_cols extension values */
int etp_shard_eval_at_ext_points(etp_shard *s, const uint64_t z0[2], const uint64_t z1[2], uint64_t *out0, uint64_t *out1);
/* FriPolynomialInfo.oracle_index 0 = the split table (polynomial_index = column of the whole table), k >= 1 =
 * extra_oracles[k - 1]; ys: the claimed openings, per batch and polynomial, extension values concatenated;
 * alpha: the challenge prove_openings draws.  Returns the FRI state holding the combined LDE values. */
int etp_shard_fri_begin(etp_shard *s, etp_batch *const *extra_oracles, size_t n_extra, const struct etp_fri_batch *batches,
                        size_t n_batches, const uint64_t *ys, const uint64_t alpha[2], const struct etp_fri_params *params,
                        etp_fri_state **out);

#ifdef __cplusplus
}
#endif
#endif
