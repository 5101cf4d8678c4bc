// placeholder: replaced by the full starky path (quotient, openings, FRI, prove)
#include "ctx.cuh"
extern "C" int etp_table_num_columns(int t) { return t == ETP_TABLE_FIBONACCI ? 2 : t == ETP_TABLE_MEMORY ? 21 : -1; }
extern "C" int etp_table_constraint_degree(int t) { return t == ETP_TABLE_FIBONACCI ? 2 : t == ETP_TABLE_MEMORY ? 3 : -1; }
extern "C" int etp_table_num_public_inputs(int t) { return t == ETP_TABLE_FIBONACCI ? 3 : t == ETP_TABLE_MEMORY ? 0 : -1; }
extern "C" int etp_table_num_aux_columns(int t, int nc) { return t == ETP_TABLE_MEMORY ? 2 * nc : 0; }
extern "C" int etp_table_quotient_degree_factor(int t) { int d = etp_table_constraint_degree(t) - 1; return d < 1 ? 1 : d; }
extern "C" int etp_lookup_helper_columns_dev(etp_ctx* ctx, int, int, const uint64_t*, size_t, const uint64_t*, int, uint64_t*) { return etp_fail(ctx, ETP_ERR_STATE, "not built"); }
extern "C" int etp_compute_quotient_polys_dev(etp_ctx* ctx, int, etp_batch*, etp_batch*, const uint64_t*, int, const uint64_t*, const uint64_t*, int, uint64_t*) { return etp_fail(ctx, ETP_ERR_STATE, "not built"); }
extern "C" int etp_pow_grind(etp_ctx* ctx, const uint64_t*, int, int, uint64_t*) { return etp_fail(ctx, ETP_ERR_STATE, "not built"); }
extern "C" size_t etp_stark_proof_words(int, int) { return 0; }
extern "C" int etp_stark_prove_host(etp_ctx* ctx, int, int, const uint64_t*, const uint64_t*, uint64_t*) { return etp_fail(ctx, ETP_ERR_STATE, "not built"); }
extern "C" int etp_stark_prove_dev(etp_ctx* ctx, int, int, const uint64_t*, size_t, const uint64_t*, uint64_t*) { return etp_fail(ctx, ETP_ERR_STATE, "not built"); }
extern "C" int etp_last_prove_timings(const etp_ctx*, const char**, float*, int) { return 0; }
