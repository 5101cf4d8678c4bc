/* A plain C11 host that binds libetp_b200.so through include/etp_b200.h exactly as a cgo / Rust-FFI caller would (no Python,
 * no torch): compiled and linked by tests/test_cabi_c_host.py.  It exercises the entry points that need no GPU — the header
 * must parse as C, every symbol used must link, the plain-data structs must have the layout the header declares — and checks
 * the host Poseidon permutation against the first upstream known-answer vector (plonky2 poseidon_goldilocks.rs test_vectors:
 * the all-zero state), the Challenger against a by-hand duplex, and that every compute call fails loudly without a device.
 * Prints one line per check; exit code 0 = all passed. */
#include <inttypes.h>
#include <stdio.h>
#include <string.h>

#include "etp_b200.h"

static int failures = 0;
#define CHECK(cond, what) do { if (cond) printf("ok   %s\n", what); else { printf("FAIL %s\n", what); failures++; } } while (0)

/* argv[1..12]: the known-answer output for the all-zero state, hex (tests/golden/poseidon_kat.json) */
int main(int argc, char **argv) {
  uint64_t want0[12];
  if (argc != 13) { fprintf(stderr, "usage: host_only <12 hex words>\n"); return 2; }
  for (int i = 0; i < 12; i++) sscanf(argv[1 + i], "%" SCNx64, &want0[i]);
  CHECK(etp_version() != NULL && strlen(etp_version()) > 0, "etp_version");

  /* Poseidon: all-zero state -> upstream's first test vector */
  uint64_t s[12] = {0};
  etp_host_poseidon_permute(s);
  CHECK(memcmp(s, want0, sizeof s) == 0, "etp_host_poseidon_permute == upstream KAT (zero state)");

  /* PoseidonGate witness: outputs (wires 12..23) of inputs 0 == the permutation of 0 */
  uint64_t in[12] = {0}, wires[135];
  etp_host_poseidon_gate_wires(in, 0, wires);
  CHECK(memcmp(wires + 12, want0, sizeof want0) == 0, "etp_host_poseidon_gate_wires outputs == permutation");

  /* Challenger: observe 8 zeros -> one duplex -> challenges are popped from the END of the rate part of the KAT state */
  etp_challenger c;
  etp_challenger_init(&c);
  uint64_t zeros[8] = {0};
  etp_challenger_observe(&c, zeros, 8);
  const uint64_t ch0 = etp_challenger_get_challenge(&c), ch1 = etp_challenger_get_challenge(&c);
  CHECK(ch0 == want0[7] && ch1 == want0[6], "Challenger: duplex on 8 inputs, pop-from-end");
  etp_challenger_compact(&c);
  CHECK(c.input_len == 0 && c.output_len == 0 && memcmp(c.sponge_state, want0, sizeof want0) == 0, "Challenger::compact keeps the sponge state only");

  /* FriConfig::fri_params: standard_fast_config at 2^22 rows -> ConstantArityBits(4, 5): 22 -> 18 -> 14 -> 10 -> 6 (4 reductions) */
  etp_fri_params fp;
  CHECK(etp_fri_params_make(22, 1, 4, 16, 84, &fp) == 0 && fp.n_reductions == 4 && fp.reduction_arity_bits[0] == 4 &&
            fp.num_query_rounds == 84 && fp.proof_of_work_bits == 16,
        "etp_fri_params_make(standard_fast_config, 2^22)");
  CHECK(etp_fri_params_make(40, 1, 4, 16, 84, &fp) != 0, "etp_fri_params_make rejects degree_bits + rate_bits > 30");

  /* no device (or an absurd ordinal): creation fails loudly, nothing falls back to the CPU */
  etp_ctx *ctx = NULL;
  const int rc = etp_ctx_create(1 << 20, &ctx);
  CHECK(rc != 0 && ctx == NULL, "etp_ctx_create(bad device) fails, no CPU fallback");
  printf("%s\n", failures ? "FAILED" : "all C-host checks passed");
  return failures ? 1 : 0;
}
