/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * First slice of plonky2's circuit prover (the recursion layers the reference runs at
 * /root/reference/ops/src/lib.rs:52,72,95).  Restates plonky2 0.2.2 (/root/reference/Cargo.lock:3441; not on disk):
 *   plonky2/src/plonk/prover.rs       all_wires_permutation_partial_products, wires_permutation_partial_products_and_zs
 *   plonky2/src/util/partial_products.rs  quotient_chunk_products, partial_products_and_z_gx
 * Parity unpinned (no upstream vector); the permutation property Z(g^n) = 1 on a consistent witness is checked by the tests.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

/* wires, sigmas: num_routed column-major (n each); out: num_challenges * (1 + num_partial_products) columns of n in the order
 * the prover commits them: [Z per challenge] ++ [partial products of challenge 0] ++ [challenge 1] ...  Returns 1 on a zero
 * denominator (upstream panics). */
int orc_plonk_partial_products_and_zs(const uint64_t *wires, const uint64_t *sigmas, const uint64_t *k_is, int num_routed,
                                      int degree_bits, int quotient_degree_factor, const uint64_t *betas, const uint64_t *gammas,
                                      int num_challenges, uint64_t *out) {
  const size_t n = (size_t)1 << degree_bits;
  const int n_chunks = (num_routed + quotient_degree_factor - 1) / quotient_degree_factor, n_pp = n_chunks - 1;
  const uint64_t g = gl_root_of_unity(degree_bits);
  uint64_t *q = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)n_chunks);
  for (int c = 0; c < num_challenges; c++) {
    const uint64_t beta = gl_canon(betas[c]), gamma = gl_canon(gammas[c]);
    uint64_t *z = out + (size_t)c * n;
    uint64_t *pp = out + ((size_t)num_challenges + (size_t)c * n_pp) * n;
    uint64_t z_x = 1, x = 1;
    for (size_t i = 0; i < n; i++) {
      /* quotient_values, then quotient_chunk_products(.., max_degree) */
      for (int k = 0; k < n_chunks; k++) q[k] = 1;
      for (int j = 0; j < num_routed; j++) {
        const uint64_t w = wires[(size_t)j * n + i];
        const uint64_t num = gl_add(gl_add(w, gl_mul(beta, gl_mul(k_is[j], x))), gamma);
        const uint64_t den = gl_add(gl_add(w, gl_mul(beta, sigmas[(size_t)j * n + i])), gamma);
        if (den == 0) { free(q); return 1; }
        q[j / quotient_degree_factor] = gl_mul(q[j / quotient_degree_factor], gl_mul(num, gl_inv(den)));
      }
      /* partial_products_and_z_gx(z_x, chunks); the last entry is Z(g x): swapped with Z(x) */
      z[i] = z_x;
      uint64_t acc = z_x;
      for (int k = 0; k < n_chunks; k++) {
        acc = gl_mul(acc, q[k]);
        if (k < n_pp) pp[(size_t)k * n + i] = acc;
      }
      z_x = acc;
      x = gl_mul(x, g);
    }
  }
  free(q);
  return 0;
}
