/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * PolynomialBatch::from_values / from_coeffs.  Restates plonky2 0.2.2
 * (/root/reference/Cargo.lock:3441; not on disk): plonky2/src/fri/oracle.rs
 *   from_values : ifft every column, then from_coeffs
 *   from_coeffs : lde_values (zero-pad to n << rate_bits, coset_fft(7)), transpose to rows,
 *                 reverse_index_bits_in_place on the rows, MerkleTree::new(rows, cap_height)
 * blinding is false on the STARK path (starky prove_with_commitment), so no salt columns.
 */
#include "oracle.h"
#include <omp.h>
#include <stdlib.h>
#include <string.h>

int orc_num_threads(void) { return omp_get_max_threads(); }

size_t orc_batch_num_digests(const orc_batch *b) {
  size_t n_leaves = (size_t)1 << (b->log_n + b->rate_bits);
  return 2 * (n_leaves - ((size_t)1 << b->cap_height));
}
const uint64_t *orc_batch_coeffs(const orc_batch *b) { return b->coeffs; }
const uint64_t *orc_batch_leaves(const orc_batch *b) { return b->leaves; }
const uint64_t *orc_batch_digests(const orc_batch *b) { return b->digests; }
const uint64_t *orc_batch_cap(const orc_batch *b) { return b->cap; }

static orc_batch *build(uint64_t *coeffs, size_t n_cols, int log_n, int rate_bits, int cap_height) {
  orc_batch *b = (orc_batch *)calloc(1, sizeof *b);
  size_t n = (size_t)1 << log_n, big = n << rate_bits;
  int log_big = log_n + rate_bits;
  b->n_cols = n_cols; b->log_n = log_n; b->rate_bits = rate_bits; b->cap_height = cap_height;
  b->coeffs = coeffs;
  b->leaves = (uint64_t *)malloc(big * (n_cols ? n_cols : 1) * sizeof(uint64_t));
  /* "FFT + blinding" then "transpose LDEs" + reverse_index_bits_in_place */
#pragma omp parallel
  {
    uint64_t *tmp = (uint64_t *)malloc(big * sizeof(uint64_t));
#pragma omp for schedule(dynamic)
    for (size_t c = 0; c < n_cols; c++) {
      orc_lde(coeffs + c * n, log_n, rate_bits, tmp);
      for (size_t k = 0; k < big; k++) b->leaves[bitrev64(k, log_big) * n_cols + c] = tmp[k];
    }
    free(tmp);
  }
  size_t nd = orc_batch_num_digests(b);
  b->digests = (uint64_t *)malloc((nd ? nd : 1) * 4 * sizeof(uint64_t));
  b->cap = (uint64_t *)malloc(((size_t)4 << cap_height) * sizeof(uint64_t));
  orc_merkle_new(b->leaves, big, n_cols, cap_height, b->digests, b->cap);
  return b;
}

orc_batch *orc_batch_from_coeffs(const uint64_t *coeffs, size_t n_cols, int log_n, int rate_bits, int cap_height) {
  size_t n = (size_t)1 << log_n;
  uint64_t *c = (uint64_t *)malloc((n_cols ? n_cols : 1) * n * sizeof(uint64_t));
  for (size_t i = 0; i < n_cols * n; i++) c[i] = gl_canon(coeffs[i]);
  return build(c, n_cols, log_n, rate_bits, cap_height);
}

orc_batch *orc_batch_from_values(const uint64_t *values, size_t n_cols, int log_n, int rate_bits, int cap_height) {
  size_t n = (size_t)1 << log_n;
  uint64_t *c = (uint64_t *)malloc((n_cols ? n_cols : 1) * n * sizeof(uint64_t));
  memcpy(c, values, n_cols * n * sizeof(uint64_t));
#pragma omp parallel for schedule(dynamic)
  for (size_t col = 0; col < n_cols; col++) orc_ifft(c + col * n, log_n);
  return build(c, n_cols, log_n, rate_bits, cap_height);
}

void orc_batch_free(orc_batch *b) {
  if (!b) return;
  free(b->coeffs); free(b->leaves); free(b->digests); free(b->cap); free(b);
}
