/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Single-column DFTs over Goldilocks, natural order in and out.  Restates plonky2_field 0.2.2
 * (/root/reference/Cargo.lock:3466; not on disk):
 *   field/src/fft.rs               fft_dispatch = reverse_index_bits + fft_classic (DIT);
 *                                  ifft = forward fft, then swap i <-> n-i and scale by 1/n
 *   field/src/polynomial/mod.rs    PolynomialCoeffs::{lde, coset_fft_with_options}, PolynomialValues::
 *                                  {ifft, coset_ifft}
 * out[k] = sum_j a[j] * w^(j k), w = primitive_root_of_unity(log_n).  Outputs are field elements and
 * therefore independent of the butterfly schedule; a plain radix-2 DIT is used.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

static void bit_reverse_inplace(uint64_t *a, int log_n) {
  size_t n = (size_t)1 << log_n;
  for (size_t i = 0; i < n; i++) {
    size_t j = bitrev64(i, log_n);
    if (i < j) { uint64_t t = a[i]; a[i] = a[j]; a[j] = t; }
  }
}

/* per-size twiddle cache: tw[log_n][k] = w_{2^log_n}^k, k < n/2 */
static uint64_t *tw_cache[33];
static const uint64_t *twiddles(int log_n) {
  uint64_t *t;
#pragma omp critical(orc_tw)
  {
    t = tw_cache[log_n];
    if (!t && log_n > 0) {
      size_t half = (size_t)1 << (log_n - 1);
      t = (uint64_t *)malloc(half * sizeof(uint64_t));
      uint64_t w = gl_root_of_unity(log_n), cur = 1;
      for (size_t k = 0; k < half; k++) { t[k] = cur; cur = gl_mul(cur, w); }
      tw_cache[log_n] = t;
    }
  }
  return t;
}

void orc_fft(uint64_t *a, int log_n) {
  size_t n = (size_t)1 << log_n;
  for (size_t i = 0; i < n; i++) a[i] = gl_canon(a[i]);
  if (log_n == 0) return;
  const uint64_t *tw = twiddles(log_n);
  bit_reverse_inplace(a, log_n);
  for (int s = 1; s <= log_n; s++) {
    size_t m = (size_t)1 << s, half = m >> 1, stride = n >> s;
    for (size_t k = 0; k < n; k += m)
      for (size_t j = 0; j < half; j++) {
        uint64_t u = a[k + j], v = gl_mul(a[k + j + half], tw[j * stride]);
        a[k + j] = gl_add(u, v);
        a[k + j + half] = gl_sub(u, v);
      }
  }
}

void orc_ifft(uint64_t *a, int log_n) {
  size_t n = (size_t)1 << log_n;
  orc_fft(a, log_n);
  uint64_t n_inv = gl_inv((uint64_t)n % GL_P);
  a[0] = gl_mul(a[0], n_inv);
  if (n > 1) a[n / 2] = gl_mul(a[n / 2], n_inv);
  for (size_t i = 1; i < n / 2; i++) {
    size_t j = n - i;
    uint64_t ci = gl_mul(a[j], n_inv), cj = gl_mul(a[i], n_inv);
    a[i] = ci; a[j] = cj;
  }
}

void orc_coset_fft(uint64_t *a, int log_n, uint64_t shift) {
  size_t n = (size_t)1 << log_n;
  uint64_t cur = 1;
  for (size_t i = 0; i < n; i++) { a[i] = gl_mul(a[i], cur); cur = gl_mul(cur, shift); }
  orc_fft(a, log_n);
}

void orc_coset_ifft(uint64_t *a, int log_n, uint64_t shift) {
  size_t n = (size_t)1 << log_n;
  orc_ifft(a, log_n);
  uint64_t si = gl_inv(shift), cur = 1;
  for (size_t i = 0; i < n; i++) { a[i] = gl_mul(a[i], cur); cur = gl_mul(cur, si); }
}

void orc_lde(const uint64_t *coeffs, int log_n, int rate_bits, uint64_t *out) {
  size_t n = (size_t)1 << log_n, big = n << rate_bits;
  memcpy(out, coeffs, n * sizeof(uint64_t));
  memset(out + n, 0, (big - n) * sizeof(uint64_t));
  orc_coset_fft(out, log_n + rate_bits, GL_GENERATOR);
}
