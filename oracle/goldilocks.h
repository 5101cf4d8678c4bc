/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this code, and only as the checker / CPU baseline.
 *
 * Goldilocks field F_p, p = 2^64 - 2^32 + 1, and its quadratic extension F_p[X]/(X^2 - 7).
 *
 * Restates (third-party, un-vendored; pinned by /root/reference/Cargo.lock:3466-3469,
 * plonky2_field 0.2.2):
 *   field/src/goldilocks_field.rs      GoldilocksField, reduce128, MULTIPLICATIVE_GROUP_GENERATOR = 7,
 *                                      POWER_OF_TWO_GENERATOR = 1753635133440165772
 *   field/src/extension/quadratic.rs   QuadraticExtension, W = 7
 * Reached from the reference at /root/reference/ops/src/lib.rs:52 (generate_txn_proof).
 *
 * All oracle functions return CANONICAL representatives (< p) and accept any u64.
 */
#ifndef ORACLE_GOLDILOCKS_H
#define ORACLE_GOLDILOCKS_H
#include <stdint.h>
#include <stddef.h>

#define GL_P 0xFFFFFFFF00000001ULL
#define GL_EPS 0xFFFFFFFFULL /* 2^64 mod p */
#define GL_GENERATOR 7ULL
#define GL_POWER_OF_TWO_GENERATOR 1753635133440165772ULL /* order 2^32 */
#define GL_TWO_ADICITY 32

static inline uint64_t gl_canon(uint64_t x) { return x >= GL_P ? x - GL_P : x; }

/* reduce128 as in goldilocks_field.rs: x = lo + 2^64*(hh*2^32 + hl) == lo - hh + hl*(2^32-1) (mod p) */
static inline uint64_t gl_reduce128(__uint128_t x) {
  uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
  uint64_t hh = hi >> 32, hl = hi & GL_EPS;
  uint64_t t0, t2;
  if (__builtin_sub_overflow(lo, hh, &t0)) t0 -= GL_EPS;
  uint64_t t1 = hl * GL_EPS;
  if (__builtin_add_overflow(t0, t1, &t2)) t2 += GL_EPS;
  return gl_canon(t2);
}
static inline uint64_t gl_add(uint64_t a, uint64_t b) {
  a = gl_canon(a); b = gl_canon(b);
  uint64_t s = a + b; /* < 2p - 1 < 2^65: detect wrap */
  if (s < a) s += GL_EPS; /* wrapped: +2^64 == +eps */
  return gl_canon(s);
}
static inline uint64_t gl_neg(uint64_t a) { a = gl_canon(a); return a ? GL_P - a : 0; }
static inline uint64_t gl_sub(uint64_t a, uint64_t b) { return gl_add(a, gl_neg(b)); }
static inline uint64_t gl_mul(uint64_t a, uint64_t b) { return gl_reduce128((__uint128_t)a * b); }
static inline uint64_t gl_sqr(uint64_t a) { return gl_mul(a, a); }
static inline uint64_t gl_pow(uint64_t a, uint64_t e) {
  uint64_t r = 1; a = gl_canon(a);
  while (e) { if (e & 1) r = gl_mul(r, a); a = gl_sqr(a); e >>= 1; }
  return r;
}
static inline uint64_t gl_inv(uint64_t a) { return gl_pow(a, GL_P - 2); }
/* primitive_root_of_unity(n_log) = POWER_OF_TWO_GENERATOR^(2^(32 - n_log)) */
static inline uint64_t gl_root_of_unity(int n_log) {
  uint64_t r = GL_POWER_OF_TWO_GENERATOR;
  for (int i = n_log; i < GL_TWO_ADICITY; i++) r = gl_sqr(r);
  return r;
}

/* ---- quadratic extension F_p[X]/(X^2 - 7): element = (c0, c1) = c0 + c1*X ---- */
typedef struct { uint64_t c0, c1; } gl2_t;
static inline gl2_t gl2(uint64_t a, uint64_t b) { gl2_t r = { gl_canon(a), gl_canon(b) }; return r; }
static inline gl2_t gl2_from_base(uint64_t a) { return gl2(a, 0); }
static inline gl2_t gl2_add(gl2_t a, gl2_t b) { return gl2(gl_add(a.c0, b.c0), gl_add(a.c1, b.c1)); }
static inline gl2_t gl2_sub(gl2_t a, gl2_t b) { return gl2(gl_sub(a.c0, b.c0), gl_sub(a.c1, b.c1)); }
static inline gl2_t gl2_neg(gl2_t a) { return gl2(gl_neg(a.c0), gl_neg(a.c1)); }
static inline gl2_t gl2_mul(gl2_t a, gl2_t b) {
  uint64_t c0 = gl_add(gl_mul(a.c0, b.c0), gl_mul(7, gl_mul(a.c1, b.c1)));
  uint64_t c1 = gl_add(gl_mul(a.c0, b.c1), gl_mul(a.c1, b.c0));
  return gl2(c0, c1);
}
static inline gl2_t gl2_scalar_mul(gl2_t a, uint64_t s) { return gl2(gl_mul(a.c0, s), gl_mul(a.c1, s)); }
static inline gl2_t gl2_inv(gl2_t a) {
  /* 1/(a0 + a1 X) = (a0 - a1 X) / (a0^2 - 7 a1^2) */
  uint64_t norm = gl_sub(gl_sqr(a.c0), gl_mul(7, gl_sqr(a.c1)));
  uint64_t ni = gl_inv(norm);
  return gl2(gl_mul(a.c0, ni), gl_mul(gl_neg(a.c1), ni));
}
static inline gl2_t gl2_pow(gl2_t a, uint64_t e) {
  gl2_t r = gl2(1, 0);
  while (e) { if (e & 1) r = gl2_mul(r, a); a = gl2_mul(a, a); e >>= 1; }
  return r;
}
static inline int gl2_eq(gl2_t a, gl2_t b) { return gl_canon(a.c0) == gl_canon(b.c0) && gl_canon(a.c1) == gl_canon(b.c1); }

static inline uint64_t bitrev64(uint64_t x, int bits) {
  uint64_t r = 0;
  for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
  return r;
}
#endif
