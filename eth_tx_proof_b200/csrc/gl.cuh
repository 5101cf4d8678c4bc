// Goldilocks field F_p, p = 2^64 - 2^32 + 1, and F_p[X]/(X^2 - 7) for sm_100a (K1 in SURVEY.md 2.4).
//
// Replaces plonky2_field 0.2.2 `GoldilocksField` / `QuadraticExtension<GoldilocksField>`
// (field/src/goldilocks_field.rs, field/src/extension/quadratic.rs; crate pinned at
// /root/reference/Cargo.lock:3466, reached from /root/reference/ops/src/lib.rs:52).
//
// Representation: one u64 per element. Like upstream, values in memory may be any u64
// ("non-canonical"); every device function accepts any u64 unless its comment says otherwise.
// Results that leave the library are canonicalised with gl_canon().
//
// 64x64->128 products use mul.lo/mul.hi.u64, which ptxas lowers to IMAD.WIDE.U32 chains on the
// FMA pipe; the reduction uses 2^64 == 2^32 - 1 and 2^96 == -1 (mod p) with carry-chain adds on the
// ALU pipe. Tensor cores are deliberately unused.
#pragma once
#if defined(__CUDACC_RTC__)  // NVRTC (constraint programs compiled at table registration): no host headers
typedef unsigned long long uint64_t;
typedef unsigned int uint32_t;
typedef int int32_t;
typedef unsigned long size_t;
#else
#include <cstdint>
#endif

#if defined(__CUDACC__)
#define GL_HD __host__ __device__ __forceinline__
#define GL_D __device__ __forceinline__
#else
#define GL_HD inline
#define GL_D inline
#endif

namespace gl {

constexpr uint64_t P = 0xFFFFFFFF00000001ULL;
constexpr uint64_t EPS = 0xFFFFFFFFULL;  // 2^64 mod p
constexpr uint64_t GENERATOR = 7ULL;     // MULTIPLICATIVE_GROUP_GENERATOR == coset_shift()
constexpr uint64_t POWER_OF_TWO_GENERATOR = 1753635133440165772ULL;
constexpr int TWO_ADICITY = 32;

GL_HD uint64_t canon(uint64_t x) { return x >= P ? x - P : x; }

#if defined(__CUDACC__)
static __constant__ int32_t MINUS_ONE = -1;  // read from the constant bank on purpose, see reduce_prod
#endif
#if defined(__CUDA_ARCH__)
// ---- device versions: explicit carry chains -------------------------------------------------
// a, b: any u64. Result: any u64, == a + b (mod p).
GL_D uint64_t add(uint64_t a, uint64_t b) {
  uint64_t s;
  uint32_t c;
  asm("add.cc.u64 %0, %2, %3;\n\taddc.u32 %1, 0, 0;" : "=l"(s), "=r"(c) : "l"(a), "l"(b));
  // +2^64 wrapped away == +EPS; the fix itself can wrap once more when both inputs were >= p
  asm("add.cc.u64 %0, %0, %2;\n\taddc.u32 %1, 0, 0;" : "+l"(s), "=r"(c) : "l"((uint64_t)(0u - c)));
  return s + (uint64_t)(0u - c);
}
// a: any u64, b: CANONICAL (< p) or at least one of the two < p. One fix-up suffices.
GL_D uint64_t add_c(uint64_t a, uint64_t b) {
  uint64_t s;
  uint32_t c;
  asm("add.cc.u64 %0, %2, %3;\n\taddc.u32 %1, 0, 0;" : "=l"(s), "=r"(c) : "l"(a), "l"(b));
  return s + (uint64_t)(0u - c);
}
// a, b: any u64. Result any u64 == a - b (mod p).
GL_D uint64_t sub(uint64_t a, uint64_t b) {
  uint64_t d;
  uint32_t br;
  asm("sub.cc.u64 %0, %2, %3;\n\tsubc.u32 %1, 0, 0;" : "=l"(d), "=r"(br) : "l"(a), "l"(b));
  // br = 0xffffffff on borrow == EPS as a u64
  asm("sub.cc.u64 %0, %0, %2;\n\tsubc.u32 %1, 0, 0;" : "+l"(d), "=r"(br) : "l"((uint64_t)br));
  return d - (uint64_t)br;
}
// add / sub with the second fix-up out of line.  The second wrap happens only when BOTH operands are >= p (add) or the
// subtrahend exceeds p + minuend (sub): never for canonical data, probability ~2^-64 for random representatives.  Keeping it
// behind a predicated call takes its two ALU instructions off the integer pipe that bounds the NTT (ptxas would if-convert
// an inline branch back into predicated instructions, which still occupy the pipe).
#if !defined(__CUDACC_RTC__)
static __device__ __noinline__ uint64_t add_eps_slow(uint64_t s) { return s + EPS; }
static __device__ __noinline__ uint64_t sub_eps_slow(uint64_t s) { return s - EPS; }
GL_D uint64_t add_r(uint64_t a, uint64_t b) {
  uint64_t s;
  uint32_t c, c2;
  asm("add.cc.u64 %0, %2, %3;\n\taddc.u32 %1, 0, 0;" : "=l"(s), "=r"(c) : "l"(a), "l"(b));
  asm("add.cc.u64 %0, %0, %2;\n\taddc.u32 %1, 0, 0;" : "+l"(s), "=r"(c2) : "l"((uint64_t)(0u - c)));
  if (__builtin_expect(c2 != 0, 0)) s = add_eps_slow(s);
  return s;
}
GL_D uint64_t sub_r(uint64_t a, uint64_t b) {
  uint64_t d;
  uint32_t br, br2;
  asm("sub.cc.u64 %0, %2, %3;\n\tsubc.u32 %1, 0, 0;" : "=l"(d), "=r"(br) : "l"(a), "l"(b));
  asm("sub.cc.u64 %0, %0, %2;\n\tsubc.u32 %1, 0, 0;" : "+l"(d), "=r"(br2) : "l"((uint64_t)br));
  if (__builtin_expect(br2 != 0, 0)) d = sub_eps_slow(d);
  return d;
}
#endif
// a: any u64, b: CANONICAL (< p).
GL_D uint64_t sub_c(uint64_t a, uint64_t b) {
  uint64_t d;
  uint32_t br;
  asm("sub.cc.u64 %0, %2, %3;\n\tsubc.u32 %1, 0, 0;" : "=l"(d), "=r"(br) : "l"(a), "l"(b));
  return d - (uint64_t)br;
}
// 128-bit (hi:lo) -> any u64 == value (mod p).  goldilocks_field.rs reduce128.
GL_D uint64_t reduce128(uint64_t lo, uint64_t hi) {
  uint32_t hl = (uint32_t)hi, hh = (uint32_t)(hi >> 32);
  uint64_t t0;
  uint32_t br, c;
  asm("sub.cc.u64 %0, %2, %3;\n\tsubc.u32 %1, 0, 0;" : "=l"(t0), "=r"(br) : "l"(lo), "l"((uint64_t)hh));
  t0 -= (uint64_t)br;                              // borrowed 2^64 == EPS: cannot underflow again
  uint64_t t1 = (uint64_t)hl * (uint64_t)0xFFFFFFFFu;  // IMAD.WIDE.U32 on the FMA pipe
  uint64_t t2;
  asm("add.cc.u64 %0, %2, %3;\n\taddc.u32 %1, 0, 0;" : "=l"(t2), "=r"(c) : "l"(t0), "l"(t1));
  return t2 + (uint64_t)(0u - c);                  // cannot overflow again (t1 <= EPS^2)
}
// same, but the result is CANONICAL: t0 + t1 < 2p, so one conditional subtraction of p is exact.
GL_D uint64_t reduce128_canon(uint64_t lo, uint64_t hi) {
  uint32_t hl = (uint32_t)hi, hh = (uint32_t)(hi >> 32);
  uint64_t t0;
  uint32_t br, c;
  asm("sub.cc.u64 %0, %2, %3;\n\tsubc.u32 %1, 0, 0;" : "=l"(t0), "=r"(br) : "l"(lo), "l"((uint64_t)hh));
  t0 -= (uint64_t)br;
  uint64_t t1 = (uint64_t)hl * (uint64_t)0xFFFFFFFFu;
  uint64_t t2;
  asm("add.cc.u64 %0, %2, %3;\n\taddc.u32 %1, 0, 0;" : "=l"(t2), "=r"(c) : "l"(t0), "l"(t1));
  return (c || t2 >= P) ? t2 + EPS : t2;           // -p == +EPS (mod 2^64)
}
#ifndef ETP_MUL_V
#define ETP_MUL_V 2
#endif
#if ETP_MUL_V == 2
// a, b: any u64.  Result any u64 == a*b (mod p).
// The ALU/FP64 issue port is what saturates in the Poseidon and NTT kernels while the FMA pipe idles
// (profiles/), so the multiplication keeps its carries on the FMA pipe: the 128-bit product comes from
// ptxas's own u128 lowering (IMAD.WIDE with carry-out / carry-in predicates), and the reduction
//   V = (w1:w0) + w2*EPS - w3,   -2^32 < V < 2^65 - 2^33
// is one wrapping multiply-add with carry c, one wrapping subtraction with borrow, k = c - borrow in
// {-1,0,1} and a single correction t + k*EPS, which can neither overflow (k = 1: t <= 2^64 - 2^33 - 1)
// nor underflow (k = -1: t >= 2^64 - 2^32 + 1).
GL_D uint64_t reduce_prod(uint64_t lo, uint64_t hi) {
  const uint32_t w0 = (uint32_t)lo, w1 = (uint32_t)(lo >> 32), w2 = (uint32_t)hi, w3 = (uint32_t)(hi >> 32);
  uint32_t t0, t1, k;
  asm("{\n\t"
      ".reg .u32 u0, u1, c;\n\t"
      "mad.lo.cc.u32 u0, %5, 0xffffffff, %3;\n\t"
      "madc.hi.cc.u32 u1, %5, 0xffffffff, %4;\n\t"
      "addc.u32 c, 0, 0;\n\t"
      "sub.cc.u32 %0, u0, %6;\n\t"
      "subc.cc.u32 %1, u1, 0;\n\t"
      "subc.u32 %2, c, 0;\n\t"
      "}"
      : "=r"(t0), "=r"(t1), "=r"(k)
      : "r"(w0), "r"(w1), "r"(w2), "r"(w3));
  // t + k*EPS = (t1 + k : t0) - sext(k).  The multiplier -1 comes from the constant bank so that ptxas
  // keeps the 64-bit multiply-add (one FMA-pipe IMAD.WIDE) instead of expanding it into ALU carry chains.
  const uint64_t t = ((uint64_t)(t1 + k) << 32) | t0;
  uint64_t r;
  asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(k), "r"(MINUS_ONE), "l"(t));
  return r;
}
// ETP_COMPACT_CODE (large NVRTC-compiled constraint programs): one out-of-line copy instead of thousands of
// inlined ones — the straight-line kernel would otherwise be megabytes of code and minutes of ptxas time.
#if defined(ETP_COMPACT_CODE)
static __device__ __noinline__ uint64_t mul(uint64_t a, uint64_t b) {
#else
GL_D uint64_t mul(uint64_t a, uint64_t b) {
#endif
  const unsigned __int128 p = (unsigned __int128)a * b;
  return reduce_prod((uint64_t)p, (uint64_t)(p >> 64));
}
GL_D uint64_t sqr(uint64_t a) { return mul(a, a); }
#else
// a, b: any u64.  Result any u64 == a*b (mod p).  Four IMAD.WIDE.U32 with a zero addend (2 clk each on
// B200; the accumulating form measures ~5 clk, tools/microbench/pipes.cu) + 32-bit carry chains, then
// the reduction x0 + 2^32 x1 + (2^32-1) x2 - x3 written out on 32-bit words.
GL_D uint64_t mul(uint64_t a, uint64_t b) {
  uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32);
  uint32_t r0, r1;
  asm("{\n\t"
      ".reg .u64 p00, p01, p10, p11, u;\n\t"
      ".reg .u32 w0, w1, w2, w3, l01, h01, l10, h10, l11, h11, t0, t1, br, u0, u1, c, m;\n\t"
      "mul.wide.u32 p00, %2, %4;\n\t"
      "mul.wide.u32 p01, %2, %5;\n\t"
      "mul.wide.u32 p10, %3, %4;\n\t"
      "mul.wide.u32 p11, %3, %5;\n\t"
      "mov.b64 {w0, w1}, p00;\n\t"
      "mov.b64 {l01, h01}, p01;\n\t"
      "mov.b64 {l10, h10}, p10;\n\t"
      "mov.b64 {l11, h11}, p11;\n\t"
      "add.cc.u32 w1, w1, l01;\n\t"
      "addc.cc.u32 w2, h01, l11;\n\t"
      "addc.u32 w3, h11, 0;\n\t"
      "add.cc.u32 w1, w1, l10;\n\t"
      "addc.cc.u32 w2, w2, h10;\n\t"
      "addc.u32 w3, w3, 0;\n\t"
      "sub.cc.u32 t0, w0, w3;\n\t"      // t = (w1:w0) - w3 ; a borrow of 2^64 is worth EPS
      "subc.cc.u32 t1, w1, 0;\n\t"
      "subc.u32 br, 0, 0;\n\t"
      "sub.cc.u32 t0, t0, br;\n\t"
      "subc.u32 t1, t1, 0;\n\t"
      "mul.wide.u32 u, w2, 0xffffffff;\n\t"  // w2 * EPS
      "mov.b64 {u0, u1}, u;\n\t"
      "add.cc.u32 %0, t0, u0;\n\t"
      "addc.cc.u32 %1, t1, u1;\n\t"
      "addc.u32 c, 0, 0;\n\t"
      "sub.u32 m, 0, c;\n\t"               // carry of 2^64 is worth EPS
      "add.cc.u32 %0, %0, m;\n\t"
      "addc.u32 %1, %1, 0;\n\t"
      "}"
      : "=r"(r0), "=r"(r1)
      : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
  return ((uint64_t)r1 << 32) | r0;
}
// a*a: the two cross products are the same IMAD.WIDE
GL_D uint64_t sqr(uint64_t a) {
  uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32);
  uint32_t r0, r1;
  asm("{\n\t"
      ".reg .u64 p00, p01, p10, p11, u;\n\t"
      ".reg .u32 w0, w1, w2, w3, l01, h01, l10, h10, l11, h11, t0, t1, br, u0, u1, c, m;\n\t"
      "mul.wide.u32 p00, %2, %2;\n\t"
      "mul.wide.u32 p01, %2, %3;\n\t"
      "mul.wide.u32 p11, %3, %3;\n\t"
      "mov.b64 {w0, w1}, p00;\n\t"
      "mov.b64 {l01, h01}, p01;\n\t"
      "mov.b64 {l10, h10}, p01;\n\t"
      "mov.b64 {l11, h11}, p11;\n\t"
      "add.cc.u32 w1, w1, l01;\n\t"
      "addc.cc.u32 w2, h01, l11;\n\t"
      "addc.u32 w3, h11, 0;\n\t"
      "add.cc.u32 w1, w1, l10;\n\t"
      "addc.cc.u32 w2, w2, h10;\n\t"
      "addc.u32 w3, w3, 0;\n\t"
      "sub.cc.u32 t0, w0, w3;\n\t"      // t = (w1:w0) - w3 ; a borrow of 2^64 is worth EPS
      "subc.cc.u32 t1, w1, 0;\n\t"
      "subc.u32 br, 0, 0;\n\t"
      "sub.cc.u32 t0, t0, br;\n\t"
      "subc.u32 t1, t1, 0;\n\t"
      "mul.wide.u32 u, w2, 0xffffffff;\n\t"  // w2 * EPS
      "mov.b64 {u0, u1}, u;\n\t"
      "add.cc.u32 %0, t0, u0;\n\t"
      "addc.cc.u32 %1, t1, u1;\n\t"
      "addc.u32 c, 0, 0;\n\t"
      "sub.u32 m, 0, c;\n\t"               // carry of 2^64 is worth EPS
      "add.cc.u32 %0, %0, m;\n\t"
      "addc.u32 %1, %1, 0;\n\t"
      "}"
      : "=r"(r0), "=r"(r1)
      : "r"(a0), "r"(a1));
  return ((uint64_t)r1 << 32) | r0;
}
#endif  // ETP_MUL_V
GL_D uint64_t mul_canon(uint64_t a, uint64_t b) { return canon(mul(a, b)); }
#else
// ---- host versions (used by the host-side Challenger / table setup; product code, not the oracle)
GL_HD uint64_t reduce128(uint64_t lo, uint64_t hi) {
  uint64_t hh = hi >> 32, hl = hi & EPS;
  uint64_t t0 = lo - hh;
  if (lo < hh) t0 -= EPS;
  uint64_t t1 = hl * EPS;
  uint64_t t2 = t0 + t1;
  if (t2 < t0) t2 += EPS;
  return t2;
}
GL_HD uint64_t reduce128_canon(uint64_t lo, uint64_t hi) { return canon(reduce128(lo, hi)); }
GL_HD uint64_t mul(uint64_t a, uint64_t b) {
  unsigned __int128 x = (unsigned __int128)a * b;
  return reduce128((uint64_t)x, (uint64_t)(x >> 64));
}
GL_HD uint64_t mul_canon(uint64_t a, uint64_t b) { return canon(mul(a, b)); }
GL_HD uint64_t add(uint64_t a, uint64_t b) {
  uint64_t s = a + b;
  if (s < a) { uint64_t t = s + EPS; s = t < s ? t + EPS : t; }
  return s;
}
GL_HD uint64_t add_c(uint64_t a, uint64_t b) { return add(a, b); }
GL_HD uint64_t sub(uint64_t a, uint64_t b) {
  uint64_t d = a - b;
  if (a < b) { uint64_t t = d - EPS; d = d < EPS ? t - EPS : t; }
  return d;
}
GL_HD uint64_t sub_c(uint64_t a, uint64_t b) { return sub(a, b); }
GL_HD uint64_t add_r(uint64_t a, uint64_t b) { return add(a, b); }
GL_HD uint64_t sub_r(uint64_t a, uint64_t b) { return sub(a, b); }
#endif

#if !defined(__CUDA_ARCH__)
GL_HD uint64_t sqr(uint64_t a) { return mul(a, a); }
#endif
GL_HD uint64_t neg(uint64_t a) { a = canon(a); return a ? P - a : 0; }
GL_HD uint64_t pow(uint64_t a, uint64_t e) {
  uint64_t r = 1;
  while (e) { if (e & 1) r = mul(r, a); a = sqr(a); e >>= 1; }
  return r;
}
GL_HD uint64_t inv(uint64_t a) { return pow(a, P - 2); }
// F::primitive_root_of_unity(n_log)
GL_HD uint64_t root_of_unity(int n_log) {
  uint64_t r = POWER_OF_TWO_GENERATOR;
  for (int i = n_log; i < TWO_ADICITY; i++) r = sqr(r);
  return canon(r);
}

// ---- quadratic extension, X^2 = 7 ---------------------------------------------------------------
struct Ext {
  uint64_t c0, c1;
};
GL_HD Ext ext(uint64_t a, uint64_t b) { Ext r; r.c0 = a; r.c1 = b; return r; }
GL_HD Ext eadd(Ext a, Ext b) { return ext(add(a.c0, b.c0), add(a.c1, b.c1)); }
GL_HD Ext esub(Ext a, Ext b) { return ext(sub(a.c0, b.c0), sub(a.c1, b.c1)); }
GL_HD Ext emul(Ext a, Ext b) {
  uint64_t t = mul(a.c1, b.c1);
  // 7*t = 8t - t, done with field ops to stay exact
  uint64_t t2 = add(t, t), t4 = add(t2, t2), t8 = add(t4, t4);
  return ext(add(mul(a.c0, b.c0), sub(t8, t)), add(mul(a.c0, b.c1), mul(a.c1, b.c0)));
}
GL_HD Ext emul_base(Ext a, uint64_t s) { return ext(mul(a.c0, s), mul(a.c1, s)); }
GL_HD Ext ecanon(Ext a) { return ext(canon(a.c0), canon(a.c1)); }
GL_HD Ext epow(Ext a, uint64_t e) {
  Ext r = ext(1, 0);
  while (e) { if (e & 1) r = emul(r, a); a = emul(a, a); e >>= 1; }
  return r;
}
GL_HD Ext einv(Ext a) {
  uint64_t t = mul(a.c1, a.c1);
  uint64_t t2 = add(t, t), t4 = add(t2, t2), t8 = add(t4, t4);
  uint64_t norm = sub(mul(a.c0, a.c0), sub(t8, t));
  uint64_t ni = inv(norm);
  return ext(mul(a.c0, ni), mul(neg(a.c1), ni));
}

GL_HD uint32_t bitrev32(uint32_t x, int bits) {
#if defined(__CUDA_ARCH__)
  return bits ? (__brev(x) >> (32 - bits)) : 0;
#else
  uint32_t r = 0;
  for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
  return r;
#endif
}

}  // namespace gl
