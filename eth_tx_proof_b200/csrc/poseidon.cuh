// Poseidon-12 over Goldilocks for sm_100a: one thread per permutation, state in registers, round
// constants in constant memory (K4/K5 building block).
//
// Replaces plonky2 0.2.2 `impl Poseidon for GoldilocksField` / `PoseidonPermutation`
// (plonky2/src/hash/poseidon.rs, poseidon_goldilocks.rs) and the sponge helpers of
// plonky2/src/hash/hashing.rs (hash_n_to_m_no_pad, compress) — crate pinned at
// /root/reference/Cargo.lock:3441, reached from /root/reference/ops/src/lib.rs:52.
//
// Round structure (identical output to upstream's naive and "fast" forms): 4 full + 22 partial + 4
// full rounds; each round = add 12 constants, S-box x^7 (all lanes / lane 0), circulant MDS
//   out[r] = sum_i in[(i+r)%12]*CIRC[i] + in[r]*DIAG[r].
//
// B200 mapping (measured, tools/microbench/pipes.cu + profiles/):
//  * S-box: 4 Goldilocks multiplications, IMAD.WIDE (FMA pipe) + carry chains (ALU pipe).
//  * MDS: the coefficients are < 2^6, so on the 32-bit halves ("planes") of the lanes every 12-term sum is
//    an integer < 2^42 - exactly representable in binary64.  B200 issues DFMA at the same rate as IMAD and
//    a DFMA accumulates for free (IMAD.WIDE with a 64-bit addend runs at ~5 clk/warp), so the layer runs
//    on the FP64 pipe.  The circulant is applied through the factorisation x^12 - 1 = (x^3 - 1)(x^3 + 1)
//    (x^6 + 1): cyclic(3) + negacyclic(3) + negacyclic(6), 97 FP64 operations per plane instead of 156
//    (the upstream matrix was built for this: the cyclic(3) coefficients are {16, 16, 32}).  u32 -> f64 is
//    the 2^52 bit trick; the next round's constants and the bias are pre-folded into the accumulators'
//    start values (constant-bank operands, tools/gen_poseidon_constants.py), and the integers come back
//    as mantissa bits.
//  * Partial rounds run two at a time WITHOUT a second full layer: with M = C + 8 e0 e0^T, s the state
//    after the S-box of round r, u = (M s)_0 + k and n0 = sbox(u),
//        M (M s + e0 (n0 - u_lin)) = C^2 s + C[:,0] (8 s_0 + n0 - u_lin) + 8 e0 n0,
//    so a pair costs one 12-term dot product (u), one application of the circulant C^2 (coefficients
//    < 2^14: the sums stay below 2^50, still exact) and a rank-one update: 124 FP64 operations per plane
//    per pair instead of 2 x 110.  Every step is exact integer arithmetic; no rounding occurs
//    (tests/test_poseidon_f64_model.py replays the same data flow on Python integers and bounds every
//    intermediate below 2^52.2).
//  * One rolled loop over the 30 rounds (full-round S-boxes behind a warp-uniform branch) keeps the
//    hot code inside the instruction cache (the fully unrolled version stalled on instruction fetch).
// Tensor cores are deliberately unused (64-bit modular arithmetic).
#pragma once
#include "gl.cuh"
#include "poseidon_constants.h"

namespace poseidon {

constexpr int WIDTH = 12, RATE = 8, HALF_FULL = 4, PARTIAL = 22, ROUNDS = 30;
#define ETP_MDS_CIRC {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20}

#if defined(__CUDACC__)
static __constant__ uint64_t RC[360] = ETP_POSEIDON_RC_TABLE;
static __constant__ uint64_t FULL_F64[8 * 24] = ETP_POSEIDON_FULL_F64_TABLE;
static __constant__ uint64_t PAIR_F64[11 * 26] = ETP_POSEIDON_PAIR_F64_TABLE;

__device__ __forceinline__ uint64_t sbox7(uint64_t x) {
  uint64_t x2 = gl::sqr(x);
  uint64_t x4 = gl::sqr(x2);
  uint64_t x3 = gl::mul(x, x2);
  return gl::mul(x3, x4);
}

// Several S-boxes behind one call.  The full rounds go through an out-of-line copy so that the whole
// permutation (full-round loop + paired partial-round loop) stays below the 32 KB instruction cache:
// with the eleven S-boxes inlined the kernel stalled on instruction fetch (profiles/, no_instruction).
// Arguments and results travel in registers (~10 instructions of call overhead per call); the default is
// three 4-lane calls per full round (sbox7_quad), -DETP_SBOX_PAIRS selects the older five 2-lane calls.
#ifndef ETP_SBOX_PAIRS
#define ETP_SBOX_QUAD 1  // full-round S-boxes of all twelve lanes through out-of-line calls (1.3 % faster than five 2-lane calls + 2 inline)
// (-DETP_SBOX_HEX: two 6-lane calls per round; 0.4 % faster in the microbenchmark, 0.5 % slower in the leaf kernel, where
// the call needs one more live register than the 72 the occupancy allows)
#endif
#if defined(ETP_COUNT_UNROLL)  // instruction-count builds (tools/sass_count.py): everything inline and unrolled
#define ETP_ROLL _Pragma("unroll")
#define ETP_SBOX_PAIR_ATTR __forceinline__
#else
#define ETP_ROLL _Pragma("unroll 1")
#define ETP_SBOX_PAIR_ATTR __noinline__
#endif
static __device__ ETP_SBOX_PAIR_ATTR ulonglong2 sbox7_pair(uint64_t a, uint64_t b) {
  ulonglong2 r;
  r.x = sbox7(a);
  r.y = sbox7(b);
  return r;
}
#if defined(ETP_SBOX_QUAD)
// four S-boxes behind one call (fewer calls per full round: 3 instead of 5 + 2 inlined S-boxes)
static __device__ ETP_SBOX_PAIR_ATTR ulonglong4 sbox7_quad(uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
  ulonglong4 r;
  r.x = sbox7(a);
  r.y = sbox7(b);
  r.z = sbox7(c);
  r.w = sbox7(d);
  return r;
}
#endif
#if defined(ETP_SBOX_HEX)
struct U64x6 { uint64_t v[6]; };
static __device__ ETP_SBOX_PAIR_ATTR U64x6 sbox7_hex(uint64_t a, uint64_t b, uint64_t c, uint64_t d, uint64_t e, uint64_t f) {
  U64x6 r;
  r.v[0] = sbox7(a); r.v[1] = sbox7(b); r.v[2] = sbox7(c); r.v[3] = sbox7(d); r.v[4] = sbox7(e); r.v[5] = sbox7(f);
  return r;
}
#endif

// ---- MDS layer on the FP64 pipe, three-level split ---------------------------------------------------
// One plane (32-bit half) of the lanes, d_k = 2^52 + x_k.  With
//   S_k = x_k + x_{k+6}, D_k = x_k - x_{k+6} (k < 6);  T_j = S_j + S_{j+3}, E_j = S_j - S_{j+3} (j < 3)
// the circulant out_r = sum_i x_{(i+r)%12} c_i becomes
//   P_j = sum_k CP_k T_{(k+j)%3}            cyclic(3);  CP has two equal entries a and one other:
//                                           P_j = a (T_0 + T_1 + T_2) + (CP_K - a) T_{(K+j)%3}
//   Q_j = sum_k +-CQ_k E_{(k+j)%3}          negacyclic(3), minus when k + j >= 3
//   B_r = sum_k +-CB_k D_{(k+r)%6}          negacyclic(6), minus when k + r >= 6
//   A_j = P_j + Q_j, A_{j+3} = P_j - Q_j ;  out_r = A_r + B_r, out_{r+6} = A_r - B_r
// (coefficients for c = MDS_MATRIX_CIRC and for c * c in poseidon_constants.h).  The chains start from
// table values that carry the 2^52 bias (on P) and the next round's constants, chosen == 0 (mod 4) per
// plane so that the halvings above stay integral.  S_k is formed as d_k + (d_{k+6} - 2^53): exact.
struct Chains {
  double P[3], Q[3], B[6];
};
struct Pre {
  double S[6], D[6], T[3], E[3];
};
__device__ __forceinline__ void split_pre(const double (&d)[12], Pre& v) {
#pragma unroll
  for (int k = 0; k < 6; k++) {
    const double e = d[k + 6] - 9007199254740992.0;  // 2^53
    v.S[k] = d[k] + e;
    v.D[k] = d[k] - d[k + 6];
  }
#pragma unroll
  for (int j = 0; j < 3; j++) {
    v.T[j] = v.S[j] + v.S[j + 3];
    v.E[j] = v.S[j] - v.S[j + 3];
  }
}
__device__ __forceinline__ double tab_f64(const uint64_t* __restrict__ tab, int i) { return __longlong_as_double((long long)tab[i]); }

template <int LV>  // 1: the circulant C, 2: C * C
__device__ __forceinline__ void split_chains(const Pre& v, const uint64_t* __restrict__ tab, Chains& c) {
  constexpr double CB1[6] = ETP_MDS_C1_CB, CB2[6] = ETP_MDS_C2_CB;
  constexpr double CQ1[3] = ETP_MDS_C1_CQ, CQ2[3] = ETP_MDS_C2_CQ;
  constexpr double CPA = LV == 1 ? ETP_MDS_C1_CPA : ETP_MDS_C2_CPA, CPD = LV == 1 ? ETP_MDS_C1_CPD : ETP_MDS_C2_CPD;
  constexpr int CPK = LV == 1 ? ETP_MDS_C1_CPK : ETP_MDS_C2_CPK;
  const double sum_t = (v.T[0] + v.T[1]) + v.T[2];
#pragma unroll
  for (int j = 0; j < 3; j++) {
    c.P[j] = fma(sum_t, CPA, fma(v.T[(CPK + j) % 3], CPD, tab_f64(tab, j)));
    double q = tab_f64(tab, 3 + j);
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const double cq = LV == 1 ? CQ1[k] : CQ2[k];
      q = fma(v.E[(k + j) % 3], (k + j >= 3) ? -cq : cq, q);
    }
    c.Q[j] = q;
  }
#pragma unroll
  for (int r = 0; r < 6; r++) {
    double b = tab_f64(tab, 6 + r);
#pragma unroll
    for (int k = 0; k < 6; k++) {
      const double cb = LV == 1 ? CB1[k] : CB2[k];
      b = fma(v.D[(k + r) % 6], (k + r >= 6) ? -cb : cb, b);
    }
    c.B[r] = b;
  }
}
__device__ __forceinline__ void split_post(const Chains& c, double (&o)[12]) {
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const double a0 = c.P[j] + c.Q[j], a1 = c.P[j] - c.Q[j];
    o[j] = a0 + c.B[j];
    o[j + 6] = a0 - c.B[j];
    o[j + 3] = a1 + c.B[j + 3];
    o[j + 9] = a1 - c.B[j + 3];
  }
}

// full-round layer: o = 2^52 + circ(C) x + 8 x_0 e0 + next round's constants.  tab: 12 start values.
__device__ __forceinline__ void mds_plane(const double (&d)[12], const uint64_t* __restrict__ tab, double (&o)[12]) {
  Pre v;
  Chains c;
  split_pre(d, v);
  split_chains<1>(v, tab, c);
  split_post(c, o);
  o[0] = fma(d[0] - 4503599627370496.0, 8.0, o[0]);  // MDS_MATRIX_DIAG[0] = 8
}

// Pair of partial rounds, phase 1: u = 2^52 + (M x)_0 + k (row 0 only: 12 DFMA on S and D, the diagonal
// 8 x_0 = 4 S_0 + 4 D_0 folded into two coefficients) and the chains of C^2 x, which do not depend on the
// S-box that follows.  tab: u start | 12 start values.
__device__ __forceinline__ double pair_phase1(const double (&d)[12], const uint64_t* __restrict__ tab, Chains& c) {
  constexpr int C[12] = ETP_MDS_CIRC;
  Pre v;
  split_pre(d, v);
  double u = tab_f64(tab, 0);
#pragma unroll
  for (int k = 0; k < 6; k++) {
    u = fma(v.S[k], (double)((C[k] + C[k + 6]) / 2 + (k == 0 ? 4 : 0)), u);
    u = fma(v.D[k], (double)((C[k] - C[k + 6]) / 2 + (k == 0 ? 4 : 0)), u);
  }
  split_chains<2>(v, tab + 1, c);
  return u;
}
// phase 2: nd = 2^52 + plane of n0 = sbox(u).  z = 8 x_0 + n0 - u_lin enters through column 0 of C
// (split like the outputs), and lane 0 receives 8 n0.
__device__ __forceinline__ void pair_phase2(double d0, double ud, double nd, Chains& c, double (&o)[12]) {
  constexpr double PV[3] = ETP_MDS_COL0_PV, QV[3] = ETP_MDS_COL0_QV, BV[6] = ETP_MDS_COL0_BV;
  const double z = fma(d0 - 4503599627370496.0, 8.0, nd - ud);
#pragma unroll
  for (int j = 0; j < 3; j++) {
    c.P[j] = fma(z, PV[j], c.P[j]);
    c.Q[j] = fma(z, QV[j], c.Q[j]);
  }
#pragma unroll
  for (int r = 0; r < 6; r++) c.B[r] = fma(z, BV[r], c.B[r]);
  split_post(c, o);
  o[0] = fma(nd - 4503599627370496.0, 8.0, o[0]);
}

// (2^52 + L, 2^52 + H) -> L + H * 2^32 (mod p) for L, H < 2^52.  With L = L1 * 2^32 + L0, H = h1 * 2^32 + h0
// and M = L1 + h0 + h1 = c * 2^32 + m the value is ((m + c) : L0) - (c + h1), which can never borrow.
// The high words of the doubles are 0x43300000 + L1 / + h1, so the exponent bits are removed by the
// constants folded into the additions instead of being masked: 7 integer instructions.
__device__ __forceinline__ uint64_t combine_planes(double al, double ah) {
  const uint32_t L0 = (uint32_t)__double2loint(al), wL = (uint32_t)__double2hiint(al);
  const uint32_t h0 = (uint32_t)__double2loint(ah), wH = (uint32_t)__double2hiint(ah);
  uint32_t v0, v1;
  asm("{\n\t"
      ".reg .u32 x, m, mc, t;\n\t"
      "add.u32 x, %2, %3;\n\t"
      "add.u32 x, x, 0x79a00000;\n\t"     // - 2 * 0x43300000: x = L1 + h1
      "add.cc.u32 m, x, %4;\n\t"          // + h0, carry c
      "addc.u32 mc, m, 0;\n\t"            // m + c
      "addc.u32 t, %3, 0xbcd00000;\n\t"   // h1 + c
      "sub.cc.u32 %0, %5, t;\n\t"
      "subc.u32 %1, mc, 0;\n\t"
      "}"
      : "=r"(v0), "=r"(v1)
      : "r"(wL), "r"(wH), "r"(h0), "r"(L0));
  return ((uint64_t)v1 << 32) | v0;
}

__device__ __forceinline__ double plane_lo(uint64_t x) { return __hiloint2double(0x43300000, (int)(uint32_t)x); }
__device__ __forceinline__ double plane_hi(uint64_t x) { return __hiloint2double(0x43300000, (int)(uint32_t)(x >> 32)); }

// In-place permutation. Input lanes: any u64. Output lanes: any u64 (canonicalise before exporting).
//
// Full round r: S-boxes of lanes 1..11 ; MDS of round r (+ constants of round r+1) ; combine ; S-box of
// lane 0 for round r+1 (issued right behind row 0 so that it overlaps the rest).
// Partial rounds (r, r+1): phase 1 (u and the C^2 chains) ; combine u ; S-box ; phase 2 ; combine all lanes
// once per pair ; S-box of lane 0 for round r+2.
__device__ __forceinline__ void permute(uint64_t (&s)[12]) {
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = gl::add_c(s[i], RC[i]);
#if !defined(ETP_SBOX_QUAD)
  s[0] = sbox7(s[0]);
#endif
  int f = 0;
  ETP_ROLL
  for (int half = 0; half < 2; half++) {
    ETP_ROLL
    for (int i = 0; i < HALF_FULL; i++, f++) {
#if defined(ETP_SBOX_HEX)
#pragma unroll
      for (int k = 0; k < 12; k += 6) {
        const U64x6 q = sbox7_hex(s[k], s[k + 1], s[k + 2], s[k + 3], s[k + 4], s[k + 5]);
#pragma unroll
        for (int j = 0; j < 6; j++) s[k + j] = q.v[j];
      }
#elif defined(ETP_SBOX_QUAD)
      // lane 0 arrives here WITHOUT its S-box (see below): all twelve lanes go through three quad calls
#pragma unroll
      for (int k = 0; k < 12; k += 4) {
        const ulonglong4 q = sbox7_quad(s[k], s[k + 1], s[k + 2], s[k + 3]);
        s[k] = q.x; s[k + 1] = q.y; s[k + 2] = q.z; s[k + 3] = q.w;
      }
#else
#pragma unroll
      for (int k = 1; k < 11; k += 2) {
        const ulonglong2 q = sbox7_pair(s[k], s[k + 1]);
        s[k] = q.x;
        s[k + 1] = q.y;
      }
      s[11] = sbox7(s[11]);
#endif
      double dl[12], dh[12], ol[12], oh[12];
#pragma unroll
      for (int k = 0; k < 12; k++) { dl[k] = plane_lo(s[k]); dh[k] = plane_hi(s[k]); }
      const uint64_t* __restrict__ tab = FULL_F64 + 24 * f;
      mds_plane(dl, tab, ol);
      mds_plane(dh, tab + 12, oh);
#if defined(ETP_SBOX_QUAD)
#pragma unroll
      for (int k = 0; k < 12; k++) s[k] = combine_planes(ol[k], oh[k]);
      // the partial rounds expect lane 0 with its S-box applied: only after the last of the first four full rounds
      if (f == HALF_FULL - 1) s[0] = sbox7(s[0]);
#else
      const uint64_t row0 = combine_planes(ol[0], oh[0]);
      const uint64_t next0 = sbox7(row0);
#pragma unroll
      for (int k = 1; k < 12; k++) s[k] = combine_planes(ol[k], oh[k]);
      s[0] = (f == 2 * HALF_FULL - 1) ? row0 : next0;  // no S-box after the last round
#endif
    }
    if (half == 0) {
      ETP_ROLL
      for (int i = 0; i < PARTIAL / 2; i++) {
        double dl[12], dh[12], ol[12], oh[12];
#pragma unroll
        for (int k = 0; k < 12; k++) { dl[k] = plane_lo(s[k]); dh[k] = plane_hi(s[k]); }
        const uint64_t* __restrict__ tab = PAIR_F64 + 26 * i;
        Chains cl, ch;
        const double ul = pair_phase1(dl, tab, cl);
        const double uh = pair_phase1(dh, tab + 13, ch);
        const uint64_t mid0 = sbox7(combine_planes(ul, uh));
        pair_phase2(dl[0], ul, plane_lo(mid0), cl, ol);
        pair_phase2(dh[0], uh, plane_hi(mid0), ch, oh);
        const uint64_t row0 = combine_planes(ol[0], oh[0]);
#if defined(ETP_SBOX_QUAD)
        uint64_t next0 = row0;  // the full round that follows the last pair applies lane 0's S-box itself
        if (i != PARTIAL / 2 - 1) next0 = sbox7(row0);
#else
        const uint64_t next0 = sbox7(row0);
#endif
#pragma unroll
        for (int k = 1; k < 12; k++) s[k] = combine_planes(ol[k], oh[k]);
        s[0] = next0;
      }
    }
  }
}
#endif  // __CUDACC__

}  // namespace poseidon
