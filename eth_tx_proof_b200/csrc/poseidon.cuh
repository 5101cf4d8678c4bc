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
//  * MDS: the coefficients are < 2^6, so on the 32-bit halves of the lanes every 12-term sum is an
//    integer < 2^42 — exactly representable in binary64.  B200 issues DFMA at the same rate as
//    IMAD, and a DFMA accumulates for free (IMAD.WIDE with a 64-bit addend runs at ~5 clk/warp), so
//    the layer runs on the FP64 pipe, in split-cyclic form (cyclic(6) + negacyclic(6): 220 FP64
//    operations per layer instead of 312).  u32 -> f64 is the 2^52 bit trick; the next round's
//    constants and the bias are pre-folded into the accumulators' initial values (constant-bank
//    operands), and the integers come back as mantissa bits.  Every step is exact integer
//    arithmetic; no rounding occurs.
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
static __constant__ uint64_t RC_F64[30 * 24] = ETP_POSEIDON_RC_F64_TABLE;

__device__ __forceinline__ uint64_t sbox7(uint64_t x) {
  uint64_t x2 = gl::sqr(x);
  uint64_t x4 = gl::sqr(x2);
  uint64_t x3 = gl::mul(x, x2);
  return gl::mul(x3, x4);
}

// Two S-boxes behind one call.  The full rounds go through this out-of-line copy so that the whole
// permutation (full-round loop + paired partial-round loop) stays below the 32 KB instruction cache:
// with the eleven S-boxes inlined the kernel stalled on instruction fetch (profiles/, no_instruction).
// Arguments and results travel in registers (10 instructions of call overhead per pair).
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

// ---- MDS layer on the FP64 pipe, split-cyclic form -------------------------------------------------
// out = circ(C) * x (+ 8 x_0 on row 0) + next-round constants, on one 32-bit half ("plane") of the
// lanes.  With S_k = x_k + x_{k+6}, D_k = x_k - x_{k+6} (k < 6):
//   out_r + out_{r+6} = sum_k (C_k + C_{k+6}) S_{(k+r)%6}            (cyclic, length 6)
//   out_r - out_{r+6} = sum_k (C_k - C_{k+6}) (+-)D_{(k+r)%6}        (negacyclic: minus when k + r >= 6)
// and both coefficient vectors are even: (C_k + C_{k+6})/2 = {15,14,40,17,18,24},
// (C_k - C_{k+6})/2 = {2,1,1,-1,-16,4}.  72 DFMA + 38 DADD per plane instead of 144 + 12, all on
// integers below 2^53 (exact).  The u32 -> f64 conversions disappear into the S/D step: the double
// with bit pattern (0x43300000 : w) is 2^52 + w, so S_k = d_k + (d_{k+6} - 2^53), D_k = d_k - d_{k+6}.
// Accumulators start at 2^51 (+ constant), so out_r = Ah + Bh = 2^52 + value and the integer is read
// back from the mantissa; out_{r+6} = (Ah - Bh) + (2^52 + k_{r+6} - k_r).
struct PlaneOut {
  double o[12];
};
__device__ __forceinline__ void mds_plane(const double (&d)[12], const uint64_t* __restrict__ tab, double (&o)[12]) {
  constexpr double CA[6] = {15, 14, 40, 17, 18, 24};
  constexpr double CB[6] = {2, 1, 1, -1, -16, 4};
  double S[6], D[6];
#pragma unroll
  for (int k = 0; k < 6; k++) {
    const double e = d[k + 6] - 9007199254740992.0;  // 2^53
    S[k] = d[k] + e;
    D[k] = d[k] - d[k + 6];
  }
#pragma unroll
  for (int r = 0; r < 6; r++) {
    double ah = __longlong_as_double((long long)tab[r]);
    double bh = 2251799813685248.0;  // 2^51
#pragma unroll
    for (int k = 0; k < 6; k++) {
      ah = fma(S[(k + r) % 6], CA[k], ah);
      bh = fma(D[(k + r) % 6], (k + r >= 6) ? -CB[k] : CB[k], bh);
    }
    o[r] = ah + bh;
    o[r + 6] = (ah - bh) + __longlong_as_double((long long)tab[6 + r]);
  }
  o[0] = fma(d[0] - 4503599627370496.0, 8.0, o[0]);  // MDS_MATRIX_DIAG[0] = 8
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

// One MDS layer (+ the next round's constants) on both planes: d -> o, all values biased by 2^52.
__device__ __forceinline__ void mds_layer(const double (&dl)[12], const double (&dh)[12], int r, double (&ol)[12], double (&oh)[12]) {
  const uint64_t* __restrict__ tab = RC_F64 + 24 * r;
  mds_plane(dl, tab, ol);
  mds_plane(dh, tab + 12, oh);
}

// In-place permutation. Input lanes: any u64. Output lanes: any u64 (canonicalise before exporting).
//
// Full round r: S-boxes of lanes 1..11 ; MDS of round r (+ constants of round r+1) ; combine ; S-box of
// lane 0 for round r+1 (issued right behind row 0 so that it overlaps the rest).
// Partial rounds run in pairs: after the first MDS only lane 0 is recombined (its S-box needs the field
// element); lanes 1..11 enter the second MDS as the 41-bit plane values they are (the second layer's sums
// stay below 2^49, still exact), and all lanes are recombined once per pair.
__device__ __forceinline__ void permute(uint64_t (&s)[12]) {
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = gl::add_c(s[i], RC[i]);
  s[0] = sbox7(s[0]);
  int r = 0;
  ETP_ROLL
  for (int half = 0; half < 2; half++) {
    ETP_ROLL
    for (int i = 0; i < HALF_FULL; i++, r++) {
#pragma unroll
      for (int k = 1; k < 11; k += 2) {
        const ulonglong2 q = sbox7_pair(s[k], s[k + 1]);
        s[k] = q.x;
        s[k + 1] = q.y;
      }
      s[11] = sbox7(s[11]);
      double dl[12], dh[12], ol[12], oh[12];
#pragma unroll
      for (int k = 0; k < 12; k++) { dl[k] = plane_lo(s[k]); dh[k] = plane_hi(s[k]); }
      mds_layer(dl, dh, r, ol, oh);
      const uint64_t row0 = combine_planes(ol[0], oh[0]);
      const uint64_t next0 = sbox7(row0);
#pragma unroll
      for (int k = 1; k < 12; k++) s[k] = combine_planes(ol[k], oh[k]);
      s[0] = (r == ROUNDS - 1) ? row0 : next0;  // no S-box after the last round
    }
    if (half == 0) {
      ETP_ROLL
      for (int i = 0; i < PARTIAL / 2; i++, r += 2) {
        double dl[12], dh[12], ol[12], oh[12];
#pragma unroll
        for (int k = 0; k < 12; k++) { dl[k] = plane_lo(s[k]); dh[k] = plane_hi(s[k]); }
        mds_layer(dl, dh, r, ol, oh);
        const uint64_t mid0 = sbox7(combine_planes(ol[0], oh[0]));
        ol[0] = plane_lo(mid0);
        oh[0] = plane_hi(mid0);
        mds_layer(ol, oh, r + 1, dl, dh);
        const uint64_t row0 = combine_planes(dl[0], dh[0]);
        const uint64_t next0 = sbox7(row0);
#pragma unroll
        for (int k = 1; k < 12; k++) s[k] = combine_planes(dl[k], dh[k]);
        s[0] = next0;
      }
    }
  }
}
#endif  // __CUDACC__

}  // namespace poseidon
