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
//    IMAD and on its own pipe, and a DFMA accumulates for free (IMAD.WIDE with a 64-bit addend
//    runs at ~5 clk/warp), so the layer is 2 x 144 DFMA on the FP64 pipe, overlapping the integer
//    pipes.  u32 -> f64 is the 2^52 trick (one DADD); the next round's constants and the 2^52 bias
//    are pre-folded into the accumulator's initial value (constant-bank operand), and the integer
//    comes back as the mantissa bits.  Every step is exact integer arithmetic; no rounding occurs.
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
  uint64_t x2 = gl::mul(x, x);
  uint64_t x4 = gl::mul(x2, x2);
  uint64_t x3 = gl::mul(x, x2);
  return gl::mul(x3, x4);
}

// exact u32 -> f64: the double with bit pattern (0x43300000 : x) is 2^52 + x
__device__ __forceinline__ double u32_to_f64(uint32_t x) {
  return __hiloint2double(0x43300000, (int)x) - 4503599627370496.0;
}

// One MDS row on the FP64 pipe: sum_i x[(i+r)%12]*C[i] (+ 8*x[0] for r = 0) + next-round constant.
template <int R>
__device__ __forceinline__ uint64_t mds_row_f64(const double (&xl)[12], const double (&xh)[12],
                                                const uint64_t* __restrict__ rcd) {
  constexpr double C[12] = ETP_MDS_CIRC;
  // initial value = 2^52 + constant half (f64 bit pattern from the constant bank)
  double al = __longlong_as_double((long long)rcd[2 * R]);
  double ah = __longlong_as_double((long long)rcd[2 * R + 1]);
#pragma unroll
  for (int i = 0; i < 12; i++) {
    const double c = (R == 0 && i == 0) ? C[0] + 8.0 : C[i];  // MDS_MATRIX_DIAG[0] = 8
    al = fma(xl[(i + R) % 12], c, al);
    ah = fma(xh[(i + R) % 12], c, ah);
  }
  // al = 2^52 + L, ah = 2^52 + H with L, H < 2^43: the mantissa bits ARE the integers
  const uint64_t L = (uint64_t)__double_as_longlong(al) & 0xFFFFFFFFFFFFFull;
  const uint64_t H = (uint64_t)__double_as_longlong(ah) & 0xFFFFFFFFFFFFFull;
  // value = L + H*2^32,  H = h1*2^32 + h0  =>  == L + h1*EPS + h0*2^32  (mod p)
  const uint32_t h0 = (uint32_t)H, h1 = (uint32_t)(H >> 32);
  const uint64_t t = L + (((uint64_t)h1 << 32) - h1);  // < 2^44, no overflow
  return gl::add_c(t, (uint64_t)h0 << 32);              // h0 << 32 < p: one fix-up is exact
}

// In-place permutation. Input lanes: any u64. Output lanes: any u64 (canonicalise before exporting).
//
// Loop iteration r = [S-boxes of lanes 1..11 if round r is full] ; MDS of round r (+ constants of
// round r+1) ; S-box of lane 0 for round r+1.  Lane 0's S-box — the only one in a partial round, a
// serial chain of four multiplications — is issued right behind MDS row 0 so that it overlaps the
// 264 DFMAs of rows 1..11 instead of stalling the warp in a basic block of its own.
__device__ __forceinline__ void permute(uint64_t (&s)[12]) {
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = gl::add_c(s[i], RC[i]);
  s[0] = sbox7(s[0]);
#pragma unroll 1
  for (int r = 0; r < ROUNDS; r++) {
    if (r < HALF_FULL || r >= HALF_FULL + PARTIAL) {  // warp-uniform
#pragma unroll
      for (int i = 1; i < 12; i++) s[i] = sbox7(s[i]);
    }
    const uint64_t* __restrict__ rcd = RC_F64 + 24 * r;
    double xl[12], xh[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
      xl[i] = u32_to_f64((uint32_t)s[i]);
      xh[i] = u32_to_f64((uint32_t)(s[i] >> 32));
    }
    const uint64_t row0 = mds_row_f64<0>(xl, xh, rcd);
    const uint64_t next0 = sbox7(row0);
    s[1] = mds_row_f64<1>(xl, xh, rcd);
    s[2] = mds_row_f64<2>(xl, xh, rcd);
    s[3] = mds_row_f64<3>(xl, xh, rcd);
    s[4] = mds_row_f64<4>(xl, xh, rcd);
    s[5] = mds_row_f64<5>(xl, xh, rcd);
    s[6] = mds_row_f64<6>(xl, xh, rcd);
    s[7] = mds_row_f64<7>(xl, xh, rcd);
    s[8] = mds_row_f64<8>(xl, xh, rcd);
    s[9] = mds_row_f64<9>(xl, xh, rcd);
    s[10] = mds_row_f64<10>(xl, xh, rcd);
    s[11] = mds_row_f64<11>(xl, xh, rcd);
    s[0] = (r == ROUNDS - 1) ? row0 : next0;  // no S-box after the last round
  }
}
#endif  // __CUDACC__

}  // namespace poseidon
