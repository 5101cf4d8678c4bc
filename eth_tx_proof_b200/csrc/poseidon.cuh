// Poseidon-12 over Goldilocks for sm_100a: one thread per permutation, state and MDS accumulators in
// registers, round constants in constant memory (K4/K5 building block).
//
// Replaces plonky2 0.2.2 `impl Poseidon for GoldilocksField` / `PoseidonPermutation`
// (plonky2/src/hash/poseidon.rs, poseidon_goldilocks.rs) and the sponge helpers of
// plonky2/src/hash/hashing.rs (hash_n_to_m_no_pad, compress) — crate pinned at
// /root/reference/Cargo.lock:3441, reached from /root/reference/ops/src/lib.rs:52.
//
// Round structure (identical output to upstream's naive and "fast" forms): 4 full + 22 partial + 4
// full rounds; each round = add 12 constants, S-box x^7 (all lanes / lane 0), circulant MDS
//   out[r] = sum_i in[(i+r)%12]*CIRC[i] + in[r]*DIAG[r].
// Here the NEXT round's constants are folded into the MDS accumulators, so a round is
//   S-box -> (MDS + RC_next) with one 96-bit reduction per lane.
// The MDS works on the 32-bit halves of each lane with IMAD.WIDE.U32 accumulation (coefficients
// < 2^6, so 12-term sums stay below 2^42) — integer pipe work, no tensor cores.
#pragma once
#include "gl.cuh"
#include "poseidon_constants.h"

namespace poseidon {

constexpr int WIDTH = 12, RATE = 8, HALF_FULL = 4, PARTIAL = 22, ROUNDS = 30;
#define ETP_MDS_CIRC {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20}

#if defined(__CUDACC__)
static __constant__ uint64_t RC[360] = ETP_POSEIDON_RC_TABLE;

__device__ __forceinline__ uint64_t sbox7(uint64_t x) {
  uint64_t x2 = gl::sqr(x);
  uint64_t x4 = gl::sqr(x2);
  uint64_t x3 = gl::mul(x, x2);
  return gl::mul(x3, x4);
}

// s <- MDS(s) + rc[0..12] (rc == nullptr: no constants, used after the last round)
__device__ __forceinline__ void mds_add_rc(uint64_t (&s)[12], const uint64_t* __restrict__ rc) {
  constexpr uint32_t C[12] = ETP_MDS_CIRC;
  uint32_t lo[12], hi[12];
#pragma unroll
  for (int i = 0; i < 12; i++) { lo[i] = (uint32_t)s[i]; hi[i] = (uint32_t)(s[i] >> 32); }
#pragma unroll
  for (int r = 0; r < 12; r++) {
    uint64_t k = rc ? rc[r] : 0;
    uint64_t al = (uint32_t)k, ah = k >> 32;
#pragma unroll
    for (int i = 0; i < 12; i++) {
      al += (uint64_t)lo[(i + r) % 12] * C[i];
      ah += (uint64_t)hi[(i + r) % 12] * C[i];
    }
    if (r == 0) { al += (uint64_t)lo[0] * 8u; ah += (uint64_t)hi[0] * 8u; }  // MDS_MATRIX_DIAG[0] = 8
    // value = al + ah*2^32, al, ah < 2^42.  ah = h1*2^32 + h0  =>  == al + h1*EPS + h0*2^32 (mod p)
    uint32_t h0 = (uint32_t)ah, h1 = (uint32_t)(ah >> 32);
    uint64_t t = al + (uint64_t)h1 * 0xFFFFFFFFu;  // < 2^43, no overflow
    s[r] = gl::add_c(t, (uint64_t)h0 << 32);       // h0 << 32 < p: one fix-up is exact
  }
}

// In-place permutation. Input lanes: any u64. Output lanes: any u64 (canonicalise before exporting).
__device__ __forceinline__ void permute(uint64_t (&s)[12]) {
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = gl::add_c(s[i], RC[i]);
  int r = 0;
#pragma unroll 1
  for (; r < HALF_FULL; r++) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]);
    mds_add_rc(s, RC + 12 * (r + 1));
  }
#pragma unroll 1
  for (; r < HALF_FULL + PARTIAL; r++) {
    s[0] = sbox7(s[0]);
    mds_add_rc(s, RC + 12 * (r + 1));
  }
#pragma unroll 1
  for (; r < ROUNDS - 1; r++) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]);
    mds_add_rc(s, RC + 12 * (r + 1));
  }
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]);
  mds_add_rc(s, nullptr);
}
#endif  // __CUDACC__

}  // namespace poseidon
