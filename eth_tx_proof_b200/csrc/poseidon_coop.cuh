// Poseidon-12 over Goldilocks, ONE permutation spread over 16 consecutive threads of a warp (lanes 0..11 of the
// group hold one state lane each, 12..15 only take part in the shuffles).  Same function as poseidon::permute
// (plonky2 0.2.2 plonky2/src/hash/poseidon.rs `Poseidon::poseidon`; crate pinned at /root/reference/Cargo.lock:3441,
// reached from /root/reference/ops/src/lib.rs:52), written in the naive round form:
//   add the round constants, x^7 on all lanes (full rounds) or lane 0 (partial rounds), circulant MDS.
//
// Why a second form: the one-thread-per-permutation kernel needs ~16 k dependent-ish instructions, 33 us for a
// single permutation on B200 — that is what every narrow Merkle level (the top of each tree, the small FRI layers)
// costs, however few nodes it has.  Here the S-boxes of a round run in parallel and the MDS row of a lane is 12
// multiply-adds on values all-gathered with warp shuffles (BASELINE.json north_star: "warp shuffles for the top
// levels and cap"): ~7 us per permutation.  Throughput per SM is several times lower than the register-resident
// form, so it is used only where a level cannot fill the machine (merkle.cuh: COOP_MAX_PARENTS).
#pragma once
#include "poseidon.cuh"

namespace poseidon {

#if defined(__CUDACC__)
constexpr int COOP_GROUP = 16;

// x: this thread's lane (any u64; lanes >= 12 pass anything).  l = thread index within the 16-group.  rc: the 360
// round constants in shared or global memory.  All 32 threads of the warp must call this together.
__device__ __forceinline__ uint64_t permute_coop(uint64_t x, int l, const uint64_t* __restrict__ rc) {
  constexpr uint32_t C[12] = ETP_MDS_CIRC;
  const int lc = l < 12 ? l : 0;
  ETP_ROLL
  for (int r = 0; r < ROUNDS; r++) {
    x = gl::add_c(x, rc[12 * r + lc]);
    const bool full = r < HALF_FULL || r >= HALF_FULL + PARTIAL;
    if (full ? (l < 12) : (l == 0)) x = sbox7(x);
    // out_l = sum_k CIRC[k] * x_{(l+k) % 12} (+ 8 x_0 on lane 0) on the two 32-bit halves (each sum < 2^42): lane l reads
    // lane (l + k) % 12 of its group, so the coefficient is a compile-time constant
    const uint32_t x0 = (uint32_t)x, x1 = (uint32_t)(x >> 32);
    uint64_t acc_lo = (uint64_t)(C[0] + (l == 0 ? 8u : 0u)) * x0, acc_hi = (uint64_t)(C[0] + (l == 0 ? 8u : 0u)) * x1;
    int src = lc;
#pragma unroll
    for (int k = 1; k < 12; k++) {
      src = (src == 11) ? 0 : src + 1;
      const uint32_t y0 = __shfl_sync(0xffffffffu, x0, src, COOP_GROUP);
      const uint32_t y1 = __shfl_sync(0xffffffffu, x1, src, COOP_GROUP);
      acc_lo += (uint64_t)C[k] * y0;
      acc_hi += (uint64_t)C[k] * y1;
    }
    // acc_lo + 2^32 acc_hi = lo + 2^64 hi with hi < 2^11
    const uint64_t sh = acc_hi << 32;
    const uint64_t lo = acc_lo + sh;
    const uint64_t hi = (acc_hi >> 32) + (lo < sh ? 1u : 0u);
    x = gl::reduce128(lo, hi);
  }
  return x;
}
#endif

}  // namespace poseidon
