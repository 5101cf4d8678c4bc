// Host-side Poseidon-12 permutation for the Fiat-Shamir transcript (product code, plain C++: compiled by g++, not nvcc).
//
// Same function as poseidon::permute on the device (plonky2 0.2.2 plonky2/src/hash/poseidon.rs; crate pinned at
// /root/reference/Cargo.lock:3441, reached from /root/reference/ops/src/lib.rs:52).  The transcript is strictly sequential
// (a duplex sponge), so it stays on a CPU core — but it is not always tiny: a table with C columns makes the challenger
// absorb 4 (C + aux + quotient) opening words, i.e. 1200 permutations for a keccak-shaped table (2400 columns), which at
// the ~10 us of a naive permutation was 12 ms of a 26 ms proof.  Hence: lazy representatives with branch-free fix-ups
// (the fix-up conditions are ~50/50, branches mispredict), the MDS layer on the 32-bit halves of the lanes with 64-bit
// accumulators, and an AVX2 version of that layer chosen at run time.  ~2 us per permutation.
// This is NOT the oracle: nothing here includes oracle/.
#include <cstdint>
#include <cstring>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "poseidon_constants.h"

namespace {
constexpr uint64_t EPS = 0xFFFFFFFFULL, P = 0xFFFFFFFF00000001ULL;
const uint64_t RC[360] = ETP_POSEIDON_RC_TABLE;
const uint32_t CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};

inline uint64_t mul(uint64_t a, uint64_t b) {
  const unsigned __int128 x = (unsigned __int128)a * b;
  const uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64), hh = hi >> 32, hl = hi & EPS;
  uint64_t t0 = lo - hh;
  t0 -= (0 - (uint64_t)(lo < hh)) & EPS;
  const uint64_t t1 = hl * EPS;
  uint64_t t2 = t0 + t1;
  t2 += (0 - (uint64_t)(t2 < t1)) & EPS;
  return t2;
}
inline uint64_t add_canonical(uint64_t a, uint64_t b /* < p */) {
  uint64_t s = a + b;
  s += (0 - (uint64_t)(s < a)) & EPS;
  return s;
}
inline uint64_t sbox7(uint64_t x) {
  const uint64_t x2 = mul(x, x), x4 = mul(x2, x2), x3 = mul(x, x2);
  return mul(x3, x4);
}
// al + 2^32 ah (each < 2^42) -> field element
inline uint64_t fold(uint64_t al, uint64_t ah) {
  const uint64_t sh = ah << 32, low = al + sh, top = (ah >> 32) + (low < sh);
  const uint64_t t1 = top * EPS;
  uint64_t t2 = low + t1;
  t2 += (0 - (uint64_t)(t2 < t1)) & EPS;
  return t2;
}
inline void split(const uint64_t s[12], uint64_t lo[24], uint64_t hi[24]) {
  for (int i = 0; i < 12; i++) {
    lo[i] = lo[i + 12] = (uint32_t)s[i];
    hi[i] = hi[i + 12] = s[i] >> 32;
  }
}
void mds_scalar(uint64_t s[12]) {
  uint64_t lo[24], hi[24], al[12], ah[12];
  split(s, lo, hi);
  for (int r = 0; r < 12; r++) {
    uint64_t a = 0, b = 0;
    for (int i = 0; i < 12; i++) {
      a += lo[i + r] * CIRC[i];
      b += hi[i + r] * CIRC[i];
    }
    al[r] = a;
    ah[r] = b;
  }
  al[0] += 8 * lo[0];
  ah[0] += 8 * hi[0];
  for (int r = 0; r < 12; r++) s[r] = fold(al[r], ah[r]);
}
#if defined(__x86_64__)
// out = sum_j x_j * column_j of the circulant: twelve broadcasts against constant column vectors (no rotated reloads of
// freshly stored data, which stall on store forwarding)
struct Columns {
  alignas(32) uint64_t c[12][12];
  Columns() {
    for (int j = 0; j < 12; j++)
      for (int r = 0; r < 12; r++) c[j][r] = CIRC[(j - r + 12) % 12];
  }
};
__attribute__((target("avx2"))) void mds_avx2(uint64_t s[12]) {
  static const Columns cols;
  uint64_t al[12], ah[12];
  __m256i l0 = _mm256_setzero_si256(), l1 = l0, l2 = l0, h0 = l0, h1 = l0, h2 = l0;
  for (int j = 0; j < 12; j++) {
    const __m256i xl = _mm256_set1_epi64x((long long)(uint32_t)s[j]), xh = _mm256_set1_epi64x((long long)(s[j] >> 32));
    const __m256i c0 = _mm256_load_si256((const __m256i*)&cols.c[j][0]), c1 = _mm256_load_si256((const __m256i*)&cols.c[j][4]),
                  c2 = _mm256_load_si256((const __m256i*)&cols.c[j][8]);
    l0 = _mm256_add_epi64(l0, _mm256_mul_epu32(xl, c0));
    l1 = _mm256_add_epi64(l1, _mm256_mul_epu32(xl, c1));
    l2 = _mm256_add_epi64(l2, _mm256_mul_epu32(xl, c2));
    h0 = _mm256_add_epi64(h0, _mm256_mul_epu32(xh, c0));
    h1 = _mm256_add_epi64(h1, _mm256_mul_epu32(xh, c1));
    h2 = _mm256_add_epi64(h2, _mm256_mul_epu32(xh, c2));
  }
  _mm256_storeu_si256((__m256i*)al, l0); _mm256_storeu_si256((__m256i*)(al + 4), l1); _mm256_storeu_si256((__m256i*)(al + 8), l2);
  _mm256_storeu_si256((__m256i*)ah, h0); _mm256_storeu_si256((__m256i*)(ah + 4), h1); _mm256_storeu_si256((__m256i*)(ah + 8), h2);
  al[0] += 8 * (uint64_t)(uint32_t)s[0];
  ah[0] += 8 * (s[0] >> 32);
  for (int r = 0; r < 12; r++) s[r] = fold(al[r], ah[r]);
}
#endif
template <void (*MDS)(uint64_t*)>
void permute(uint64_t s[12]) {
  const uint64_t* rc = RC;
  for (int r = 0; r < 30; r++, rc += 12) {
    for (int i = 0; i < 12; i++) s[i] = add_canonical(s[i], rc[i]);
    if (r < 4 || r >= 26) {
      for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]);
    } else {
      s[0] = sbox7(s[0]);
    }
    MDS(s);
  }
  for (int i = 0; i < 12; i++) s[i] = s[i] >= P ? s[i] - P : s[i];
}
}  // namespace

// in place; input lanes: any u64; output lanes canonical
extern "C" void etp_host_poseidon_permute(uint64_t s[12]) {
#if defined(__x86_64__)
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (avx2) return permute<mds_avx2>(s);
#endif
  permute<mds_scalar>(s);
}

// Witness of one PoseidonGate row (plonky2/src/gates/poseidon.rs PoseidonGenerator::run_once; wire layout of
// eth_tx_proof_b200/circuit.py PoseidonGate): inputs 0..11, outputs 12..23, swap 24, deltas 25..28, the S-box inputs of full
// rounds 1..3 at 29 + 12 (r - 1) + i, of the 22 partial rounds at 65 + r, of the last four full rounds at 87 + 12 r + i.
// The recursion layers' circuits are mostly PoseidonGate rows (Merkle paths, sponges, the in-circuit challenger), so their
// witness generation is this function called a few ten thousand times.  All values canonical.
extern "C" void etp_host_poseidon_gate_wires(const uint64_t inputs[12], int swap, uint64_t wires_out[135]) {
  const auto canon = [](uint64_t x) { return x >= P ? x - P : x; };
  uint64_t in[12], s[12];
  memset(wires_out, 0, 135 * sizeof(uint64_t));
  for (int i = 0; i < 12; i++) wires_out[i] = s[i] = in[i] = canon(inputs[i]);
  wires_out[24] = swap ? 1 : 0;
  for (int i = 0; i < 4; i++) {
    const uint64_t delta = swap ? canon(add_canonical(in[i + 4], in[i] ? P - in[i] : 0)) : 0;
    wires_out[25 + i] = delta;
    s[i] = canon(add_canonical(in[i], delta));
    s[i + 4] = canon(add_canonical(in[i + 4], delta ? P - delta : 0));
  }
#if defined(__x86_64__)
  static const bool avx2 = __builtin_cpu_supports("avx2");
#else
  const bool avx2 = false;
#endif
  const uint64_t* rc = RC;
  for (int r = 0; r < 30; r++, rc += 12) {
    for (int i = 0; i < 12; i++) s[i] = canon(add_canonical(s[i], rc[i]));
    if (r < 4 || r >= 26) {
      if (r >= 1 && r < 4)
        for (int i = 0; i < 12; i++) wires_out[29 + 12 * (r - 1) + i] = s[i];
      if (r >= 26)
        for (int i = 0; i < 12; i++) wires_out[87 + 12 * (r - 26) + i] = s[i];
      for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]);
    } else {
      wires_out[65 + (r - 4)] = s[0];
      s[0] = sbox7(s[0]);
    }
#if defined(__x86_64__)
    if (avx2) mds_avx2(s); else
#endif
    mds_scalar(s);
  }
  for (int i = 0; i < 12; i++) wires_out[12 + i] = canon(s[i]);
}

// The parameters of the permutation, for callers that build circuits over it (plonky2's PoseidonGate evaluates the
// permutation symbolically): ALL_ROUND_CONSTANTS (30 x 12), MDS_MATRIX_CIRC, MDS_MATRIX_DIAG.
extern "C" void etp_poseidon_constants(uint64_t round_constants_out[360], uint64_t mds_circ_out[12], uint64_t mds_diag_out[12]) {
  static const uint64_t circ[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
  for (int i = 0; i < 360 && round_constants_out; i++) round_constants_out[i] = RC[i];
  for (int i = 0; i < 12 && mds_circ_out; i++) mds_circ_out[i] = circ[i];
  for (int i = 0; i < 12 && mds_diag_out; i++) mds_diag_out[i] = i == 0 ? 8 : 0;
}
