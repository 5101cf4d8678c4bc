// Column-batched Goldilocks NTT for sm_100a (K2/K3 in SURVEY.md 2.4).
//
// Replaces plonky2_field 0.2.2 `fft_with_options` / `ifft_with_options` and
// `PolynomialCoeffs::{lde, coset_fft_with_options}`, `PolynomialValues::{ifft, coset_ifft}`
// (field/src/fft.rs, field/src/polynomial/mod.rs; crate pinned at /root/reference/Cargo.lock:3466,
// reached from /root/reference/ops/src/lib.rs:52 through PolynomialBatch::from_values/from_coeffs).
//
// Definition (same as upstream): out[k] = sum_j in[j] * w^(j k), w = primitive_root_of_unity(log_n);
// the inverse uses w^-1 and scales by 1/n; coset variants pre-multiply in[j] by shift^j (forward) or
// post-multiply out[k] by shift^-k (inverse).  Results are field elements, hence independent of the
// butterfly schedule used here.
//
// Schedule: the index bits are cut into digits of B <= 8 bits, most significant first.  One kernel
// launch ("pass") transforms one digit for every (prefix, low) pair:
//   * a CTA of 256 threads owns a tile of 4096 elements = 2^B digit values x U "units"
//     (strided passes: U neighbouring low indices -> 128-byte runs; last pass: U contiguous chunks);
//   * every thread keeps 16 elements in registers: stage 1 = the top min(B,4) DIF layers straight
//     from global memory, one exchange through shared memory, stage 2 = the remaining layers, which
//     are a plain 2^(B-4)-point DFT with compile-time constant twiddles (powers of w16; the trivial
//     ones cost nothing), then the inter-digit twiddle w_{2^(s+B)}^(low*k) from a precomputed table
//     (L2 resident: CTAs of the same tile across columns are adjacent in the grid) and the store;
//   * every pass leaves its digit bit-reversed in place, so after the last pass position p holds
//     out[bitrev(p)] — plonky2's leaf order (reverse_index_bits_in_place) — for free.  The natural
//     variant of the last pass reads U chunks with bit-reversed prefixes so its store is coalesced.
// Arithmetic is integer-pipe bound (profiles/): ~13 ALU + 8 FMA-pipe instructions per multiplication.
#pragma once
#include <cuda_runtime.h>

#include "gl.cuh"
#include "powtable.cuh"

namespace ntt {

// butterfly add / sub: -DETP_NTT_RARE_FIX=0 selects the fully inline double fix-up
#ifndef ETP_NTT_RARE_FIX
#define ETP_NTT_RARE_FIX 1
#endif
#if ETP_NTT_RARE_FIX
__device__ __forceinline__ uint64_t badd(uint64_t a, uint64_t b) { return gl::add_r(a, b); }
__device__ __forceinline__ uint64_t bsub(uint64_t a, uint64_t b) { return gl::sub_r(a, b); }
#else
__device__ __forceinline__ uint64_t badd(uint64_t a, uint64_t b) { return gl::add(a, b); }
__device__ __forceinline__ uint64_t bsub(uint64_t a, uint64_t b) { return gl::sub(a, b); }
#endif

constexpr int THREADS = 256;
constexpr int TILE_LOG = 12;       // 4096 elements per CTA, 16 per thread
constexpr int MAX_DIGIT_BITS = 8;
#ifndef ETP_NTT_MIN_BLOCKS
#define ETP_NTT_MIN_BLOCKS 4
#endif

// powers of w16 = primitive_root_of_unity(4) and of its inverse (all of them are powers of two mod p)
static __device__ __constant__ const uint64_t W16[8] = {0x0000000000000001ULL, 0xefffffff00000001ULL, 0xfffffffeff000001ULL, 0x000ffffffff00000ULL,
                                                 0x0001000000000000ULL, 0x0000000000001000ULL, 0xfffffeff00000101ULL, 0xffffffef00000001ULL};
static __device__ __constant__ const uint64_t W16I[8] = {0x0000000000000001ULL, 0x0000001000000000ULL, 0x000000ffffffff00ULL, 0xfffffffefffff001ULL,
                                                  0xfffeffff00000001ULL, 0xffefffff00100001ULL, 0x0000000001000000ULL, 0x1000000000000000ULL};

struct PassParams {
  const uint64_t* in;   // column 0
  uint64_t* out;
  size_t in_col_stride, out_col_stride;  // elements between columns
  uint32_t n_cols;      // grid = tiles x columns, column fastest (blockIdx.x = tile * n_cols + column)
  int log_n;            // total transform size
  int s;                // bits below this digit
  uint32_t n_in;        // inputs at positions >= n_in are zero (LDE zero padding); first pass only
  int natural_out;      // last pass only: store out[bitrev(position)] (natural order) instead of in place
  int inverse;          // use w^-1
  PowTable tw;          // powers of w_{2^log_n} (forward) — inverse indexes it with n - e
  const uint64_t* tw_full;  // strided passes: twiddle per (digit slot << s | low), 2^(s+B) entries, or nullptr
  int in_scale;         // multiply input j by shift^j (coset_fft), first pass
  PowTable in_pow;
  const uint64_t* in_full;  // shift^j for j < n_in, or nullptr
  int out_scale;        // 0: none, 1: multiply output k by out_pow.get(k) (coset_ifft incl. 1/n), 2: by out_const
  PowTable out_pow;
  uint64_t out_const;   // 1/n for plain ifft
};

__device__ __forceinline__ uint64_t tw_get(const PassParams& p, uint32_t e /* < 2^log_n */) {
  if (p.inverse && e) e = (1u << p.log_n) - e;
  return p.tw.get(e);
}
// out-of-line fallbacks for transforms whose full tables would not fit (keeps the hot code small:
// the unrolled kernels otherwise overflow the instruction cache)
static __device__ __noinline__ uint64_t slow_pow(const uint64_t* lo, const uint64_t* hi, int lo_bits, uint32_t mask, uint32_t e) {
  return gl::mul(__ldg(hi + (e >> lo_bits)), __ldg(lo + (e & mask)));
}
static __device__ __noinline__ uint64_t slow_twiddle(const uint64_t* lo, const uint64_t* hi, int lo_bits, uint32_t mask, int inverse, int log_n,
                                              uint32_t e) {
  if (inverse && e) e = (1u << log_n) - e;
  return gl::mul(__ldg(hi + (e >> lo_bits)), __ldg(lo + (e & mask)));
}

// 2^Q-point DIF with compile-time constant twiddles on registers v[0..2^Q), stride 1: output slot j holds
// frequency bitrev_Q(j).  Twiddle of layer q, pair (j, j+half): w_{2^Q}^((j mod half) << q).
template <int Q>
__device__ __forceinline__ void dft_const(uint64_t* v, bool inverse) {
#pragma unroll
  for (int q = 0; q < Q; q++) {
    const int half = 1 << (Q - 1 - q);
#pragma unroll
    for (int j = 0; j < (1 << Q); j++) {
      if ((j & half) == 0) {
        const int e = ((j & (half - 1)) << q) << (4 - Q);  // exponent in units of w16
        const uint64_t a = v[j], c = v[j + half];
        v[j] = badd(a, c);
        const uint64_t d = bsub(a, c);
        v[j + half] = (e == 0) ? d : gl::mul(d, inverse ? W16I[e] : W16[e]);
      }
    }
  }
}

// Q top DIF layers of a 2^B-point transform on v[j] = x[(j << (B-Q)) | d_lo]: general twiddles
// w_R^(((j mod half) << (B-Q) | d_lo) << q) read from the shared table tw_r (R/2 entries).
template <int B, int Q>
__device__ __forceinline__ void dif_top(uint64_t* v, const uint64_t* tw_r, int d_lo) {
#pragma unroll
  for (int q = 0; q < Q; q++) {
    const int half = 1 << (Q - 1 - q);
#pragma unroll
    for (int j = 0; j < (1 << Q); j++) {
      if ((j & half) == 0) {
        const int dm = ((j & (half - 1)) << (B - Q)) | d_lo;
        const uint64_t w = tw_r[dm << q];
        const uint64_t a = v[j], c = v[j + half];
        v[j] = badd(a, c);
        v[j + half] = gl::mul(bsub(a, c), w);
      }
    }
  }
}

template <int B>
struct Geo {
  static constexpr int Q1 = B < 4 ? B : 4;      // stage 1 layers (top)
  static constexpr int Q2 = B - Q1;             // stage 2 layers (bottom, constant twiddles)
  static constexpr int U_LOG = TILE_LOG - B;    // units per CTA
  static constexpr int U = 1 << U_LOG;
  static constexpr int G1 = 16 >> Q1;           // groups per thread in stage 1
  static constexpr int G2 = 16 >> Q2;           // groups per thread in stage 2
  static constexpr int PITCH = (1 << B) + (B >= 4 ? (1 << (B - 4)) : 0) + 1;  // last pass: padded row pitch
  static constexpr int SMEM_WORDS_STRIDED = B > 4 ? (1 << TILE_LOG) + (1 << B) / 2 : 0;  // B <= 4: registers only
  static constexpr int SMEM_WORDS_LAST = B > 4 ? U * PITCH + (1 << B) / 2 : 0;
};

// ---- strided pass (s >= U_LOG): grid = (columns, tiles) --------------------------------------------
// MODE: compile-time copy of the run-time switches, so that the hot instantiations carry no dead code (the
// straight-line kernels are executed once per CTA; ncu showed instruction-fetch stalls, profiles/).
//   strided: -1 read p.in_scale | 0 no input scaling | 1 input scaling from the full table
//   last   : -1 read the flags  | 0 in place, no scaling (LDE) | 1 natural order, times 1/n (plain iFFT)
template <int B, int MODE>
static __global__ void __launch_bounds__(THREADS, ETP_NTT_MIN_BLOCKS) pass_strided(PassParams p, int units_log) {
  using G = Geo<B>;
  const bool in_scale = MODE < 0 ? (p.in_scale != 0) : (MODE == 1);
  const int U = 1 << units_log;  // <= G::U; smaller only when the transform is smaller than a tile
  extern __shared__ uint64_t smem[];
  uint64_t* tw_r = smem;              // R/2 inner twiddles
  uint64_t* sm = smem + ((1 << B) >> 1);
  const int t = threadIdx.x, s = p.s;
  const uint32_t col = blockIdx.x % p.n_cols, tile = blockIdx.x / p.n_cols;
  const uint64_t* in = p.in + (size_t)col * p.in_col_stride;
  uint64_t* out = p.out + (size_t)col * p.out_col_stride;
  const uint32_t tiles_per_prefix = 1u << (s - units_log);
  const uint32_t prefix = tile / tiles_per_prefix;
  const uint32_t low0 = (tile % tiles_per_prefix) << units_log;
  const size_t base = ((size_t)prefix << (s + B)) | low0;
  if (B > 4)
    for (int j = t; j < ((1 << B) >> 1); j += THREADS) tw_r[j] = tw_get(p, (uint32_t)j << (p.log_n - B));
  uint64_t v[16];
  // stage 1: G1 groups of 2^Q1 elements, group g -> (u = g % U, d_lo = g / U), element j at digit (j << (B-Q1)) | d_lo
#pragma unroll
  for (int gi = 0; gi < G::G1; gi++) {
    const int g = t + gi * THREADS, u = g & (G::U - 1), d_lo = g >> G::U_LOG;
#pragma unroll
    for (int j = 0; j < (1 << G::Q1); j++) {
      const int d = (j << (B - G::Q1)) | d_lo;
      const size_t pos = base | ((size_t)d << s) | u;
      uint64_t x = 0;
      if (u < U && pos < p.n_in) {
        x = in[pos];
        if (in_scale) x = gl::mul(x, (MODE == 1 || p.in_full) ? __ldg(p.in_full + pos) : slow_pow(p.in_pow.lo, p.in_pow.hi, p.in_pow.lo_bits, p.in_pow.mask, (uint32_t)pos));
      }
      v[gi * (1 << G::Q1) + j] = x;
    }
  }
  if (B > 4) {
    __syncthreads();  // tw_r ready
    {
      const int u = t & (G::U - 1), d_lo = t >> G::U_LOG;
      dif_top<B, 4>(v, tw_r, d_lo);
#pragma unroll
      for (int j = 0; j < 16; j++) sm[(((j << (B - 4)) | d_lo) << G::U_LOG) | u] = v[j];
    }
    __syncthreads();
    // stage 2: group g -> (u = g % U, d_hi = g / U), element j at digit (d_hi << Q2) | j
#pragma unroll
    for (int gi = 0; gi < G::G2; gi++) {
      const int g = t + gi * THREADS, u = g & (G::U - 1), d_hi = g >> G::U_LOG;
#pragma unroll
      for (int j = 0; j < (1 << G::Q2); j++) v[gi * (1 << G::Q2) + j] = sm[(((d_hi << G::Q2) | j) << G::U_LOG) | u];
      dft_const<G::Q2>(v + gi * (1 << G::Q2), p.inverse);
    }
  } else {
#pragma unroll
    for (int gi = 0; gi < G::G1; gi++) dft_const<G::Q1>(v + gi * (1 << G::Q1), p.inverse);
  }
  // inter-digit twiddle + store in place (digit slot d holds frequency bitrev_B(d))
  constexpr int QL = (B > 4) ? G::Q2 : G::Q1, GL = (B > 4) ? G::G2 : G::G1;
#pragma unroll
  for (int gi = 0; gi < GL; gi++) {
    const int g = t + gi * THREADS, u = g & (G::U - 1), dg = g >> G::U_LOG;
#pragma unroll
    for (int j = 0; j < (1 << QL); j++) {
      const int d = (B > 4) ? ((dg << G::Q2) | j) : ((j << (B - G::Q1)) | dg);
      const uint32_t low = low0 + u;
      if (u < U) {
        uint64_t w;
        if (MODE >= 0 || p.tw_full) w = __ldg(p.tw_full + (((size_t)d << s) | low));
        else w = slow_twiddle(p.tw.lo, p.tw.hi, p.tw.lo_bits, p.tw.mask, p.inverse, p.log_n, (uint32_t)(((uint64_t)low * gl::bitrev32(d, B)) << (p.log_n - s - B)));
        out[base | ((size_t)d << s) | u] = gl::mul(v[gi * (1 << QL) + j], w);
      }
    }
  }
}

// ---- last pass (s == 0): a CTA owns U chunks of 2^B contiguous elements; grid = (columns, tiles) ----
//   in place (bit-reversed order): chunks q0 .. q0+U-1 ; natural_out: chunks bitrev(q0 + u)
template <int B, int MODE>
static __global__ void __launch_bounds__(THREADS, ETP_NTT_MIN_BLOCKS) pass_last(PassParams p, int units_log) {
  using G = Geo<B>;
  const bool in_scale = MODE < 0 ? (p.in_scale != 0) : false;
  const bool natural_out = MODE < 0 ? (p.natural_out != 0) : (MODE == 1);
  const int out_scale = MODE < 0 ? p.out_scale : (MODE == 1 ? 2 : 0);
  extern __shared__ uint64_t smem[];
  uint64_t* tw_r = smem;
  uint64_t* sm = smem + ((1 << B) >> 1);
  const int t = threadIdx.x;
  const int pb = p.log_n - B;
  const uint32_t col = blockIdx.x % p.n_cols, tile = blockIdx.x / p.n_cols;
  const uint64_t* in = p.in + (size_t)col * p.in_col_stride;
  uint64_t* out = p.out + (size_t)col * p.out_col_stride;
  const int U = 1 << units_log;  // <= G::U (small transforms have fewer chunks than a full tile)
  const uint32_t q0 = tile << units_log;
  if (B > 4)
    for (int j = t; j < ((1 << B) >> 1); j += THREADS) tw_r[j] = tw_get(p, (uint32_t)j << (p.log_n - B));
  uint64_t v[16];
  // stage 1: group g -> (d_lo = g % 2^(B-Q1), u = g / 2^(B-Q1)); loads run along d_lo
#pragma unroll
  for (int gi = 0; gi < G::G1; gi++) {
    const int g = t + gi * THREADS, d_lo = g & ((1 << (B - G::Q1)) - 1), u = g >> (B - G::Q1);
    const uint32_t prefix = natural_out ? gl::bitrev32(q0 + u, pb) : (q0 + u);
#pragma unroll
    for (int j = 0; j < (1 << G::Q1); j++) {
      const int d = (j << (B - G::Q1)) | d_lo;
      const size_t pos = ((size_t)prefix << B) | d;
      uint64_t x = 0;
      if (u < U && pos < p.n_in) {
        x = in[pos];
        if (in_scale) x = gl::mul(x, p.in_full ? __ldg(p.in_full + pos) : slow_pow(p.in_pow.lo, p.in_pow.hi, p.in_pow.lo_bits, p.in_pow.mask, (uint32_t)pos));
      }
      v[gi * (1 << G::Q1) + j] = x;
    }
  }
  if (B > 4) {
    __syncthreads();
    {
      const int d_lo = t & ((1 << (B - 4)) - 1), u = t >> (B - 4);
      dif_top<B, 4>(v, tw_r, d_lo);
#pragma unroll
      for (int j = 0; j < 16; j++) { const int d = (j << (B - 4)) | d_lo; sm[u * G::PITCH + d + (d >> 4)] = v[j]; }
    }
    __syncthreads();
    // stage 2: group g -> (u = g % U_full, d_hi = g / U_full): reads and natural-order stores run along u
#pragma unroll
    for (int gi = 0; gi < G::G2; gi++) {
      const int g = t + gi * THREADS, u = g & (G::U - 1), d_hi = g >> G::U_LOG;
#pragma unroll
      for (int j = 0; j < (1 << G::Q2); j++) { const int d = (d_hi << G::Q2) | j; v[gi * (1 << G::Q2) + j] = sm[u * G::PITCH + d + (d >> 4)]; }
      dft_const<G::Q2>(v + gi * (1 << G::Q2), p.inverse);
    }
  } else {
#pragma unroll
    for (int gi = 0; gi < G::G1; gi++) dft_const<G::Q1>(v + gi * (1 << G::Q1), p.inverse);
  }
  constexpr int QL = (B > 4) ? G::Q2 : G::Q1, GL = (B > 4) ? G::G2 : G::G1;
  if (natural_out) {
    // out[(k << pb) | (q0 + u)], k = bitrev_B(d)
#pragma unroll
    for (int gi = 0; gi < GL; gi++) {
      const int g = t + gi * THREADS;
      int u, dg;
      if (B > 4) { u = g & (G::U - 1); dg = g >> G::U_LOG; }
      else { dg = g & ((1 << (B - G::Q1)) - 1); u = g >> (B - G::Q1); }
#pragma unroll
      for (int j = 0; j < (1 << QL); j++) {
        const int d = (B > 4) ? ((dg << G::Q2) | j) : ((j << (B - G::Q1)) | dg);
        if (u < U) {
          const uint32_t idx = (gl::bitrev32(d, B) << pb) | (q0 + u);
          uint64_t x = v[gi * (1 << QL) + j];
          if (out_scale == 2) x = gl::mul(x, p.out_const);
          else if (out_scale == 1) x = gl::mul(x, slow_pow(p.out_pow.lo, p.out_pow.hi, p.out_pow.lo_bits, p.out_pow.mask, idx));
          out[idx] = gl::canon(x);
        }
      }
    }
  } else {
    // in place: go back through shared memory so that the store runs along d
    if (B > 4) {
      __syncthreads();
#pragma unroll
      for (int gi = 0; gi < G::G2; gi++) {
        const int g = t + gi * THREADS, u = g & (G::U - 1), d_hi = g >> G::U_LOG;
#pragma unroll
        for (int j = 0; j < (1 << G::Q2); j++) { const int d = (d_hi << G::Q2) | j; sm[u * G::PITCH + d + (d >> 4)] = v[gi * (1 << G::Q2) + j]; }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 16; k++) {
        const int e = t + k * THREADS, d = e & ((1 << B) - 1), u = e >> B;
        if (u < U) {
          const size_t pos = ((size_t)(q0 + u) << B) | d;
          uint64_t x = sm[u * G::PITCH + d + (d >> 4)];
          if (out_scale == 2) x = gl::mul(x, p.out_const);
          else if (out_scale == 1) x = gl::mul(x, slow_pow(p.out_pow.lo, p.out_pow.hi, p.out_pow.lo_bits, p.out_pow.mask, gl::bitrev32((uint32_t)pos, p.log_n)));
          out[pos] = gl::canon(x);
        }
      }
    } else {
#pragma unroll
      for (int gi = 0; gi < G::G1; gi++) {
        const int g = t + gi * THREADS, d_lo = g & ((1 << (B - G::Q1)) - 1), u = g >> (B - G::Q1);
#pragma unroll
        for (int j = 0; j < (1 << G::Q1); j++) {
          const int d = (j << (B - G::Q1)) | d_lo;
          if (u < U) {
            const size_t pos = ((size_t)(q0 + u) << B) | d;
            uint64_t x = v[gi * (1 << G::Q1) + j];
            if (out_scale == 2) x = gl::mul(x, p.out_const);
            else if (out_scale == 1) x = gl::mul(x, slow_pow(p.out_pow.lo, p.out_pow.hi, p.out_pow.lo_bits, p.out_pow.mask, gl::bitrev32((uint32_t)pos, p.log_n)));
            out[pos] = gl::canon(x);
          }
        }
      }
    }
  }
}

// twiddle table of a strided pass: tab[(d << s) | low] = w^(+-(low * bitrev_B(d)) << (log_n - s - B))
static __global__ void fill_pass_twiddles(PassParams p, int B, uint64_t* tab) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= ((size_t)1 << (p.s + B))) return;
  const uint32_t d = (uint32_t)(i >> p.s), low = (uint32_t)(i & (((size_t)1 << p.s) - 1));
  tab[i] = gl::canon(tw_get(p, (uint32_t)(((uint64_t)low * gl::bitrev32(d, B)) << (p.log_n - p.s - B))));
}
// shift^j table
static __global__ void fill_pow_full(PowTable t, size_t n, uint64_t* tab) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) tab[i] = gl::canon(t.get((uint32_t)i));
}

}  // namespace ntt
