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
// Schedule: the index bits are cut into digits of <= MAX_DIGIT_BITS bits, most significant first.
// One kernel launch ("pass") transforms one digit for every (prefix, low) pair:
//   * tile = 2^b digit values x U neighbouring "units" staged in shared memory,
//   * radix-16 register stages (4 DIF layers per shared-memory round trip),
//   * inter-digit twiddle w_{2^(s+b)}^(low*k) applied on the way out (two-level power table),
//   * loads and stores are 64/128-byte coalesced runs (units are neighbouring low indices in the
//     strided passes and whole contiguous chunks in the last pass).
// Every pass leaves its digit bit-reversed in place, so after the last pass position p holds
// out[bitrev(p)] — plonky2's leaf order (reverse_index_bits_in_place) — for free.  The "natural"
// variant of the last pass scatters T chunks at once so that the natural-order store is coalesced.
#pragma once
#include <cuda_runtime.h>

#include "gl.cuh"

namespace ntt {

constexpr int MAX_DIGIT_BITS = 11;   // strided passes: 2^11 x 8 units x 8 B = 128 KiB of shared memory
constexpr int STRIDED_UNITS_LOG = 3; // 8 neighbouring low indices = 64-byte runs
constexpr int LAST_UNITS_LOG = 3;    // last pass: 8 chunks per CTA
constexpr int THREADS = 512;

// base^e for e < 2^bits via two tables: hi[e >> lo_bits] * lo[e & mask].  `hi` may carry a scale.
struct PowTable {
  const uint64_t* lo;
  const uint64_t* hi;
  int lo_bits;
  uint32_t mask;
  __device__ __forceinline__ uint64_t get(uint32_t e) const {
    return gl::mul(__ldg(hi + (e >> lo_bits)), __ldg(lo + (e & mask)));
  }
};

struct PassParams {
  const uint64_t* in;   // column 0
  uint64_t* out;
  size_t in_col_stride, out_col_stride;  // elements between columns
  int log_n;            // total transform size
  int s;                // bits below this digit
  int b;                // digit bits
  uint32_t n_in;        // inputs at positions >= n_in are zero (LDE zero padding); first pass only
  int natural_out;      // last pass only: store out[bitrev(position)] (natural order) instead of in place
  int inverse;          // use w^-1
  PowTable tw;          // powers of w_{2^log_n} (forward) — inverse indexes it with n - e
  int in_scale;         // multiply input j by in_pow^j (coset_fft), first pass
  PowTable in_pow;
  int out_scale;        // 0: none, 1: multiply output k by out_pow.get(k) (coset_ifft incl. 1/n), 2: by out_const
  PowTable out_pow;
  uint64_t out_const;   // 1/n for plain ifft
};

// ------------------------------------------------------------------------------------------------
// shared-memory tile indexing.  Element (unit u, digit d).
//   strided passes: [d][u], u fastest (loads/stores run along u)
//   last pass:      [u][d], d fastest, padded (+1 per 16, row pitch == 1 mod 16) against bank conflicts
struct LayoutStrided {
  int ulog;
  __device__ __forceinline__ int operator()(int u, int d) const { return (d << ulog) | u; }
};
struct LayoutLast {
  int pitch;
  __device__ __forceinline__ int operator()(int u, int d) const { return u * pitch + d + (d >> 4); }
};
__host__ __device__ inline int last_pitch(int b) { return (1 << b) + (b >= 4 ? (1 << (b - 4)) : 0) + 1; }

// One register stage: Q DIF layers starting at layer `l0` of a size-2^b transform, for all units.
// DIF layer l pairs digits d and d + h, h = 2^(b-1-l), twiddle w_R^((d mod h) << l), R = 2^b.
template <int Q, class Layout>
__device__ __forceinline__ void dif_stage(uint64_t* sm, const Layout& at, const uint64_t* tw_r /* w_R^j, j < R/2 */,
                                          int b, int l0, int n_units_log, int tid, int nthreads, bool units_fastest) {
  const int fpos = b - l0 - Q;                 // bit position of the Q-bit field inside the digit
  const int groups_log = b - Q + n_units_log;  // (digit without field) x units
  for (int g = tid; g < (1 << groups_log); g += nthreads) {
    int u, gd;
    if (units_fastest) { u = g & ((1 << n_units_log) - 1); gd = g >> n_units_log; }
    else { gd = g & ((1 << (b - Q)) - 1); u = g >> (b - Q); }
    const int d_lo = gd & ((1 << fpos) - 1), d_hi = gd >> fpos;
    const int d0 = (d_hi << (fpos + Q)) | d_lo;
    uint64_t v[1 << Q];
#pragma unroll
    for (int j = 0; j < (1 << Q); j++) v[j] = sm[at(u, d0 | (j << fpos))];
#pragma unroll
    for (int q = 0; q < Q; q++) {
      // layer l = l0 + q: half = 2^(Q-1-q) in field units; pairs (j, j + half)
      const int half = 1 << (Q - 1 - q);
      const int l = l0 + q;
#pragma unroll
      for (int j = 0; j < (1 << Q); j++) {
        if ((j & half) == 0) {
          // digit of element j: d0 | j << fpos ; (d mod h) with h = 2^(b-1-l) = half << fpos
          const int dm = ((j & (half - 1)) << fpos) | d_lo;
          const uint64_t w = tw_r[dm << l];
          const uint64_t a = v[j], c = v[j + half];
          v[j] = gl::add(a, c);
          v[j + half] = gl::mul(gl::sub(a, c), w);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < (1 << Q); j++) sm[at(u, d0 | (j << fpos))] = v[j];
  }
}

template <class Layout>
__device__ __forceinline__ void dif_all(uint64_t* sm, const Layout& at, const uint64_t* tw_r, int b, int n_units_log,
                                        bool units_fastest) {
  const int tid = threadIdx.x, nt = blockDim.x;
  int l = 0;
  while (b - l >= 4) { dif_stage<4>(sm, at, tw_r, b, l, n_units_log, tid, nt, units_fastest); l += 4; __syncthreads(); }
  if (b - l == 3) { dif_stage<3>(sm, at, tw_r, b, l, n_units_log, tid, nt, units_fastest); __syncthreads(); }
  else if (b - l == 2) { dif_stage<2>(sm, at, tw_r, b, l, n_units_log, tid, nt, units_fastest); __syncthreads(); }
  else if (b - l == 1) { dif_stage<1>(sm, at, tw_r, b, l, n_units_log, tid, nt, units_fastest); __syncthreads(); }
}

__device__ __forceinline__ uint64_t tw_get(const PassParams& p, uint32_t e /* < 2^log_n */) {
  if (p.inverse && e) e = (1u << p.log_n) - e;
  return p.tw.get(e);
}

// fills tw_r[j] = w_R^(+-j), j < R/2 (R = 2^b) from the global table
__device__ __forceinline__ void fill_inner_twiddles(const PassParams& p, uint64_t* tw_r) {
  const int half = (1 << p.b) >> 1;
  for (int j = threadIdx.x; j < half; j += blockDim.x) tw_r[j] = tw_get(p, (uint32_t)j << (p.log_n - p.b));
}

// ---- strided pass: s >= units_log. grid.x = (2^log_n >> b) >> units_log tiles, grid.y = columns
static __global__ void __launch_bounds__(THREADS) pass_strided(PassParams p, int units_log) {
  extern __shared__ uint64_t smem[];
  const int b = p.b, s = p.s, R = 1 << b, U = 1 << units_log;
  uint64_t* tw_r = smem;            // R/2
  uint64_t* sm = smem + (R >> 1);   // R * U
  const LayoutStrided at{units_log};
  const uint64_t* in = p.in + (size_t)blockIdx.y * p.in_col_stride;
  uint64_t* out = p.out + (size_t)blockIdx.y * p.out_col_stride;
  // tile -> (prefix, low0)
  const uint32_t tiles_per_prefix = 1u << (s - units_log);
  const uint32_t prefix = blockIdx.x / tiles_per_prefix;
  const uint32_t low0 = (blockIdx.x % tiles_per_prefix) << units_log;
  const size_t base = ((size_t)prefix << (s + b)) | low0;
  fill_inner_twiddles(p, tw_r);
  for (int e = threadIdx.x; e < R * U; e += blockDim.x) {
    const int u = e & (U - 1), d = e >> units_log;
    const size_t pos = base | ((size_t)d << s) | u;
    uint64_t v = 0;
    if (pos < p.n_in) {
      v = in[pos];
      if (p.in_scale) v = gl::mul(v, p.in_pow.get((uint32_t)pos));
    }
    sm[at(u, d)] = v;
  }
  __syncthreads();
  dif_all(sm, at, tw_r, b, units_log, true);
  // slot d holds output digit k = bitrev_b(d); twiddle w_{2^(s+b)}^(low*k)
  for (int e = threadIdx.x; e < R * U; e += blockDim.x) {
    const int u = e & (U - 1), d = e >> units_log;
    const uint32_t k = gl::bitrev32(d, b);
    const uint32_t low = low0 + u;
    uint64_t v = sm[at(u, d)];
    const uint32_t ex = (uint32_t)(((uint64_t)low * k) << (p.log_n - s - b));  // < 2^log_n
    v = gl::mul(v, tw_get(p, ex));
    out[base | ((size_t)d << s) | u] = v;
  }
}

// ---- last pass: s == 0. A CTA owns U chunks of R contiguous elements.
//   in place (bit-reversed order):  chunks prefix0 .. prefix0+U-1 (one contiguous region)
//   natural_out: chunks bitrev(q0 + u) so that out[(k << (log_n-b)) | (q0+u)] runs along u
static __global__ void __launch_bounds__(THREADS) pass_last(PassParams p, int units_log) {
  extern __shared__ uint64_t smem[];
  const int b = p.b, R = 1 << b, U = 1 << units_log;
  const int pb = p.log_n - b;  // prefix bits
  uint64_t* tw_r = smem;
  uint64_t* sm = smem + (R >> 1);
  const LayoutLast at{last_pitch(b)};
  const uint64_t* in = p.in + (size_t)blockIdx.y * p.in_col_stride;
  uint64_t* out = p.out + (size_t)blockIdx.y * p.out_col_stride;
  const uint32_t q0 = blockIdx.x << units_log;
  fill_inner_twiddles(p, tw_r);
  for (int e = threadIdx.x; e < R * U; e += blockDim.x) {
    const int d = e & (R - 1), u = e >> b;
    const uint32_t prefix = p.natural_out ? gl::bitrev32(q0 + u, pb) : (q0 + u);
    const size_t pos = ((size_t)prefix << b) | d;
    uint64_t v = 0;
    if (pos < p.n_in) {
      v = in[pos];
      if (p.in_scale) v = gl::mul(v, p.in_pow.get((uint32_t)pos));
    }
    sm[at(u, d)] = v;
  }
  __syncthreads();
  dif_all(sm, at, tw_r, b, units_log, false);
  if (!p.natural_out) {
    for (int e = threadIdx.x; e < R * U; e += blockDim.x) {
      const int d = e & (R - 1), u = e >> b;
      uint64_t v = sm[at(u, d)];
      const size_t pos = ((size_t)(q0 + u) << b) | d;
      if (p.out_scale == 2) v = gl::mul(v, p.out_const);
      // (out_scale == 1 needs the natural index: position p holds out[bitrev(p)])
      else if (p.out_scale == 1) v = gl::mul(v, p.out_pow.get(gl::bitrev32((uint32_t)pos, p.log_n)));
      out[pos] = gl::canon(v);
    }
  } else {
    for (int e = threadIdx.x; e < R * U; e += blockDim.x) {
      const int u = e & (U - 1), k = e >> units_log;
      uint64_t v = sm[at(u, gl::bitrev32(k, b))];
      const uint32_t idx = ((uint32_t)k << pb) | (q0 + u);
      if (p.out_scale == 2) v = gl::mul(v, p.out_const);
      else if (p.out_scale == 1) v = gl::mul(v, p.out_pow.get(idx));
      out[idx] = gl::canon(v);
    }
  }
}

}  // namespace ntt
