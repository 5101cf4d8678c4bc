// Device kernels of the starky / FRI part of the path (K6, K7, K8 in SURVEY.md 2.4).
//
// Replaces (third-party crates, not on disk; pins at /root/reference/Cargo.lock:3441,4529,1675; reached
// from /root/reference/ops/src/lib.rs:52):
//   starky 0.4.0   src/prover.rs              compute_quotient_polys
//                  src/constraint_consumer.rs ConstraintConsumer
//                  src/lookup.rs              lookup_helper_columns, eval_packed_lookups_generic
//                  src/proof.rs               StarkOpeningSet::new
//                  src/fibonacci_stark.rs     FibonacciStark::eval_packed_generic
//   evm_arithmetization 0.1.3 src/memory/memory_stark.rs  (memory-shaped table)
//   plonky2 0.2.2  src/fri/oracle.rs          prove_openings (reduce_polys_base, divide_by_linear, lde)
//                  src/fri/prover.rs          fri_committed_trees (fold), fri_proof_of_work
// All matrices are column-major; LDE matrices are in bit-reversed row order (position p <-> point
// index bitrev(p)), which is plonky2's leaf order.
#pragma once
#include <cuda_runtime.h>

#include "gl.cuh"
#include "ntt.cuh"
#include "poseidon.cuh"
#include "quotient_rt.cuh"

namespace stark {

constexpr uint64_t MEM_TRIE_DATA_SEGMENT = 13;
enum { M_FILTER = 0, M_TIMESTAMP, M_IS_READ, M_CTX, M_SEG, M_VIRT, M_VALUE0, M_CFC = 14, M_SFC, M_VFC, M_INIT_AUX,
       M_RANGE_CHECK, M_COUNTER, M_FREQ };

template <int TABLE>
struct Table;

// starky/src/fibonacci_stark.rs
template <>
struct Table<0> {
  static constexpr int COLS = 2, AUX = 0;
  static constexpr int CONSTRAINTS = 5, LOOKUP_CONSTRAINTS_PER_CHALLENGE = 0;
  __device__ static __forceinline__ void eval(const uint64_t* lv, const uint64_t* nv, const uint64_t* pi, Consumer& c) {
    c.first_row(gl::sub(lv[0], pi[0]));
    c.first_row(gl::sub(lv[1], pi[1]));
    c.last_row(gl::sub(lv[1], pi[2]));
    c.transition(gl::sub(nv[0], lv[1]));
    c.transition(gl::sub(gl::sub(nv[1], lv[0]), lv[1]));
  }
  __device__ static __forceinline__ void eval_lookups(const uint64_t*, const uint64_t*, const uint64_t*, const uint64_t*, int, Consumer&) {}
};

// memory-shaped table: constraint list and order of SURVEY.md Appendix A
template <>
struct Table<1> {
  static constexpr int COLS = 21, AUX = 4;
  static constexpr int CONSTRAINTS = 40, LOOKUP_CONSTRAINTS_PER_CHALLENGE = 3;
  __device__ static __forceinline__ void eval(const uint64_t* lv, const uint64_t* nv, const uint64_t*, Consumer& c) {
    const uint64_t one = 1;
    const uint64_t filter = lv[M_FILTER];
    c.constraint(gl::mul(filter, gl::sub(filter, one)));
    c.constraint(gl::mul(gl::sub(one, filter), gl::sub(one, lv[M_IS_READ])));
    const uint64_t cfc = lv[M_CFC], sfc = lv[M_SFC], vfc = lv[M_VFC];
    const uint64_t unchanged = gl::sub(gl::sub(gl::sub(one, cfc), sfc), vfc);
    c.constraint(gl::mul(cfc, gl::sub(one, cfc)));
    c.constraint(gl::mul(sfc, gl::sub(one, sfc)));
    c.constraint(gl::mul(vfc, gl::sub(one, vfc)));
    c.constraint(gl::mul(unchanged, gl::sub(one, unchanged)));
    const uint64_t d_ctx = gl::sub(nv[M_CTX], lv[M_CTX]), d_seg = gl::sub(nv[M_SEG], lv[M_SEG]);
    const uint64_t d_virt = gl::sub(nv[M_VIRT], lv[M_VIRT]), d_ts = gl::sub(nv[M_TIMESTAMP], lv[M_TIMESTAMP]);
    c.transition(gl::mul(sfc, d_ctx));
    c.transition(gl::mul(vfc, d_ctx));
    c.transition(gl::mul(vfc, d_seg));
    c.transition(gl::mul(unchanged, d_ctx));
    c.transition(gl::mul(unchanged, d_seg));
    c.transition(gl::mul(unchanged, d_virt));
    const uint64_t computed =
        gl::add(gl::add(gl::mul(cfc, gl::sub(d_ctx, one)), gl::mul(sfc, gl::sub(d_seg, one))),
                gl::add(gl::mul(vfc, gl::sub(d_virt, one)), gl::mul(unchanged, d_ts)));
    c.transition(gl::sub(lv[M_RANGE_CHECK], computed));
    const uint64_t init_aux = lv[M_INIT_AUX];
    c.transition(gl::sub(init_aux, gl::mul(gl::mul(nv[M_SEG], gl::sub(one, unchanged)), nv[M_IS_READ])));
    const uint64_t read_unchanged = gl::mul(nv[M_IS_READ], unchanged);
    const uint64_t ctx_init = gl::mul(nv[M_CTX], init_aux);
    const uint64_t seg_init = gl::mul(gl::sub(nv[M_SEG], MEM_TRIE_DATA_SEGMENT), init_aux);
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const uint64_t v = lv[M_VALUE0 + i], nvv = nv[M_VALUE0 + i];
      c.transition(gl::mul(read_unchanged, gl::sub(nvv, v)));
      c.transition(gl::mul(ctx_init, nvv));
      c.transition(gl::mul(seg_init, nvv));
    }
    c.first_row(lv[M_COUNTER]);
    c.transition(gl::sub(gl::sub(nv[M_COUNTER], lv[M_COUNTER]), one));
  }
  // eval_packed_lookups_generic: one lookup (RANGE_CHECK in COUNTER, multiplicities FREQUENCIES), per challenge
  // one helper column h = 1/(f + ch) and one Z column
  __device__ static __forceinline__ void eval_lookups(const uint64_t* lv, const uint64_t* al, const uint64_t* an,
                                                      const uint64_t* challenges, int n_ch, Consumer& c) {
#pragma unroll
    for (int k = 0; k < MAX_CHALLENGES; k++) {
      if (k >= n_ch) break;
      const uint64_t ch = challenges[k];
      const uint64_t h = al[2 * k], z = al[2 * k + 1], next_z = an[2 * k + 1];
      c.constraint(gl::sub(gl::mul(gl::add(lv[M_RANGE_CHECK], ch), h), 1));
      const uint64_t twc = gl::add(lv[M_COUNTER], ch);
      const uint64_t y = gl::sub(gl::mul(h, twc), lv[M_FREQ]);
      c.first_row(z);
      c.constraint(gl::sub(gl::mul(gl::sub(next_z, z), twc), y));
    }
  }
};

// One thread per position p of the quotient coset inside the LDE (p < size): local row = LDE row p,
// next row = the row of point index k + next_step*step.
template <int TABLE, bool SPLIT = false>
static __global__ void __launch_bounds__(128) quotient_kernel(const __grid_constant__ QuotientParams q) {
  using T = Table<TABLE>;
  RowCtxT<SPLIT> r;
  if (!quotient_begin(q, r)) return;
  uint64_t lv[T::COLS], nv[T::COLS];
#pragma unroll
  for (int c = 0; c < T::COLS; c++) {
    lv[c] = r.lv(q, c);
    nv[c] = r.nv(q, c);
  }
  T::eval(lv, nv, q.pi, r.cs);
  if (T::AUX > 0) {
    uint64_t al[T::AUX > 0 ? T::AUX : 1], an[T::AUX > 0 ? T::AUX : 1];
#pragma unroll
    for (int c = 0; c < T::AUX; c++) {
      al[c] = r.la(q, c);
      an[c] = r.na(q, c);
    }
    T::eval_lookups(lv, al, an, q.lookup_ch, q.n_lookup_ch, r.cs);
  }
  quotient_end(q, r);
}

// ---- Lagrange selectors on the coset: L_first(x) = Z_H(x) / (n (x - 1)), L_last(x) = Z_H(x) / (n (g x - 1)) ----
// step 1: denominators (two per position), step 2: batch inverse, step 3: multiply by Z_H(x)
static __global__ void lagrange_denominators(int log_lde, int log_size, int step_log, ntt::PowTable coset, uint64_t n_field,
                                             uint64_t g, uint64_t* den /* 2 x size */) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t size = 1u << log_size;
  if (p >= size) return;
  const uint32_t i = gl::bitrev32(p, log_lde) >> step_log;
  const uint64_t x = coset.get(i);
  den[p] = gl::mul(n_field, gl::sub(x, 1));
  den[size + p] = gl::mul(n_field, gl::sub(gl::mul(g, x), 1));
}
struct ZhVals { uint64_t v[8]; };  // Z_H(x_i) per (i mod 2^qbits), qbits <= 3
static __global__ void lagrange_finish(int log_lde, int log_size, int step_log, int qmask, ZhVals zhv,
                                       uint64_t* inv /* 2 x size, in place */) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t size = 1u << log_size;
  if (p >= size) return;
  const uint32_t i = gl::bitrev32(p, log_lde) >> step_log;
  const uint64_t z = zhv.v[i & qmask];
  inv[p] = gl::canon(gl::mul(inv[p], z));
  inv[size + p] = gl::canon(gl::mul(inv[size + p], z));
}

// ---- batch inverse (Montgomery's trick inside each thread, K strided elements per thread) ----------
// A zero input (upstream: batch_multiplicative_inverse panics, "Tried to invert zero") raises *zero_flag; the strip's
// outputs are then meaningless and the caller reports ETP_ERR_PROOF at its next synchronisation point.
constexpr int INV_K = 8;
static __global__ void batch_inverse(const uint64_t* in, uint64_t* out, size_t n, unsigned long long* zero_flag) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  uint64_t v[INV_K], pre[INV_K];
  uint64_t acc = 1;
#pragma unroll
  for (int k = 0; k < INV_K; k++) {
    const size_t idx = t + k * stride;
    v[k] = idx < n ? in[idx] : 1;
    pre[k] = acc;
    acc = gl::mul(acc, v[k]);
  }
  if (gl::canon(acc) == 0) atomicOr(zero_flag, 1ull);
  uint64_t inv = gl::inv(acc);
#pragma unroll
  for (int k = INV_K - 1; k >= 0; k--) {
    const size_t idx = t + k * stride;
    if (idx < n) out[idx] = gl::canon(gl::mul(inv, pre[k]));
    inv = gl::mul(inv, v[k]);
  }
}

// ---- lookup helper columns -----------------------------------------------------------------------
// den[j][i] = trace[cols[j]][i] + ch for the m looking columns of a lookup followed by its table column (j = m)
static __global__ void lookup_denominators(const uint64_t* __restrict__ trace, size_t stride, const int* __restrict__ cols, int n_cols,
                                           uint64_t ch, size_t n, uint64_t* __restrict__ den) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= n * n_cols) return;
  const size_t j = t / n, i = t % n;
  den[t] = gl::add(trace[(size_t)cols[j] * stride + i], ch);
}
// inv: (m + 1) x n inverted denominators.  helper c = sum of the inverses of chunk c (`chunk` looking columns
// each); term[i] = sum_c helper_c[i] - freq[i] * inv[m][i]  (the increment of Z)
static __global__ void lookup_terms(const uint64_t* __restrict__ inv, int m, int chunk, const uint64_t* __restrict__ freq, size_t n,
                                    uint64_t* __restrict__ h_out /* ceil(m / chunk) columns, stride n */, uint64_t* __restrict__ term) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t total = 0;
  int c = 0;
  for (int j0 = 0; j0 < m; j0 += chunk, c++) {
    uint64_t h = 0;
    for (int j = j0; j < j0 + chunk && j < m; j++) h = gl::add(h, inv[(size_t)j * n + i]);
    h = gl::canon(h);
    h_out[(size_t)c * n + i] = h;
    total = gl::add(total, h);
  }
  term[i] = gl::canon(gl::sub(total, gl::mul(freq[i], inv[(size_t)m * n + i])));
}
// ---- general auxiliary columns: Lookups with linear-combination Columns / Filters, CTL Z data -----------------------
// (starky/src/lookup.rs Column::eval_table, Filter::eval_table, get_helper_cols; cross_table_lookup.rs partial_sums)
// The descriptors are the aux-spec words of include/etp_b200.h, interpreted per row; every thread walks the same words
// (uniform, cached loads), the trace reads are coalesced along the rows.
//   Column: n_local, (col, coeff)*, n_next, (col, coeff)*, constant       Filter: n_prod, (Column, Column)*, n_const, Column*
__device__ __forceinline__ uint64_t aux_eval_column(const uint64_t* __restrict__& w, const uint64_t* __restrict__ trace, size_t stride,
                                                    uint32_t i, uint32_t i_next) {
  uint64_t acc = 0;
  const uint32_t nl = (uint32_t)__ldg(w++);
  for (uint32_t k = 0; k < nl; k++) {
    const uint64_t col = __ldg(w++), coeff = __ldg(w++);
    const uint64_t v = __ldg(trace + col * stride + i);
    acc = gl::add(acc, coeff == 1 ? v : gl::mul(v, coeff));
  }
  const uint32_t nn = (uint32_t)__ldg(w++);
  for (uint32_t k = 0; k < nn; k++) {
    const uint64_t col = __ldg(w++), coeff = __ldg(w++);
    const uint64_t v = __ldg(trace + col * stride + i_next);
    acc = gl::add(acc, coeff == 1 ? v : gl::mul(v, coeff));
  }
  return gl::add(acc, __ldg(w++));
}
__device__ __forceinline__ uint64_t aux_eval_filter(const uint64_t* __restrict__& w, const uint64_t* __restrict__ trace, size_t stride,
                                                    uint32_t i, uint32_t i_next) {
  uint64_t acc = 0;
  const uint32_t np = (uint32_t)__ldg(w++);
  for (uint32_t k = 0; k < np; k++) {
    const uint64_t a = aux_eval_column(w, trace, stride, i, i_next);
    const uint64_t b = aux_eval_column(w, trace, stride, i, i_next);
    acc = gl::add(acc, gl::mul(a, b));
  }
  const uint32_t nc = (uint32_t)__ldg(w++);
  for (uint32_t k = 0; k < nc; k++) acc = gl::add(acc, aux_eval_column(w, trace, stride, i, i_next));
  return acc;
}
// One (columns, filter) pair on every row: den[i] = reduce_with_powers(column evals, beta) + gamma, filt[i] = filter(i).
// cols: n_cols Column descriptors back to back; filter: one Filter descriptor.
static __global__ void aux_colset_eval(const uint64_t* __restrict__ trace, size_t stride, uint32_t n, const uint64_t* __restrict__ cols,
                                       int n_cols, const uint64_t* __restrict__ filter, uint64_t beta, uint64_t gamma,
                                       uint64_t* __restrict__ den, uint64_t* __restrict__ filt) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t i_next = (i + 1 == n) ? 0 : i + 1;
  // sum_j col_j * beta^j, ascending with a running power (the descriptors can only be walked forwards)
  const uint64_t* w = cols;
  uint64_t acc = 0, bp = 1;
  for (int j = 0; j < n_cols; j++) {
    const uint64_t v = aux_eval_column(w, trace, stride, i, i_next);
    acc = gl::add(acc, j == 0 ? v : gl::mul(v, bp));
    if (j + 1 < n_cols) bp = j == 0 ? beta : gl::mul(bp, beta);
  }
  den[i] = gl::add(acc, gamma);
  const uint64_t* f = filter;
  filt[i] = aux_eval_filter(f, trace, stride, i, i_next);
}
// a single Column on every row
static __global__ void aux_column_eval(const uint64_t* __restrict__ trace, size_t stride, uint32_t n, const uint64_t* __restrict__ col,
                                       uint64_t add_const, uint64_t* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t* w = col;
  out[i] = gl::add(aux_eval_column(w, trace, stride, i, (i + 1 == n) ? 0 : i + 1), add_const);
}
// helper columns of one Lookup / CtlZData from the inverted denominators: helper c = sum over its chunk of filt_s * inv_s;
// term[i] = sum_c helper_c[i] (- freq[i] * table_inv[i] for a Lookup).  h_out == nullptr: helpers are not kept (single colset).
static __global__ void aux_helper_terms(const uint64_t* __restrict__ inv, const uint64_t* __restrict__ filt, int n_sets, int chunk, size_t n,
                                        const uint64_t* __restrict__ freq, const uint64_t* __restrict__ table_inv,
                                        uint64_t* __restrict__ h_out, uint64_t* __restrict__ term) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t total = 0;
  int c = 0;
  for (int s0 = 0; s0 < n_sets; s0 += chunk, c++) {
    uint64_t h = 0;
    for (int s = s0; s < s0 + chunk && s < n_sets; s++) h = gl::add(h, gl::mul(inv[(size_t)s * n + i], filt[(size_t)s * n + i]));
    h = gl::canon(h);
    if (h_out) h_out[(size_t)c * n + i] = h;
    total = gl::add(total, h);
  }
  if (freq) total = gl::sub(total, gl::mul(freq[i], table_inv[i]));
  term[i] = gl::canon(total);
}
// partial_sums: z[i] = sum_{j >= i} term[j] = total - exclusive_prefix[i], total = prefix[n-1] + term[n-1]
static __global__ void suffix_from_prefix(const uint64_t* __restrict__ prefix, const uint64_t* __restrict__ term, size_t n,
                                          uint64_t* __restrict__ z) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t total = gl::add(prefix[n - 1], term[n - 1]);
  z[i] = gl::canon(gl::sub(total, prefix[i]));
}
// exclusive prefix sum over the field, 3 phases; SCAN_BLOCK elements per block
constexpr int SCAN_THREADS = 256, SCAN_PER_THREAD = 8, SCAN_BLOCK = SCAN_THREADS * SCAN_PER_THREAD;
static __global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums(const uint64_t* __restrict__ in, size_t n, uint64_t* __restrict__ sums) {
  __shared__ uint64_t sh[SCAN_THREADS];
  const size_t base = (size_t)blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_PER_THREAD;
  uint64_t acc = 0;
#pragma unroll
  for (int k = 0; k < SCAN_PER_THREAD; k++)
    if (base + k < n) acc = gl::add(acc, in[base + k]);
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = SCAN_THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] = gl::add(sh[threadIdx.x], sh[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) sums[blockIdx.x] = gl::canon(sh[0]);
}
static __global__ void scan_sums_serial(uint64_t* sums, size_t n_blocks) {  // tiny: exclusive scan in place
  if (blockIdx.x || threadIdx.x) return;
  uint64_t acc = 0;
  for (size_t i = 0; i < n_blocks; i++) { const uint64_t v = sums[i]; sums[i] = acc; acc = gl::canon(gl::add(acc, v)); }
}
static __global__ void __launch_bounds__(SCAN_THREADS) scan_finish(const uint64_t* __restrict__ in, size_t n, const uint64_t* __restrict__ sums,
                                                                   uint64_t* __restrict__ out) {
  __shared__ uint64_t sh[SCAN_THREADS];
  const size_t base = (size_t)blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_PER_THREAD;
  uint64_t v[SCAN_PER_THREAD];
  uint64_t acc = 0;
#pragma unroll
  for (int k = 0; k < SCAN_PER_THREAD; k++) { v[k] = base + k < n ? in[base + k] : 0; acc = gl::add(acc, v[k]); }
  sh[threadIdx.x] = acc;
  __syncthreads();
  // exclusive scan of the per-thread sums (Hillis-Steele; 256 entries)
  uint64_t mine = acc;
  for (int s = 1; s < SCAN_THREADS; s <<= 1) {
    uint64_t add = threadIdx.x >= s ? sh[threadIdx.x - s] : 0;
    __syncthreads();
    mine = gl::add(mine, add);
    sh[threadIdx.x] = mine;
    __syncthreads();
  }
  uint64_t run = gl::add(sums[blockIdx.x], gl::sub(mine, acc));  // exclusive prefix of this thread
#pragma unroll
  for (int k = 0; k < SCAN_PER_THREAD; k++) {
    if (base + k < n) out[base + k] = gl::canon(run);
    run = gl::add(run, v[k]);
  }
}

// ---- openings: evaluate every polynomial of a batch at two extension points ------------------------
// partial[block][poly][point] (ext).  pw0 / pw1: powers of the two points via two-level ext tables.
struct ExtPowTable {
  const uint64_t* lo;  // interleaved ext, 2^lo_bits entries
  const uint64_t* hi;
  int lo_bits;
  __device__ __forceinline__ gl::Ext get(uint32_t e) const {
    const uint32_t a = e >> lo_bits, b = e & ((1u << lo_bits) - 1);
    return gl::emul(gl::ext(hi[2 * a], hi[2 * a + 1]), gl::ext(lo[2 * b], lo[2 * b + 1]));
  }
};
constexpr int OPEN_THREADS = 256, OPEN_CHUNK = 32;
// (Acc192 / mac192 / reduce192: quotient_rt.cuh)
// grid: (ceil(n / (OPEN_THREADS*OPEN_CHUNK)), n_polys).  Thread t of a block owns the strip of OPEN_CHUNK consecutive
// coefficients starting at s = (block * OPEN_THREADS + t) * OPEN_CHUNK of ONE polynomial and computes
//   z^s * sum_b c_{s+b} z^b   for z = z0 and z = z1,
// the inner sums as four dot products (two extension components x two points) against the shared table z^b, b < OPEN_CHUNK,
// accumulated WITHOUT modular reduction (192-bit), i.e. 4 wide multiply-adds per coefficient instead of 4 field
// multiplications + 4 field additions.  The block then adds its strips; the host adds the blocks.
static __global__ void __launch_bounds__(OPEN_THREADS) eval_polys_at_two_points(const uint64_t* __restrict__ coeffs, size_t col_stride,
                                                                                int n_polys, uint32_t n, ExtPowTable t0, ExtPowTable t1,
                                                                                uint64_t* __restrict__ partial) {
  __shared__ uint64_t pw[OPEN_CHUNK * 4];
  __shared__ uint64_t sh[OPEN_THREADS * 4];
  const int poly = blockIdx.y;
  if (threadIdx.x < OPEN_CHUNK) {
    const gl::Ext a = gl::ecanon(t0.get(threadIdx.x)), b = gl::ecanon(t1.get(threadIdx.x));
    pw[4 * threadIdx.x + 0] = a.c0; pw[4 * threadIdx.x + 1] = a.c1;
    pw[4 * threadIdx.x + 2] = b.c0; pw[4 * threadIdx.x + 3] = b.c1;
  }
  __syncthreads();
  const uint32_t start = (blockIdx.x * OPEN_THREADS + threadIdx.x) * OPEN_CHUNK;
  gl::Ext r0 = gl::ext(0, 0), r1 = gl::ext(0, 0);
  if (start < n) {
    Acc192 a[4] = {};
    const uint64_t* __restrict__ src = coeffs + (size_t)poly * col_stride + start;
    const int len = (n - start) < (uint32_t)OPEN_CHUNK ? (int)(n - start) : OPEN_CHUNK;
    if (len == OPEN_CHUNK) {
#pragma unroll 4
      for (int b = 0; b < OPEN_CHUNK; b += 2) {
        const ulonglong2 c = __ldg(reinterpret_cast<const ulonglong2*>(src + b));
#pragma unroll
        for (int w = 0; w < 4; w++) {
          mac192(a[w], c.x, pw[4 * b + w]);
          mac192(a[w], c.y, pw[4 * b + 4 + w]);
        }
      }
    } else {
      for (int b = 0; b < len; b++) {
        const uint64_t c = __ldg(src + b);
#pragma unroll
        for (int w = 0; w < 4; w++) mac192(a[w], c, pw[4 * b + w]);
      }
    }
    r0 = gl::emul(gl::ext(reduce192(a[0]), reduce192(a[1])), t0.get(start));
    r1 = gl::emul(gl::ext(reduce192(a[2]), reduce192(a[3])), t1.get(start));
  }
  sh[4 * threadIdx.x + 0] = r0.c0; sh[4 * threadIdx.x + 1] = r0.c1;
  sh[4 * threadIdx.x + 2] = r1.c0; sh[4 * threadIdx.x + 3] = r1.c1;
  __syncthreads();
  for (int s = OPEN_THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s)
      for (int w = 0; w < 4; w++) sh[4 * threadIdx.x + w] = gl::add(sh[4 * threadIdx.x + w], sh[4 * (threadIdx.x + s) + w]);
    __syncthreads();
  }
  if (threadIdx.x < 4) partial[((size_t)blockIdx.x * n_polys + poly) * 4 + threadIdx.x] = gl::canon(sh[threadIdx.x]);
}

// ---- prove_openings in evaluation form -----------------------------------------------------------
// PolynomialBatch::prove_openings for a general FriInstanceInfo (plonky2/src/fri/oracle.rs), point by point on the LDE coset:
//   final(x) = sum_b alpha^(shift_b) * (sum_{k < n_b} alpha^k f_{b,k}(x) - y_b) / (x - z_b),   shift_b = sum_{b' > b} n_b'
// which equals coset_fft(final_poly.lde()) of upstream's coefficient-form computation (reduce_polys_base, divide_by_linear,
// shift_poly).  Every committed column that occurs in some batch is read ONCE: `cols` lists the unique columns in order of
// first appearance with, per batch, the alpha power it carries there (or NONE).  A batch whose polynomial list is a prefix
// of batch 0's (starky: the g*zeta batch = trace ++ aux, a prefix of trace ++ aux ++ quotient) costs nothing extra: its sum
// is a snapshot of batch 0's accumulator.
constexpr int MAX_FRI_BATCHES = 4;
constexpr uint32_t COMBINE_NONE = 0xFFFFFFFFu;
struct CombineCol {
  const uint64_t* ptr;             // LDE column (bit-reversed rows)
  uint32_t idx[MAX_FRI_BATCHES];   // alpha power of this column in batch b, or COMBINE_NONE
};
struct CombineParams {
  const CombineCol* cols;  // device
  int n_cols, n_batches;
  int prefix_len[MAX_FRI_BATCHES];  // > 0: batch b (>= 1) == the first prefix_len[b] unique columns with batch 0's powers
  int log_lde;
  ntt::PowTable coset;         // 7 * w^k
  const uint64_t* alpha_pows;  // device: ext interleaved, alpha^k
  gl::Ext y[MAX_FRI_BATCHES], z[MAX_FRI_BATCHES], shift[MAX_FRI_BATCHES];
  uint64_t seven_zc1_sq[MAX_FRI_BATCHES];
  uint64_t* den;   // n_batches x lde_n: norms to invert (phase 1) / inverted norms (phase 2)
  uint64_t* out;   // lde_n ext interleaved, bit-reversed order
  // wide tables (hundreds of columns on few rows): the column sums are split over blockIdx.y chunks of `col_chunk`
  // columns by combine_accumulate into partial[chunk][p][batch][c0, c1]; n_chunks == 0: inline
  uint64_t* partial;
  int col_chunk, n_chunks;
};
// acc[b] += sum over the unique columns [u0, u1) of alpha^idx * f(x_p).  The two components of alpha^k * v are accumulated
// without modular reduction (192-bit, reduced once at the end): two wide multiply-adds per column and batch.
template <int B>
__device__ __forceinline__ void combine_columns(const CombineParams& c, uint32_t p, int u0, int u1, gl::Ext (&acc)[B]) {
  Acc192 s[B][2] = {};
  Acc192 snap[B][2] = {};
  constexpr int G = 4;  // columns in flight per thread: the loads of a group are issued before its multiply-adds
  for (int ug = u0; ug < u1; ug += G) {
    const uint64_t* ptr[G];
    uint64_t v[G];
#pragma unroll
    for (int j = 0; j < G; j++)
      if (ug + j < u1) ptr[j] = reinterpret_cast<const uint64_t*>(__ldg(reinterpret_cast<const unsigned long long*>(&c.cols[ug + j].ptr)));
#pragma unroll
    for (int j = 0; j < G; j++)
      if (ug + j < u1) v[j] = __ldg(ptr[j] + p);
#pragma unroll
    for (int j = 0; j < G; j++) {
      const int u = ug + j;
      if (u >= u1) break;
#pragma unroll
      for (int b = 1; b < B; b++)
        if (c.prefix_len[b] == u) { snap[b][0] = s[0][0]; snap[b][1] = s[0][1]; }
#pragma unroll
      for (int b = 0; b < B; b++) {
        if (b >= c.n_batches || (b > 0 && c.prefix_len[b] > 0)) continue;
        const uint32_t k = __ldg(&c.cols[u].idx[b]);
        if (k == COMBINE_NONE) continue;
        mac192(s[b][0], v[j], __ldg(c.alpha_pows + 2 * (size_t)k));
        mac192(s[b][1], v[j], __ldg(c.alpha_pows + 2 * (size_t)k + 1));
      }
    }
  }
#pragma unroll
  for (int b = 0; b < B; b++) {
    if (b >= c.n_batches) continue;
    if (b > 0 && c.prefix_len[b] > 0) {
      // this range's share of the prefix sum: everything if it ends inside the prefix, the snapshot if it straddles the
      // end of the prefix, nothing if it starts after it
      const int L = c.prefix_len[b];
      if (L >= u1) acc[b] = gl::eadd(acc[b], gl::ext(reduce192(s[0][0]), reduce192(s[0][1])));
      else if (L > u0) acc[b] = gl::eadd(acc[b], gl::ext(reduce192(snap[b][0]), reduce192(snap[b][1])));
    } else {
      acc[b] = gl::eadd(acc[b], gl::ext(reduce192(s[b][0]), reduce192(s[b][1])));
    }
  }
}
template <int B>
static __global__ void __launch_bounds__(128) combine_accumulate(CombineParams c) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n = 1u << c.log_lde;
  if (p >= n) return;
  const int u0 = blockIdx.y * c.col_chunk, u1 = (u0 + c.col_chunk) < c.n_cols ? (u0 + c.col_chunk) : c.n_cols;
  gl::Ext acc[B];
#pragma unroll
  for (int b = 0; b < B; b++) acc[b] = gl::ext(0, 0);
  combine_columns<B>(c, p, u0, u1, acc);
  uint64_t* dst = c.partial + ((size_t)blockIdx.y * n + p) * (2 * B);
#pragma unroll
  for (int b = 0; b < B; b++) { dst[2 * b] = acc[b].c0; dst[2 * b + 1] = acc[b].c1; }
}
static __global__ void combine_norms(CombineParams c) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n = 1u << c.log_lde;
  if (p >= n) return;
  const uint64_t x = c.coset.get(gl::bitrev32(p, c.log_lde));
  // norm of (x - z) = (x - z.c0)^2 - 7 z.c1^2   (7 z.c1^2 is precomputed on the host)
  for (int b = 0; b < c.n_batches; b++) {
    const uint64_t a = gl::sub(x, c.z[b].c0);
    c.den[(size_t)b * n + p] = gl::canon(gl::sub(gl::mul(a, a), c.seven_zc1_sq[b]));
  }
}
template <int B>
static __global__ void __launch_bounds__(128) combine_values(CombineParams c) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n = 1u << c.log_lde;
  if (p >= n) return;
  gl::Ext acc[B];
#pragma unroll
  for (int b = 0; b < B; b++) acc[b] = gl::ext(0, 0);
  if (c.n_chunks == 0) {
    combine_columns<B>(c, p, 0, c.n_cols, acc);
  } else {
    for (int j = 0; j < c.n_chunks; j++) {
      const uint64_t* src = c.partial + ((size_t)j * n + p) * (2 * B);
#pragma unroll
      for (int b = 0; b < B; b++) acc[b] = gl::eadd(acc[b], gl::ext(src[2 * b], src[2 * b + 1]));
    }
  }
  const uint64_t x = c.coset.get(gl::bitrev32(p, c.log_lde));
  gl::Ext f = gl::ext(0, 0);
#pragma unroll
  for (int b = 0; b < B; b++) {
    if (b >= c.n_batches) continue;
    // 1/(x - z) = conj(x - z) / norm
    const uint64_t i = c.den[(size_t)b * n + p];
    const gl::Ext inv = gl::ext(gl::mul(gl::sub(x, c.z[b].c0), i), gl::mul(c.z[b].c1, i));
    gl::Ext q = gl::emul(gl::esub(acc[b], c.y[b]), inv);
    if (b + 1 < c.n_batches) q = gl::emul(q, c.shift[b]);
    f = gl::eadd(f, q);
  }
  f = gl::ecanon(f);
  c.out[2 * (size_t)p] = f.c0;
  c.out[2 * (size_t)p + 1] = f.c1;
}

// ---- FRI fold in evaluation form (arity 2^4) -------------------------------------------------------
// Leaf j of a layer holds the 16 values at points x0 * w16^bitrev4(t), x0 = shift * w_n^bitrev(j).
// With u = IDFT16(values in natural order m), the folded value at y = x0^16 is sum_i (beta/x0)^i u_i —
// the value at y of upstream's reduce_with_powers(coefficient chunks, beta).  Output position j is the
// bit-reversed position of y in the next layer.
struct FoldParams {
  const uint64_t* in;   // n ext interleaved
  uint64_t* out;        // n/16 ext interleaved
  int log_n;            // current layer size
  ntt::PowTable x0_inv; // shift^-1 * (w_n^-1)^k
  gl::Ext beta;
  uint64_t w16_inv_pows[16];  // (w16^-1)^e
  uint64_t inv16;
};
// (Measured, profiles/r02_ncu_stark_kernels_2p22.csv: staging the leaves through padded shared memory for fully coalesced
// loads does not help — layer 0 is bound by the integer ALU (66 %), not by its 134 MB of reads; the L1 absorbs the
// per-thread 256-byte leaf reads.)
static __global__ void __launch_bounds__(128) fri_fold16(FoldParams f) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n_out = 1u << (f.log_n - 4);
  if (j >= n_out) return;
  gl::Ext v[16];
#pragma unroll
  for (int t = 0; t < 16; t++) {
    const ulonglong2 e = reinterpret_cast<const ulonglong2*>(f.in)[16 * (size_t)j + t];
    v[gl::bitrev32(t, 4)] = gl::ext(e.x, e.y);  // natural order m = bitrev4(t)
  }
  gl::Ext gamma = gl::emul_base(f.beta, f.x0_inv.get(gl::bitrev32(j, f.log_n - 4)));
  // result = (1/16) * sum_i gamma^i * sum_m v_m w16^(-i m), computed as four arity-2 folds with gamma, gamma^2,
  // gamma^4, gamma^8:  v'[m] = (v[m] + v[m+h]) + gamma_k * (v[m] - v[m+h]) * w_{2h}^(-m)   (15 pair folds instead of
  // the 16 x 16 matrix; same field elements)
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int h = 8 >> k;
#pragma unroll
    for (int m = 0; m < h; m++) {
      const gl::Ext a = v[m], c = v[m + h];
      gl::Ext d = gl::esub(a, c);
      if (m) d = gl::emul_base(d, f.w16_inv_pows[m << k]);
      v[m] = gl::eadd(gl::eadd(a, c), gl::emul(d, gamma));
    }
    if (k < 3) gamma = gl::emul(gamma, gamma);
  }
  gl::Ext acc = v[0];
  acc = gl::ecanon(gl::emul_base(acc, f.inv16));
  reinterpret_cast<ulonglong2*>(f.out)[j] = make_ulonglong2(acc.c0, acc.c1);
}

// ---- proof of work --------------------------------------------------------------------------------
static __global__ void __launch_bounds__(128) pow_grind(const uint64_t* __restrict__ state, int pos, int bits, uint64_t base,
                                                        unsigned long long* __restrict__ result) {
  const uint64_t cand = base + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (cand >= gl::P) return;
  uint64_t s[12];
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = (i == pos) ? cand : state[i];
  poseidon::permute(s);
  const uint64_t resp = gl::canon(s[7]);
  if (bits == 0 || (resp >> (64 - bits)) == 0) atomicMin(result, (unsigned long long)cand);
}

static __global__ void canon_copy(const uint64_t* __restrict__ src, uint64_t* __restrict__ dst, size_t n) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = gl::canon(src[i]);
}

// ---- gathers for the query phase --------------------------------------------------------------------
// out[q][i] = levels[i][(idx[q] >> i) ^ 1], i < num_layers
static __global__ void gather_paths(const uint64_t* __restrict__ levels, uint32_t n_leaves, int num_layers,
                                    const uint64_t* __restrict__ idx, int n_idx, uint64_t* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_idx * num_layers) return;
  const int q = t / num_layers, i = t % num_layers;
  size_t off = 0;
  for (int l = 0; l < i; l++) off += 4 * (size_t)(n_leaves >> l);
  const uint64_t node = (idx[q] >> i) ^ 1;
  const ulonglong2* src = reinterpret_cast<const ulonglong2*>(levels + off + 4 * node);
  ulonglong2* dst = reinterpret_cast<ulonglong2*>(out + 4 * (size_t)t);
  dst[0] = src[0];
  dst[1] = src[1];
}
static __global__ void gather_rows_rowmajor(const uint64_t* __restrict__ rows, int row_len, const uint64_t* __restrict__ idx, int n_idx,
                                            uint64_t* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_idx * row_len) return;
  const int q = t / row_len, c = t % row_len;
  out[t] = gl::canon(rows[(size_t)idx[q] * row_len + c]);
}

}  // namespace stark
