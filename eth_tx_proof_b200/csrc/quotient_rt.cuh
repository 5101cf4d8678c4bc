// Runtime of the quotient kernels: starky's ConstraintConsumer, the kernel parameters and the per-row prologue /
// epilogue of compute_quotient_polys (starky 0.4.0 src/prover.rs, src/constraint_consumer.rs; pins at
// /root/reference/Cargo.lock:4529, reached from /root/reference/ops/src/lib.rs:52).  Shared by the built-in
// tables (stark_kernels.cuh, nvcc) and by the constraint programs compiled with NVRTC when a table is registered
// (etp_jit.cu) — this header, gl.cuh and powtable.cuh are embedded in the library as source text for that.
#pragma once
#include "gl.cuh"
#include "powtable.cuh"

namespace stark {

constexpr int MAX_CHALLENGES = 2;   // StarkConfig::num_challenges: alphas, lookup challenges, CTL (beta, gamma) pairs
constexpr int MAX_CH_SCALARS = 8;   // challenge scalars a program may read: lookup challenges [0, 2), then CTL beta_k, gamma_k
constexpr int MAX_PUBLIC_INPUTS = 16;

// ---- 192-bit lazy accumulator: a sum of products of two field elements, reduced once -------------------------
struct Acc192 {
  uint64_t w0, w1;
  uint32_t w2;
};
__device__ __forceinline__ void mac192(Acc192& a, uint64_t c, uint64_t u) {
  const unsigned __int128 p = (unsigned __int128)c * u;
  const uint64_t pl = (uint64_t)p, ph = (uint64_t)(p >> 64);
  asm("add.cc.u64 %0, %0, %3;\n\taddc.cc.u64 %1, %1, %4;\n\taddc.u32 %2, %2, 0;" : "+l"(a.w0), "+l"(a.w1), "+r"(a.w2) : "l"(pl), "l"(ph));
}
// w0 + 2^64 w1 + 2^128 w2 with 2^128 == -2^32 (mod p); w2 < 2^31
__device__ __forceinline__ uint64_t reduce192(const Acc192& a) {
  return gl::sub_c(gl::reduce128(a.w0, a.w1), (uint64_t)a.w2 << 32);
}

// ---- ConstraintConsumer ------------------------------------------------------------------------
// Upstream folds constraint i into acc_j <- acc_j * alpha_j + c_i * mult_i (mult = 1, x - g^-1, L_first, L_last).  With N
// constraints in all that is  sum_i c_i mult_i alpha_j^(N-1-i): here the powers come from a table (computed per proof on
// the host) and the sums are accumulated without modular reduction, one 192-bit accumulator per challenge for the plain
// constraints and one for the transition constraints (their common factor x - g^-1 is applied once at the end): a wide
// multiply-add per constraint and challenge instead of a field multiplication and a field addition.
struct Consumer {
  Acc192 plain[MAX_CHALLENGES], trans[MAX_CHALLENGES];
  const uint64_t* apow;  // [n][n_constraints]: alpha_j^e
  uint64_t z_last, lagrange_first, lagrange_last;
  int n, n_constraints, idx;
#if defined(ETP_COMPACT_CODE)
  __device__ __noinline__ void emit(uint64_t c, int is_transition) {
#else
  __device__ __forceinline__ void emit(uint64_t c, int is_transition) {
#endif
    const int e = n_constraints - 1 - idx;
    idx++;
#pragma unroll
    for (int j = 0; j < MAX_CHALLENGES; j++)
      if (j < n) {
        const uint64_t a = __ldg(apow + (size_t)j * n_constraints + e);
        if (is_transition) mac192(trans[j], c, a);
        else mac192(plain[j], c, a);
      }
  }
  __device__ __forceinline__ void constraint(uint64_t c) { emit(c, 0); }
  __device__ __forceinline__ void transition(uint64_t c) { emit(c, 1); }
  __device__ __forceinline__ void first_row(uint64_t c) { emit(gl::mul(c, lagrange_first), 0); }
  __device__ __forceinline__ void last_row(uint64_t c) { emit(gl::mul(c, lagrange_last), 0); }
  __device__ __forceinline__ uint64_t result(int j) const { return gl::add(reduce192(plain[j]), gl::mul(reduce192(trans[j]), z_last)); }
};

struct QuotientParams {
  const uint64_t* trace;  // LDE, bit-reversed rows
  size_t trace_stride;
  // column-split tables (etp_shard): the LDE base of every trace column, in this GPU's HBM or in a peer's (read over
  // NVLink); used instead of trace / trace_stride by the kernels compiled with SPLIT
  const uint64_t* const* trace_cols;
  const uint64_t* aux;
  size_t aux_stride;
  int log_lde;      // degree_bits + rate_bits
  int log_size;     // degree_bits + quotient_degree_bits
  int step_log;     // rate_bits - quotient_degree_bits
  int next_step;    // 1 << quotient_degree_bits
  ntt::PowTable coset;  // 7 * w_size^i
  const uint64_t* lag_first;  // per position p < size: L_first, L_last at the point of position p
  const uint64_t* lag_last;
  uint64_t zh_inv[8];    // 1/Z_H per (i mod 2^qbits), qbits <= 3
  uint64_t last;         // g^-1
  uint64_t alphas[MAX_CHALLENGES];
  int n_alphas;
  const uint64_t* alpha_pows;  // device, [n_alphas][n_constraints]: alpha_j^e (Consumer)
  int n_constraints;           // constraints the table emits per row, lookups included
  uint64_t lookup_ch[MAX_CH_SCALARS];  // lookup challenges, then the CTL challenges (programs index all of them)
  int n_lookup_ch;
  uint64_t pi[MAX_PUBLIC_INPUTS];
  uint64_t* out;  // n_alphas columns x size, NATURAL order (input of coset_ifft)
};


// per-thread row context: position p of the quotient coset inside the LDE, its "next" row, and the consumer
template <bool SPLIT>
struct RowCtxT {
  uint32_t p, p_next, i;
  Consumer cs;
  __device__ __forceinline__ const uint64_t* col(const QuotientParams& q, int c) const {
    if (SPLIT) return (const uint64_t*)__ldg((const unsigned long long*)q.trace_cols + c);
    return q.trace + (size_t)c * q.trace_stride;
  }
  __device__ __forceinline__ uint64_t lv(const QuotientParams& q, int c) const { return __ldg(col(q, c) + p); }
  __device__ __forceinline__ uint64_t nv(const QuotientParams& q, int c) const { return __ldg(col(q, c) + p_next); }
  __device__ __forceinline__ uint64_t la(const QuotientParams& q, int c) const { return __ldg(q.aux + (size_t)c * q.aux_stride + p); }
  __device__ __forceinline__ uint64_t na(const QuotientParams& q, int c) const { return __ldg(q.aux + (size_t)c * q.aux_stride + p_next); }
};
#if defined(ETP_SPLIT_COLUMNS)
typedef RowCtxT<true> RowCtx;
#else
typedef RowCtxT<false> RowCtx;
#endif
// Block order.  Position p = (H << 7) | t (t = thread) holds point index i = (bitrev_7(t) << m) | bitrev_m(H), m = log_size - 7,
// so the "next" row i + next_step of EVERY thread of a block lies in the block whose low index bits are I + next_step,
// I = bitrev_m(H).  Blocks are therefore issued along the chains I, I + next_step, I + 2 next_step, ...: the rows block b
// reads as `next` are the rows block b + 1 reads as `local`, the two run side by side and the second read is served by L2
// instead of DRAM (ncu, round 1: 3.35 GB read for 1.68 GB of LDE rows with the natural block order).
__device__ __forceinline__ uint32_t quotient_block_position(const QuotientParams& q) {
  const uint32_t b = blockIdx.x;
  const int m = q.log_size - 7;
  const int qb = 31 - __clz(q.next_step);  // next_step is a power of two
  if (blockDim.x != 128 || m < qb + 1) return b * blockDim.x + threadIdx.x;
  const uint32_t I = ((b << qb) & ((1u << m) - 1)) | (b >> (m - qb));
  return (gl::bitrev32(I, m) << 7) | threadIdx.x;
}
template <bool SPLIT>
__device__ __forceinline__ bool quotient_begin(const QuotientParams& q, RowCtxT<SPLIT>& r) {
  r.p = quotient_block_position(q);
  const uint32_t size = 1u << q.log_size;
  if (r.p >= size) return false;
  const uint32_t k = gl::bitrev32(r.p, q.log_lde);                 // LDE point index, multiple of step
  r.i = k >> q.step_log;                                           // index on the quotient coset
  const uint32_t i_next = (r.i + q.next_step) & (size - 1);
  r.p_next = gl::bitrev32(i_next << q.step_log, q.log_lde);
  r.cs.n = q.n_alphas;
  r.cs.n_constraints = q.n_constraints;
  r.cs.idx = 0;
  r.cs.apow = q.alpha_pows;
#pragma unroll
  for (int j = 0; j < MAX_CHALLENGES; j++) { r.cs.plain[j] = Acc192{0, 0, 0}; r.cs.trans[j] = Acc192{0, 0, 0}; }
  const uint64_t x = q.coset.get(r.i);
  r.cs.z_last = gl::sub(x, q.last);
  r.cs.lagrange_first = q.lag_first[r.p];
  r.cs.lagrange_last = q.lag_last[r.p];
  return true;
}
template <bool SPLIT>
__device__ __forceinline__ void quotient_end(const QuotientParams& q, const RowCtxT<SPLIT>& r) {
  const uint32_t size = 1u << q.log_size;
  const uint64_t dinv = q.zh_inv[r.i & (q.next_step - 1)];
#pragma unroll
  for (int j = 0; j < MAX_CHALLENGES; j++)
    if (j < q.n_alphas) q.out[(size_t)j * size + r.i] = gl::mul(r.cs.result(j), dinv);
}

}  // namespace stark
