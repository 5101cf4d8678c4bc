// Runtime of the quotient kernels: starky's ConstraintConsumer, the kernel parameters and the per-row prologue /
// epilogue of compute_quotient_polys (starky 0.4.0 src/prover.rs, src/constraint_consumer.rs; pins at
// /root/reference/Cargo.lock:4529, reached from /root/reference/ops/src/lib.rs:52).  Shared by the built-in
// tables (stark_kernels.cuh, nvcc) and by the constraint programs compiled with NVRTC when a table is registered
// (etp_jit.cu) — this header, gl.cuh and powtable.cuh are embedded in the library as source text for that.
#pragma once
#include "gl.cuh"
#include "powtable.cuh"

namespace stark {

constexpr int MAX_CHALLENGES = 2;
constexpr int MAX_PUBLIC_INPUTS = 16;

// ---- ConstraintConsumer ------------------------------------------------------------------------
struct Consumer {
  uint64_t alphas[MAX_CHALLENGES], acc[MAX_CHALLENGES];
  uint64_t z_last, lagrange_first, lagrange_last;
  int n;
#if defined(ETP_COMPACT_CODE)
  __device__ __noinline__ void constraint(uint64_t c) {
#else
  __device__ __forceinline__ void constraint(uint64_t c) {
#endif
#pragma unroll
    for (int j = 0; j < MAX_CHALLENGES; j++)
      if (j < n) acc[j] = gl::add(gl::mul(acc[j], alphas[j]), c);
  }
  __device__ __forceinline__ void transition(uint64_t c) { constraint(gl::mul(c, z_last)); }
  __device__ __forceinline__ void first_row(uint64_t c) { constraint(gl::mul(c, lagrange_first)); }
  __device__ __forceinline__ void last_row(uint64_t c) { constraint(gl::mul(c, lagrange_last)); }
};

struct QuotientParams {
  const uint64_t* trace;  // LDE, bit-reversed rows
  size_t trace_stride;
  const uint64_t* aux;
  size_t aux_stride;
  int log_lde;      // degree_bits + rate_bits
  int log_size;     // degree_bits + quotient_degree_bits
  int step_log;     // rate_bits - quotient_degree_bits
  int next_step;    // 1 << quotient_degree_bits
  ntt::PowTable coset;  // 7 * w_size^i
  const uint64_t* lag_first;  // per position p < size: L_first, L_last at the point of position p
  const uint64_t* lag_last;
  uint64_t zh_inv[4];    // 1/Z_H per (i mod 2^qbits)
  uint64_t last;         // g^-1
  uint64_t alphas[MAX_CHALLENGES];
  int n_alphas;
  uint64_t lookup_ch[MAX_CHALLENGES];
  int n_lookup_ch;
  uint64_t pi[MAX_PUBLIC_INPUTS];
  uint64_t* out;  // n_alphas columns x size, NATURAL order (input of coset_ifft)
};


// per-thread row context: position p of the quotient coset inside the LDE, its "next" row, and the consumer
struct RowCtx {
  uint32_t p, p_next, i;
  Consumer cs;
  __device__ __forceinline__ uint64_t lv(const QuotientParams& q, int c) const { return __ldg(q.trace + (size_t)c * q.trace_stride + p); }
  __device__ __forceinline__ uint64_t nv(const QuotientParams& q, int c) const { return __ldg(q.trace + (size_t)c * q.trace_stride + p_next); }
  __device__ __forceinline__ uint64_t la(const QuotientParams& q, int c) const { return __ldg(q.aux + (size_t)c * q.aux_stride + p); }
  __device__ __forceinline__ uint64_t na(const QuotientParams& q, int c) const { return __ldg(q.aux + (size_t)c * q.aux_stride + p_next); }
};
__device__ __forceinline__ bool quotient_begin(const QuotientParams& q, RowCtx& r) {
  r.p = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t size = 1u << q.log_size;
  if (r.p >= size) return false;
  const uint32_t k = gl::bitrev32(r.p, q.log_lde);                 // LDE point index, multiple of step
  r.i = k >> q.step_log;                                           // index on the quotient coset
  const uint32_t i_next = (r.i + q.next_step) & (size - 1);
  r.p_next = gl::bitrev32(i_next << q.step_log, q.log_lde);
  r.cs.n = q.n_alphas;
#pragma unroll
  for (int j = 0; j < MAX_CHALLENGES; j++) { r.cs.alphas[j] = q.alphas[j]; r.cs.acc[j] = 0; }
  const uint64_t x = q.coset.get(r.i);
  r.cs.z_last = gl::sub(x, q.last);
  r.cs.lagrange_first = q.lag_first[r.p];
  r.cs.lagrange_last = q.lag_last[r.p];
  return true;
}
__device__ __forceinline__ void quotient_end(const QuotientParams& q, const RowCtx& r) {
  const uint32_t size = 1u << q.log_size;
  const uint64_t dinv = q.zh_inv[r.i & (q.next_step - 1)];
#pragma unroll
  for (int j = 0; j < MAX_CHALLENGES; j++)
    if (j < q.n_alphas) q.out[(size_t)j * size + r.i] = gl::mul(r.cs.acc[j], dinv);
}

}  // namespace stark
