// Device kernels of the first slice of plonky2's circuit prover (the recursion layers): the permutation
// argument's partial products and Z polynomials.
//
// Replaces plonky2 0.2.2 plonky2/src/plonk/prover.rs `wires_permutation_partial_products_and_zs` /
// `all_wires_permutation_partial_products` and plonk/vanishing_poly.rs-independent helpers
// `quotient_chunk_products`, `partial_products_and_z_gx` (util/partial_products.rs) — crate pinned at
// /root/reference/Cargo.lock:3441; this is what every shrink / root / aggregation / block proof of the reference
// (/root/reference/ops/src/lib.rs:52,72,95) runs between its wires commitment and its Z commitment.
// Row i is the subgroup point x_i = g^i (natural order), wires and sigmas are column-major.
#pragma once
#include <cuda_runtime.h>

#include "gl.cuh"
#include "powtable.cuh"

namespace plonk {

// den[j][i] = wire_j(i) + beta * sigma_j(i) + gamma   (to be batch-inverted)
static __global__ void permutation_denominators(const uint64_t* __restrict__ wires, size_t w_stride, const uint64_t* __restrict__ sigmas,
                                                size_t s_stride, int n_routed, uint32_t n, uint64_t beta, uint64_t gamma,
                                                uint64_t* __restrict__ den) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= (size_t)n * n_routed) return;
  const size_t j = t / n, i = t % n;
  den[t] = gl::add(gl::add(wires[j * w_stride + i], gl::mul(beta, sigmas[j * s_stride + i])), gamma);
}
// per row: q_c = prod over chunk c of (wire_j + beta k_j x + gamma) * den_inv_j ; chunk[c][i] = q_c ; rowprod[i] = prod_c q_c
static __global__ void permutation_chunk_products(const uint64_t* __restrict__ wires, size_t w_stride, const uint64_t* __restrict__ den_inv,
                                                  const uint64_t* __restrict__ k_is, int n_routed, int chunk_size, uint32_t n,
                                                  ntt::PowTable subgroup, uint64_t beta, uint64_t gamma, uint64_t* __restrict__ chunk,
                                                  uint64_t* __restrict__ rowprod) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t bx = gl::mul(beta, subgroup.get(i));
  uint64_t all = 1;
  int c = 0;
  for (int j0 = 0; j0 < n_routed; j0 += chunk_size, c++) {
    uint64_t q = 1;
    for (int j = j0; j < j0 + chunk_size && j < n_routed; j++) {
      const uint64_t num = gl::add(gl::add(wires[(size_t)j * w_stride + i], gl::mul(bx, __ldg(k_is + j))), gamma);
      q = gl::mul(q, gl::mul(num, den_inv[(size_t)j * n + i]));
    }
    chunk[(size_t)c * n + i] = q;
    all = gl::mul(all, q);
  }
  rowprod[i] = gl::canon(all);
}
// z[i] = exclusive running PRODUCT of rowprod (Z(x_0) = 1), three phases like the additive scan of stark_kernels.cuh
constexpr int PSCAN_THREADS = 256, PSCAN_PER_THREAD = 8, PSCAN_BLOCK = PSCAN_THREADS * PSCAN_PER_THREAD;
static __global__ void __launch_bounds__(PSCAN_THREADS) prod_block_totals(const uint64_t* __restrict__ in, size_t n, uint64_t* __restrict__ totals) {
  __shared__ uint64_t sh[PSCAN_THREADS];
  const size_t base = (size_t)blockIdx.x * PSCAN_BLOCK + threadIdx.x * PSCAN_PER_THREAD;
  uint64_t acc = 1;
#pragma unroll
  for (int k = 0; k < PSCAN_PER_THREAD; k++)
    if (base + k < n) acc = gl::mul(acc, in[base + k]);
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = PSCAN_THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] = gl::mul(sh[threadIdx.x], sh[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) totals[blockIdx.x] = gl::canon(sh[0]);
}
static __global__ void prod_totals_serial(uint64_t* totals, size_t n_blocks) {  // tiny: exclusive running product in place
  if (blockIdx.x || threadIdx.x) return;
  uint64_t acc = 1;
  for (size_t i = 0; i < n_blocks; i++) { const uint64_t v = totals[i]; totals[i] = acc; acc = gl::canon(gl::mul(acc, v)); }
}
static __global__ void __launch_bounds__(PSCAN_THREADS) prod_finish(const uint64_t* __restrict__ in, size_t n, const uint64_t* __restrict__ totals,
                                                                    uint64_t* __restrict__ out) {
  __shared__ uint64_t sh[PSCAN_THREADS];
  const size_t base = (size_t)blockIdx.x * PSCAN_BLOCK + threadIdx.x * PSCAN_PER_THREAD;
  uint64_t v[PSCAN_PER_THREAD];
  uint64_t acc = 1;
#pragma unroll
  for (int k = 0; k < PSCAN_PER_THREAD; k++) { v[k] = base + k < n ? in[base + k] : 1; acc = gl::mul(acc, v[k]); }
  sh[threadIdx.x] = acc;
  __syncthreads();
  // inclusive running product of the per-thread products (Hillis-Steele; 256 entries)
  uint64_t mine = acc;
  for (int s = 1; s < PSCAN_THREADS; s <<= 1) {
    const uint64_t other = threadIdx.x >= s ? sh[threadIdx.x - s] : 1;
    __syncthreads();
    mine = gl::mul(mine, other);
    sh[threadIdx.x] = mine;
    __syncthreads();
  }
  // exclusive prefix of this thread = inclusive prefix of the previous thread
  uint64_t run = gl::mul(totals[blockIdx.x], threadIdx.x ? sh[threadIdx.x - 1] : 1);
#pragma unroll
  for (int k = 0; k < PSCAN_PER_THREAD; k++) {
    if (base + k < n) out[base + k] = gl::canon(run);
    run = gl::mul(run, v[k]);
  }
}
// partial products of row i: pp_c(i) = Z(x_i) * q_0(i) * ... * q_c(i) for c < n_chunks - 1 (the last running product is
// Z(g x_i), which is the next row's Z and is not stored)
static __global__ void permutation_partial_products(const uint64_t* __restrict__ z, const uint64_t* __restrict__ chunk, int n_chunks, uint32_t n,
                                                    uint64_t* __restrict__ pp /* (n_chunks - 1) columns, stride n */) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t acc = z[i];
  for (int c = 0; c + 1 < n_chunks; c++) {
    acc = gl::mul(acc, chunk[(size_t)c * n + i]);
    pp[(size_t)c * n + i] = gl::canon(acc);
  }
}

}  // namespace plonk
