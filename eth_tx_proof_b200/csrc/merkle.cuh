// Poseidon Merkle trees on device (K4/K5 in SURVEY.md 2.4).
//
// Replaces plonky2 0.2.2 `MerkleTree::new` / `fill_digests_buf` / `fill_subtree` / `prove`
// (plonky2/src/hash/merkle_tree.rs) and `PoseidonHash::{hash_no_pad, hash_or_noop, two_to_one}`
// (plonky2/src/hash/poseidon.rs, hashing.rs) — crate pinned at /root/reference/Cargo.lock:3441,
// reached from /root/reference/ops/src/lib.rs:52 via PolynomialBatch::from_coeffs and fri_committed_trees.
//
// Device layout: digests are kept LEVEL BY LEVEL (level 0 = leaf digests, level i has n_leaves >> i
// nodes, the last level is the cap), 4 x u64 per digest.  A sibling is then levels[i][(leaf>>i)^1].
// plonky2's recursive `digests` layout (left subtree || left child || right child || right subtree,
// per cap subtree) is produced on demand by scatter_to_plonky2_layout with the closed-form map
//   node j of layer i  ->  2*(((j>>1) << (i+1)) + 2^i - 1) + (j&1)          (SURVEY.md 8(a), row M).
//
// Leaves are read straight from COLUMN-MAJOR matrices (leaf i = element i of every column), so the
// "transpose LDEs" + "reverse_index_bits_in_place" steps of from_coeffs disappear: the NTT already
// leaves every column in bit-reversed order and consecutive threads read consecutive addresses.
#pragma once
#include <cuda_runtime.h>

#include "poseidon.cuh"
#include "poseidon_coop.cuh"

namespace merkle {

#ifndef ETP_HASH_THREADS
#define ETP_HASH_THREADS 128
#endif
constexpr int HASH_THREADS = ETP_HASH_THREADS;
#ifndef ETP_LEAF_MIN_BLOCKS
#define ETP_LEAF_MIN_BLOCKS 7
#endif

__device__ __forceinline__ void store_digest(uint64_t* dst, const uint64_t (&s)[12]) {
  ulonglong2 a, b;
  a.x = gl::canon(s[0]); a.y = gl::canon(s[1]); b.x = gl::canon(s[2]); b.y = gl::canon(s[3]);
  reinterpret_cast<ulonglong2*>(dst)[0] = a;
  reinterpret_cast<ulonglong2*>(dst)[1] = b;
}

// Where the columns of a leaf live.  Column c is base[c / cols_per_src] + (c % cols_per_src) * col_stride:
// one source for an ordinary batch; one per GPU for a column-split batch, where the peers' LDE
// matrices are mapped through CUDA IPC and read over NVLink by the hashing kernel itself (the
// all-gather of row tiles is fused into the hash: SURVEY.md 8(e)).  cols_per_src is a multiple of 8
// whenever there is more than one source, so an 8-column sponge chunk never straddles two sources.
constexpr int MAX_SRC = 8;
struct LeafSrc {
  const uint64_t* base[MAX_SRC];
  size_t col_stride;
  int cols_per_src;
};

// leaf row0 + i = (col_0[row0 + i], col_1[row0 + i], ...) -> digests[i].  hash_or_noop semantics.
// The sponge (overwrite mode, rate 8) absorbs columns [c_begin, c_end) of n_cols_total; a call with
// c_begin > 0 resumes from the capacity lanes s[8..12) that the previous call left in digests[i]
// (c_begin is a multiple of 8, so the rate lanes are overwritten anyway), and a call with
// c_end < n_cols_total leaves them there instead of the digest.  This is what lets a host commit hash
// column group k while group k+1 is still crossing PCIe (etp_batch_from_values_host).
static __global__ void __launch_bounds__(HASH_THREADS, ETP_LEAF_MIN_BLOCKS) hash_leaves_colmajor(const __grid_constant__ LeafSrc src, int c_begin, int c_end,
                                                                     int n_cols_total, uint32_t row0, uint32_t n_rows,
                                                                     uint64_t* __restrict__ digests) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  const size_t row = (size_t)row0 + i;
  uint64_t s[12];
#pragma unroll
  for (int k = 0; k < 12; k++) s[k] = 0;
  if (n_cols_total <= 4) {  // hash_or_noop: copy and zero-pad
#pragma unroll
    for (int c = 0; c < 4; c++)
      if (c < n_cols_total) s[c] = src.base[0][(size_t)c * src.col_stride + row];
    store_digest(digests + 4 * (size_t)i, s);
    return;
  }
  if (c_begin > 0) {
    const ulonglong2* st = reinterpret_cast<const ulonglong2*>(digests + 4 * (size_t)i);
    const ulonglong2 a = st[0], b = st[1];
    s[8] = a.x; s[9] = a.y; s[10] = b.x; s[11] = b.y;
  }
  // one call site for the permutation (instruction-cache footprint); a ragged last chunk overwrites
  // only the first (n_cols - c) rate lanes
  for (int c = c_begin; c < c_end; c += 8) {
    const int si = c / src.cols_per_src;
    const uint64_t* __restrict__ p = src.base[si] + (size_t)(c - si * src.cols_per_src) * src.col_stride + row;
#pragma unroll
    for (int k = 0; k < 8; k++)
      if (c + k < c_end) s[k] = __ldg(p + (size_t)k * src.col_stride);
    poseidon::permute(s);
  }
  if (c_end < n_cols_total) {  // park the capacity lanes (kept as they are: any u64 is a valid lane value)
    ulonglong2 a, b;
    a.x = s[8]; a.y = s[9]; b.x = s[10]; b.y = s[11];
    reinterpret_cast<ulonglong2*>(digests + 4 * (size_t)i)[0] = a;
    reinterpret_cast<ulonglong2*>(digests + 4 * (size_t)i)[1] = b;
    return;
  }
  store_digest(digests + 4 * (size_t)i, s);
}
inline LeafSrc single_src(const uint64_t* base, size_t col_stride, int n_cols) {
  LeafSrc s{};
  s.base[0] = base;
  s.col_stride = col_stride;
  s.cols_per_src = n_cols > 0 ? n_cols : 1;
  return s;
}

// leaves stored row-major (n_leaves x leaf_len): MerkleTree::new on caller-provided rows, FRI layers.
static __global__ void __launch_bounds__(HASH_THREADS) hash_leaves_rowmajor(const uint64_t* __restrict__ rows, int leaf_len,
                                                                     uint32_t n_leaves, uint64_t* __restrict__ digests) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_leaves) return;
  const uint64_t* row = rows + (size_t)i * leaf_len;
  uint64_t s[12];
#pragma unroll
  for (int k = 0; k < 12; k++) s[k] = 0;
  if (leaf_len <= 4) {
#pragma unroll
    for (int c = 0; c < 4; c++)
      if (c < leaf_len) s[c] = row[c];
    store_digest(digests + 4 * (size_t)i, s);
    return;
  }
  for (int c = 0; c < leaf_len; c += 8) {
#pragma unroll
    for (int k = 0; k < 8; k++)
      if (c + k < leaf_len) s[k] = __ldg(row + c + k);
    poseidon::permute(s);
  }
  store_digest(digests + 4 * (size_t)i, s);
}

// parent[j] = two_to_one(child[2j], child[2j+1])
static __global__ void __launch_bounds__(HASH_THREADS) hash_level(const uint64_t* __restrict__ child, uint32_t n_parents,
                                                           uint64_t* __restrict__ parent) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_parents) return;
  const ulonglong2* c = reinterpret_cast<const ulonglong2*>(child + 8 * (size_t)j);
  ulonglong2 a = c[0], b = c[1], d = c[2], e = c[3];
  uint64_t s[12] = {a.x, a.y, b.x, b.y, d.x, d.y, e.x, e.y, 0, 0, 0, 0};
  poseidon::permute(s);
  store_digest(parent + 4 * (size_t)j, s);
}

// ---- narrow levels: one permutation per 16 threads (poseidon_coop.cuh) -----------------------------------
// A level that cannot fill the machine costs one permutation LATENCY whatever its width (33 us with the
// register-resident form); the cooperative form brings that to ~7 us.  Used when n <= COOP_MAX_PARENTS.
constexpr uint32_t COOP_MAX_PARENTS = 8192;
constexpr int COOP_THREADS = 128;  // 8 permutations per CTA

__device__ __forceinline__ void coop_load_rc(uint64_t* rc_s) {
  for (int i = threadIdx.x; i < 360; i += blockDim.x) rc_s[i] = poseidon::RC[i];
  __syncthreads();
}

// parent[j] = two_to_one(child[2j], child[2j+1]), group g of the grid computes parent g
static __global__ void __launch_bounds__(COOP_THREADS) hash_level_coop(const uint64_t* __restrict__ child, uint32_t n_parents,
                                                                uint64_t* __restrict__ parent) {
  __shared__ uint64_t rc_s[360];
  coop_load_rc(rc_s);
  const uint32_t j = (blockIdx.x * blockDim.x + threadIdx.x) / poseidon::COOP_GROUP;
  const int l = threadIdx.x % poseidon::COOP_GROUP;
  const bool live = j < n_parents;  // whole groups are live or not; dead groups still take part in the warp's shuffles
  uint64_t x = (live && l < 8) ? child[8 * (size_t)j + l] : 0;
  x = poseidon::permute_coop(x, l, rc_s);
  if (live && l < 4) parent[4 * (size_t)j + l] = gl::canon(x);
}

// hash_leaves_colmajor for batches with few rows (same arguments and sponge-resume semantics; n_cols_total > 4): one
// 16-thread group per row.  The thread-per-row kernel above needs ceil(C / 8) permutation LATENCIES (33 us each) however few
// rows there are — 0.56 ms for the 135 wire columns of a 2^12-row circuit, 10 ms for a 2400-column keccak table at 2^14 rows —
// here a permutation takes ~7 us.  Throughput is another matter: the shuffle-bound cooperative form sustains ~0.23 G perm/s
// against 1.5 G perm/s (measured: 2^15 rows x 135 columns 2.4 ms cooperative, 0.56 ms thread-per-row), so it pays only while
// rows / 0.23e9 < 33 us, i.e. below ~7.7 k rows — the same crossover as COOP_MAX_PARENTS.
constexpr uint32_t COOP_LEAF_MAX_ROWS = 4096;
static __global__ void __launch_bounds__(COOP_THREADS) hash_leaves_colmajor_coop(const __grid_constant__ LeafSrc src, int c_begin, int c_end,
                                                                          int n_cols_total, uint32_t row0, uint32_t n_rows,
                                                                          uint64_t* __restrict__ digests) {
  __shared__ uint64_t rc_s[360];
  coop_load_rc(rc_s);
  const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) / poseidon::COOP_GROUP;
  const int l = threadIdx.x % poseidon::COOP_GROUP;
  const bool live = i < n_rows;  // whole groups are live or not; dead groups still take part in the warp's shuffles
  const size_t row = (size_t)row0 + i;
  uint64_t x = 0;
  if (live && c_begin > 0 && l >= 8 && l < 12) x = digests[4 * (size_t)i + (l - 8)];  // resume: the parked capacity lanes
  for (int c = c_begin; c < c_end; c += 8) {
    const int cc = c + l;
    if (live && l < 8 && cc < c_end) {
      const int si = cc / src.cols_per_src;
      x = __ldg(src.base[si] + (size_t)(cc - si * src.cols_per_src) * src.col_stride + row);
    }
    x = poseidon::permute_coop(x, l, rc_s);
  }
  if (!live) return;
  if (c_end < n_cols_total) {
    if (l >= 8 && l < 12) digests[4 * (size_t)i + (l - 8)] = x;  // park the capacity lanes (any u64 is a valid lane value)
  } else if (l < 4) {
    digests[4 * (size_t)i + l] = gl::canon(x);
  }
}

// leaves stored row-major (n_leaves x leaf_len), leaf_len > 4: one 16-thread group per leaf, the sponge's
// permutations in sequence (FRI layers: 32 words = 4 permutations)
static __global__ void __launch_bounds__(COOP_THREADS) hash_leaves_rowmajor_coop(const uint64_t* __restrict__ rows, int leaf_len,
                                                                          uint32_t n_leaves, uint64_t* __restrict__ digests) {
  __shared__ uint64_t rc_s[360];
  coop_load_rc(rc_s);
  const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) / poseidon::COOP_GROUP;
  const int l = threadIdx.x % poseidon::COOP_GROUP;
  const bool live = i < n_leaves;
  const uint64_t* row = rows + (size_t)i * leaf_len;
  uint64_t x = 0;
  for (int c0 = 0; c0 < leaf_len; c0 += 8) {
    if (live && l < 8 && c0 + l < leaf_len) x = row[c0 + l];
    x = poseidon::permute_coop(x, l, rc_s);
  }
  if (live && l < 4) digests[4 * (size_t)i + l] = gl::canon(x);
}

// level i (n_nodes nodes) -> plonky2 digest layout. num_layers = log2(n_leaves) - cap_height.
static __global__ void scatter_to_plonky2_layout(const uint64_t* __restrict__ level, int i, uint32_t n_nodes, int num_layers,
                                          uint64_t* __restrict__ digests) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_nodes) return;
  const uint32_t per_sub_log = num_layers - i;  // nodes of this level per cap subtree
  const uint32_t sub = t >> per_sub_log, j = t & ((1u << per_sub_log) - 1);
  const size_t sub_size = ((size_t)1 << (num_layers + 1)) - 2;
  const size_t slot = sub * sub_size + 2 * ((((size_t)j >> 1) << (i + 1)) + ((size_t)1 << i) - 1) + (j & 1);
  const ulonglong2* src = reinterpret_cast<const ulonglong2*>(level + 4 * (size_t)t);
  ulonglong2* dst = reinterpret_cast<ulonglong2*>(digests + 4 * slot);
  dst[0] = src[0];
  dst[1] = src[1];
}

// rows[q][c] = cols[c][idx[q]]
static __global__ void gather_rows_colmajor(const uint64_t* __restrict__ base, size_t col_stride, int n_cols,
                                     const uint64_t* __restrict__ idx, int n_idx, uint64_t* __restrict__ rows) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_idx * n_cols) return;
  const int q = t / n_cols, c = t % n_cols;
  rows[t] = gl::canon(base[(size_t)c * col_stride + idx[q]]);
}

}  // namespace merkle
