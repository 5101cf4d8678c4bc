// libetp_b200: context, transforms, Merkle trees and PolynomialBatch behind the C ABI of
// include/etp_b200.h.  See that header for the upstream plonky2 item each entry point replaces
// (plonky2/src/fri/oracle.rs, hash/merkle_tree.rs, field/src/fft.rs; reached from
// /root/reference/ops/src/lib.rs:52).  Product code: no oracle, no CPU fallback.
#include "ctx.cuh"

// =================================================================================================
// context
// =================================================================================================
extern "C" const char* etp_version(void) { return "etp_b200 0.1 (sm_100a)"; }

extern "C" int etp_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

extern "C" int etp_ctx_create(int device, etp_ctx** out) {
  if (!out) return ETP_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device < 0 || device >= n) return ETP_ERR_CUDA;
  etp_ctx* ctx = new etp_ctx();
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return ETP_ERR_CUDA;
  }
  if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMalloc((void**)&ctx->d_pow_result, 16) != cudaSuccess) {
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return ETP_ERR_CUDA;
  }
  *out = ctx;
  return ETP_OK;
}

// returns every cached (free) block to the runtime; live blocks are untouched
int dev_cache_trim(etp_ctx* ctx) {
  ETP_CUDA(ctx, cudaSetDevice(ctx->device));
  // the streamed host commits also write staging / coefficient blocks from copy_stream
  ETP_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
  ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  std::lock_guard<std::mutex> lock(ctx->cache_mutex);
  for (auto& kv : ctx->cache_free) cudaFree(kv.second);
  ctx->cache_free.clear();
  ctx->cache_free_bytes = 0;
  // full-size twiddle / coset tables (up to 128 MiB each) are rebuilt on demand
  for (auto& kv : ctx->full_tables) cudaFree(kv.second);
  ctx->full_tables.clear();
  return ETP_OK;
}
extern "C" int etp_ctx_trim(etp_ctx* ctx) {
  etp_bind(ctx);
  if (!ctx) return ETP_ERR_INVALID;
  return dev_cache_trim(ctx);
}
extern "C" size_t etp_ctx_cached_bytes(const etp_ctx* ctx) { return ctx ? ctx->cache_free_bytes : 0; }

extern "C" void etp_ctx_destroy(etp_ctx* ctx) {
  etp_bind(ctx);
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->copy_stream);
  cudaStreamSynchronize(ctx->stream);
  free_registered_tables(ctx);
  for (auto& kv : ctx->pow_tables) { cudaFree(kv.second.lo); cudaFree(kv.second.hi); }
  cudaFree(ctx->d_pow_result);
  dev_cache_trim(ctx);  // also frees the full-size tables
  for (auto& kv : ctx->cache_live) cudaFree(kv.first);  // objects the caller never freed
  for (auto e : ctx->sync_events) cudaEventDestroy(e);
  cudaStreamDestroy(ctx->copy_stream);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char* etp_last_error(const etp_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context"; }
extern "C" int etp_ctx_synchronize(etp_ctx* ctx) {
  etp_bind(ctx);
  if (!ctx) return ETP_ERR_INVALID;
  ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ETP_OK;
}
extern "C" void* etp_ctx_stream(etp_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" uint64_t etp_ctx_launch_count(const etp_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int etp_dev_alloc(etp_ctx* ctx, size_t bytes, void** out) {
  etp_bind(ctx);
  if (!ctx || !out) return ETP_ERR_INVALID;
  ETP_CUDA(ctx, cudaSetDevice(ctx->device));
  ETP_CUDA(ctx, cudaMalloc(out, bytes ? bytes : 8));
  return ETP_OK;
}
extern "C" int etp_dev_free(etp_ctx* ctx, void* ptr) {
  etp_bind(ctx);
  if (!ctx) return ETP_ERR_INVALID;
  ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ETP_CUDA(ctx, cudaFree(ptr));
  return ETP_OK;
}
// Page-locks caller-owned host memory (cudaHostRegister) so that the *_host entry points copy from it at full PCIe speed
// and asynchronously; pageable memory works too, but its copies are staged by the driver and block the calling thread.
extern "C" int etp_host_pin(etp_ctx* ctx, void* ptr, size_t bytes) {
  etp_bind(ctx);
  if (!ctx || !ptr || !bytes) return ETP_ERR_INVALID;
  ETP_CUDA(ctx, cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
  return ETP_OK;
}
extern "C" int etp_host_unpin(etp_ctx* ctx, void* ptr) {
  etp_bind(ctx);
  if (!ctx || !ptr) return ETP_ERR_INVALID;
  ETP_CUDA(ctx, cudaHostUnregister(ptr));
  return ETP_OK;
}
extern "C" int etp_dev_upload(etp_ctx* ctx, void* dst, const void* src, size_t bytes) {
  etp_bind(ctx);
  if (!ctx) return ETP_ERR_INVALID;
  ETP_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ETP_OK;
}
extern "C" int etp_dev_download(etp_ctx* ctx, void* dst, const void* src, size_t bytes) {
  etp_bind(ctx);
  if (!ctx) return ETP_ERR_INVALID;
  ETP_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ETP_OK;
}

// =================================================================================================
// power tables
// =================================================================================================
int get_pow_table(etp_ctx* ctx, uint64_t base, int bits, uint64_t scale, ntt::PowTable* out) {
  auto key = std::make_tuple(base, bits, scale);
  auto it = ctx->pow_tables.find(key);
  if (it == ctx->pow_tables.end()) {
    DevPowTable t;
    t.lo_bits = (bits + 1) / 2;
    const size_t n_lo = (size_t)1 << t.lo_bits, n_hi = (size_t)1 << (bits - t.lo_bits);
    std::vector<uint64_t> lo(n_lo), hi(n_hi);
    uint64_t cur = 1;
    for (size_t i = 0; i < n_lo; i++) { lo[i] = gl::canon(cur); cur = gl::mul(cur, base); }
    const uint64_t step = gl::canon(cur);  // base^(2^lo_bits)
    cur = gl::canon(scale);
    for (size_t i = 0; i < n_hi; i++) { hi[i] = gl::canon(cur); cur = gl::mul(cur, step); }
    ETP_CUDA(ctx, cudaMalloc((void**)&t.lo, n_lo * 8));
    ETP_CUDA(ctx, cudaMalloc((void**)&t.hi, n_hi * 8));
    ETP_CUDA(ctx, cudaMemcpyAsync(t.lo, lo.data(), n_lo * 8, cudaMemcpyHostToDevice, ctx->stream));
    ETP_CUDA(ctx, cudaMemcpyAsync(t.hi, hi.data(), n_hi * 8, cudaMemcpyHostToDevice, ctx->stream));
    ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // host vectors die here
    it = ctx->pow_tables.emplace(key, t).first;
  }
  out->lo = it->second.lo;
  out->hi = it->second.hi;
  out->lo_bits = it->second.lo_bits;
  out->mask = (1u << it->second.lo_bits) - 1;
  return ETP_OK;
}

__global__ void k_canon_copy_strided(const uint64_t* src, size_t src_stride, uint64_t* dst, size_t dst_stride, size_t n_cols, int nonzero) {
  const size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (c < n_cols) dst[c * dst_stride] = nonzero ? gl::canon(src[c * src_stride]) : 0;
}

// =================================================================================================
// NTT driver
// =================================================================================================
static int plan_digits(int L, int b[8]) {
  // digits of <= 8 bits, as even as possible, the larger ones last (the last pass wants long contiguous chunks)
  const int m = (L + ntt::MAX_DIGIT_BITS - 1) / ntt::MAX_DIGIT_BITS;
  const int base = L / m, extra = L % m;
  for (int j = 0; j < m; j++) b[j] = base + (j >= m - extra ? 1 : 0);
  return m;
}
int ntt_num_passes(int log_n) { int b[8]; return log_n <= 0 ? 1 : plan_digits(log_n, b); }

template <int B, int MODE>
static void launch_strided(const ntt::PassParams& p, int ul, dim3 grid, cudaStream_t st) {
  ntt::pass_strided<B, MODE><<<grid, ntt::THREADS, ntt::Geo<B>::SMEM_WORDS_STRIDED * 8, st>>>(p, ul);
}
template <int B, int MODE>
static void launch_last(const ntt::PassParams& p, int ul, dim3 grid, cudaStream_t st) {
  ntt::pass_last<B, MODE><<<grid, ntt::THREADS, ntt::Geo<B>::SMEM_WORDS_LAST * 8, st>>>(p, ul);
}
// The passes of the big transforms (digits of 5..8 bits with their tables resident) run the instantiations whose
// switches are compile-time constants; everything else takes the generic kernel of the same digit size.
template <int B>
static void dispatch_digit(bool last, const ntt::PassParams& p, int ul, dim3 grid, cudaStream_t st) {
  if (B >= 5) {
    if (!last && p.tw_full && (!p.in_scale || p.in_full)) {
      if (p.in_scale) launch_strided<B, 1>(p, ul, grid, st); else launch_strided<B, 0>(p, ul, grid, st);
      return;
    }
    if (last && !p.in_scale && !p.natural_out && p.out_scale == 0) { launch_last<B, 0>(p, ul, grid, st); return; }
    if (last && !p.in_scale && p.natural_out && p.out_scale == 2) { launch_last<B, 1>(p, ul, grid, st); return; }
  }
  if (last) launch_last<B, -1>(p, ul, grid, st); else launch_strided<B, -1>(p, ul, grid, st);
}
static void dispatch_pass(bool last, int B, const ntt::PassParams& p, int ul, dim3 grid, cudaStream_t st) {
#define ETP_CASE(n) case n: dispatch_digit<n>(last, p, ul, grid, st); break;
  switch (B) { ETP_CASE(1) ETP_CASE(2) ETP_CASE(3) ETP_CASE(4) ETP_CASE(5) ETP_CASE(6) ETP_CASE(7) ETP_CASE(8) default: break; }
#undef ETP_CASE
}

// cached full-size tables (inter-digit twiddles, coset shift powers); nullptr when too large to keep
static const size_t kMaxFullTable = (size_t)1 << 24;
static int get_full_table(etp_ctx* ctx, const std::tuple<int, uint64_t, int, int, int>& key, size_t n, uint64_t** out) {
  auto it = ctx->full_tables.find(key);
  if (it != ctx->full_tables.end()) { *out = it->second; return ETP_OK; }
  uint64_t* d = nullptr;
  ETP_CUDA(ctx, cudaMalloc((void**)&d, n * 8));
  *out = d;
  return 1;  // caller must fill it, then keep_full_table() (or cudaFree it if the fill launch failed)
}
static int keep_full_table(etp_ctx* ctx, const std::tuple<int, uint64_t, int, int, int>& key, uint64_t* d) {
  ctx->launches++;
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    cudaFree(d);
    return etp_fail(ctx, ETP_ERR_CUDA, "table fill kernel failed: %s", cudaGetErrorString(e));
  }
  ctx->full_tables.emplace(key, d);
  return ETP_OK;
}

int ntt_run(etp_ctx* ctx, const NttArgs& a) {
  if (a.log_n < 0 || a.log_n > 31) return etp_fail(ctx, ETP_ERR_INVALID, "ntt: log_n %d out of range", a.log_n);
  if (a.n_cols == 0) return ETP_OK;
  const int L = a.log_n;
  if (L == 0) {  // size-1 transform: identity (1/n = 1, shift^0 = 1)
    k_canon_copy_strided<<<(unsigned)((a.n_cols + 255) / 256), 256, 0, ctx->stream>>>(a.in, a.in_stride, a.out, a.out_stride, a.n_cols,
                                                                                   a.n_in ? 1 : 0);
    ETP_LAUNCH_CHECK(ctx);
    return ETP_OK;
  }
  int b[8];
  const int m = plan_digits(L, b);
  if (a.natural_out && m > 1 && !a.scratch) return etp_fail(ctx, ETP_ERR_INVALID, "ntt: scratch required");
  ntt::PassParams p{};
  p.log_n = L;
  p.n_cols = (uint32_t)a.n_cols;
  p.inverse = a.inverse ? 1 : 0;
  ETP_TRY(get_pow_table(ctx, gl::root_of_unity(L), L, 1, &p.tw));
  ntt::PowTable in_pow{}, out_pow{};
  const uint64_t n_inv = gl::inv((uint64_t)1 << L);
  uint64_t* in_full = nullptr;
  if (a.coset_shift && !a.inverse) {
    ETP_TRY(get_pow_table(ctx, a.coset_shift, L, 1, &in_pow));
    if (a.n_in <= kMaxFullTable) {
      const auto key = std::make_tuple(-1, a.coset_shift, (int)a.n_in, 0, 0);
      int rc = get_full_table(ctx, key, a.n_in, &in_full);
      if (rc < 0) return rc;
      if (rc == 1) {
        ntt::fill_pow_full<<<(unsigned)((a.n_in + 255) / 256), 256, 0, ctx->stream>>>(in_pow, a.n_in, in_full);
        ETP_TRY(keep_full_table(ctx, key, in_full));
      }
    }
  }
  if (a.coset_shift && a.inverse) ETP_TRY(get_pow_table(ctx, gl::inv(a.coset_shift), L, n_inv, &out_pow));
  int s = L;
  for (int j = 0; j < m; j++) {
    s -= b[j];
    const bool first = j == 0, last = j == m - 1;
    // buffers: forward/bitrev: in -> out, then in place. natural: in -> scratch ... -> out
    const uint64_t* src; size_t src_stride; uint64_t* dst; size_t dst_stride;
    if (first) { src = a.in; src_stride = a.in_stride; }
    else if (a.natural_out) { src = a.scratch; src_stride = a.scratch_stride; }
    else { src = a.out; src_stride = a.out_stride; }
    if (a.natural_out && !last) { dst = a.scratch; dst_stride = a.scratch_stride; }
    else { dst = a.out; dst_stride = a.out_stride; }
    p.in = src; p.in_col_stride = src_stride; p.out = dst; p.out_col_stride = dst_stride;
    p.s = s;
    p.n_in = first ? a.n_in : (1u << L);
    p.in_scale = (first && a.coset_shift && !a.inverse) ? 1 : 0;
    p.in_pow = in_pow;
    p.in_full = in_full;
    p.natural_out = (last && a.natural_out) ? 1 : 0;
    p.out_scale = 0;
    p.tw_full = nullptr;
    if (last && a.inverse) {
      if (a.coset_shift) { p.out_scale = 1; p.out_pow = out_pow; }
      else { p.out_scale = 2; p.out_const = n_inv; }
    }
    const int B = b[j];
    int ul = ntt::TILE_LOG - B;
    if (!last) {
      if (ul > s) ul = s;
      if (((size_t)1 << (s + B)) <= kMaxFullTable) {
        uint64_t* tab = nullptr;
        const auto key = std::make_tuple(L, (uint64_t)0, s, B, p.inverse);
        int rc = get_full_table(ctx, key, (size_t)1 << (s + B), &tab);
        if (rc < 0) return rc;
        if (rc == 1) {
          ntt::fill_pass_twiddles<<<(unsigned)((((size_t)1 << (s + B)) + 255) / 256), 256, 0, ctx->stream>>>(p, B, tab);
          ETP_TRY(keep_full_table(ctx, key, tab));
        }
        p.tw_full = tab;
      }
      dim3 grid((unsigned)(a.n_cols * (((size_t)1 << (L - B)) >> ul)));
      dispatch_pass(false, B, p, ul, grid, ctx->stream);
    } else {
      const int pb = L - B;
      if (ul > pb) ul = pb;
      dim3 grid((unsigned)(a.n_cols * (((size_t)1 << pb) >> ul)));
      dispatch_pass(true, B, p, ul, grid, ctx->stream);
    }
    ETP_LAUNCH_CHECK(ctx);
  }
  return ETP_OK;
}

// one Merkle level: wide levels with the register-resident permutation (throughput), narrow ones with the
// 16-thread cooperative permutation (latency) — merkle.cuh
int launch_hash_level(etp_ctx* ctx, const uint64_t* child, uint32_t parents, uint64_t* parent) {
  if (parents <= merkle::COOP_MAX_PARENTS) {
    const unsigned threads = parents * poseidon::COOP_GROUP;
    merkle::hash_level_coop<<<(threads + merkle::COOP_THREADS - 1) / merkle::COOP_THREADS, merkle::COOP_THREADS, 0, ctx->stream>>>(child, parents, parent);
  } else {
    merkle::hash_level<<<(parents + merkle::HASH_THREADS - 1) / merkle::HASH_THREADS, merkle::HASH_THREADS, 0, ctx->stream>>>(child, parents, parent);
  }
  ETP_LAUNCH_CHECK(ctx);
  return ETP_OK;
}
// row-major leaves (MerkleTree::new on caller rows, FRI layers)
int launch_leaf_hash_rowmajor(etp_ctx* ctx, const uint64_t* rows, int leaf_len, size_t n_leaves, uint64_t* digests) {
  if (leaf_len > 4 && n_leaves <= merkle::COOP_MAX_PARENTS) {
    const unsigned threads = (unsigned)n_leaves * poseidon::COOP_GROUP;
    merkle::hash_leaves_rowmajor_coop<<<(threads + merkle::COOP_THREADS - 1) / merkle::COOP_THREADS, merkle::COOP_THREADS, 0, ctx->stream>>>(
        rows, leaf_len, (uint32_t)n_leaves, digests);
  } else {
    merkle::hash_leaves_rowmajor<<<(unsigned)((n_leaves + merkle::HASH_THREADS - 1) / merkle::HASH_THREADS), merkle::HASH_THREADS, 0, ctx->stream>>>(
        rows, leaf_len, (uint32_t)n_leaves, digests);
  }
  ETP_LAUNCH_CHECK(ctx);
  return ETP_OK;
}

// =================================================================================================
// Merkle driver
// =================================================================================================
int merkle_build_levels(etp_ctx* ctx, uint64_t* levels, size_t n_leaves, int cap_height, uint64_t* cap_host) {
  size_t n = n_leaves;
  uint64_t* cur = levels;
  while (n > ((size_t)1 << cap_height)) {
    uint64_t* nxt = cur + 4 * n;
    const uint32_t parents = (uint32_t)(n >> 1);
    ETP_TRY(launch_hash_level(ctx, cur, parents, nxt));
    cur = nxt;
    n >>= 1;
  }
  ETP_CUDA(ctx, cudaMemcpyAsync(cap_host, cur, ((size_t)32) << cap_height, cudaMemcpyDeviceToHost, ctx->stream));
  ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ETP_OK;
}

int merkle_prove_from_levels(etp_ctx* ctx, const uint64_t* levels, size_t n_leaves, int cap_height, size_t leaf_index,
                             uint64_t* out) {
  const int num_layers = log2_exact(n_leaves) - cap_height;
  if (leaf_index >= n_leaves) return etp_fail(ctx, ETP_ERR_INVALID, "prove: leaf index out of range");
  for (int i = 0; i < num_layers; i++) {
    const size_t node = (leaf_index >> i) ^ 1;
    ETP_CUDA(ctx, cudaMemcpyAsync(out + 4 * i, levels + level_offset(n_leaves, i) + 4 * node, 32, cudaMemcpyDeviceToHost,
                                  ctx->stream));
  }
  ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ETP_OK;
}

int merkle_download_digests(etp_ctx* ctx, const uint64_t* levels, size_t n_leaves, int cap_height, uint64_t* out_host) {
  const int num_layers = log2_exact(n_leaves) - cap_height;
  const size_t nd = 2 * (n_leaves - ((size_t)1 << cap_height));
  if (nd == 0) return ETP_OK;
  DevBuf<uint64_t> tmp(ctx);
  ETP_TRY(tmp.alloc(nd * 4));
  for (int i = 0; i < num_layers; i++) {
    const uint32_t nodes = (uint32_t)(n_leaves >> i);
    merkle::scatter_to_plonky2_layout<<<(nodes + 255) / 256, 256, 0, ctx->stream>>>(levels + level_offset(n_leaves, i), i,
                                                                                    nodes, num_layers, tmp.p);
    ETP_LAUNCH_CHECK(ctx);
  }
  ETP_CUDA(ctx, cudaMemcpyAsync(out_host, tmp.p, nd * 32, cudaMemcpyDeviceToHost, ctx->stream));
  ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ETP_OK;
}

// =================================================================================================
// primitives exposed for parity tests
// =================================================================================================
__global__ void k_permute_states(uint64_t* st, size_t n) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= n) return;
  uint64_t s[12];
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = st[12 * t + i];
  poseidon::permute(s);
#pragma unroll
  for (int i = 0; i < 12; i++) st[12 * t + i] = gl::canon(s[i]);
}

extern "C" int etp_poseidon_permute_host(etp_ctx* ctx, uint64_t* states, size_t n) {
  etp_bind(ctx);
  if (!ctx || (!states && n)) return ETP_ERR_INVALID;
  if (n == 0) return ETP_OK;
  DevBuf<uint64_t> d(ctx);
  ETP_TRY(d.alloc(12 * n));
  ETP_CUDA(ctx, cudaMemcpyAsync(d.p, states, 96 * n, cudaMemcpyHostToDevice, ctx->stream));
  k_permute_states<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(d.p, n);
  ETP_LAUNCH_CHECK(ctx);
  ETP_CUDA(ctx, cudaMemcpyAsync(states, d.p, 96 * n, cudaMemcpyDeviceToHost, ctx->stream));
  ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ETP_OK;
}

static int transform_host(etp_ctx* ctx, uint64_t* cols, size_t n_cols, int log_n, bool inverse, uint64_t shift) {
  if (!ctx) return ETP_ERR_INVALID;
  if (log_n < 0 || log_n > 31 || (!cols && n_cols)) return etp_fail(ctx, ETP_ERR_INVALID, "transform: bad arguments");
  if (n_cols == 0) return ETP_OK;
  const size_t n = (size_t)1 << log_n;
  DevBuf<uint64_t> a(ctx), b(ctx), c(ctx);
  ETP_TRY(a.alloc(n_cols * n));
  ETP_TRY(b.alloc(n_cols * n));
  ETP_TRY(c.alloc(n_cols * n));
  ETP_CUDA(ctx, cudaMemcpyAsync(a.p, cols, n_cols * n * 8, cudaMemcpyHostToDevice, ctx->stream));
  NttArgs args;
  args.in = a.p; args.in_stride = n; args.n_in = (uint32_t)n; args.out = b.p; args.out_stride = n;
  args.scratch = c.p; args.scratch_stride = n; args.log_n = log_n; args.n_cols = n_cols;
  args.inverse = inverse; args.natural_out = true; args.coset_shift = shift;
  ETP_TRY(ntt_run(ctx, args));
  ETP_CUDA(ctx, cudaMemcpyAsync(cols, b.p, n_cols * n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ETP_OK;
}
extern "C" int etp_ifft_host(etp_ctx* ctx, uint64_t* cols, size_t n_cols, int log_n) {
  etp_bind(ctx);
  return transform_host(ctx, cols, n_cols, log_n, true, 0);
}
extern "C" int etp_fft_host(etp_ctx* ctx, uint64_t* cols, size_t n_cols, int log_n) {
  etp_bind(ctx);
  return transform_host(ctx, cols, n_cols, log_n, false, 0);
}
extern "C" int etp_coset_ifft_host(etp_ctx* ctx, uint64_t* cols, size_t n_cols, int log_n, uint64_t shift) {
  etp_bind(ctx);
  if (gl::canon(shift) == 0) return etp_fail(ctx, ETP_ERR_INVALID, "coset shift must be non-zero");
  return transform_host(ctx, cols, n_cols, log_n, true, gl::canon(shift));
}
extern "C" int etp_coset_lde_host(etp_ctx* ctx, const uint64_t* coeffs, size_t n_cols, int log_n, int rate_bits,
                                  uint64_t shift, uint64_t* out) {
  etp_bind(ctx);
  if (!ctx) return ETP_ERR_INVALID;
  if (log_n < 0 || rate_bits < 0 || log_n + rate_bits > 31 || gl::canon(shift) == 0)
    return etp_fail(ctx, ETP_ERR_INVALID, "coset_lde: bad arguments");
  if (n_cols == 0) return ETP_OK;
  const size_t n = (size_t)1 << log_n, big = n << rate_bits;
  DevBuf<uint64_t> a(ctx), b(ctx), c(ctx);
  ETP_TRY(a.alloc(n_cols * n));
  ETP_TRY(b.alloc(n_cols * big));
  ETP_TRY(c.alloc(n_cols * big));
  ETP_CUDA(ctx, cudaMemcpyAsync(a.p, coeffs, n_cols * n * 8, cudaMemcpyHostToDevice, ctx->stream));
  NttArgs args;
  args.in = a.p; args.in_stride = n; args.n_in = (uint32_t)n; args.out = b.p; args.out_stride = big;
  args.scratch = c.p; args.scratch_stride = big; args.log_n = log_n + rate_bits; args.n_cols = n_cols;
  args.natural_out = true; args.coset_shift = gl::canon(shift);
  ETP_TRY(ntt_run(ctx, args));
  ETP_CUDA(ctx, cudaMemcpyAsync(out, b.p, n_cols * big * 8, cudaMemcpyDeviceToHost, ctx->stream));
  ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ETP_OK;
}

// =================================================================================================
// MerkleTree::new
// =================================================================================================
extern "C" int etp_merkle_new_host(etp_ctx* ctx, const uint64_t* leaves, size_t n_leaves, size_t leaf_len, int cap_height,
                                   etp_tree** out) {
  etp_bind(ctx);
  if (!ctx || !out) return ETP_ERR_INVALID;
  *out = nullptr;
  const int lg = log2_exact(n_leaves);
  if (lg < 0) return etp_fail(ctx, ETP_ERR_INVALID, "MerkleTree::new: number of leaves must be a power of two");
  if (cap_height < 0 || cap_height > lg)
    return etp_fail(ctx, ETP_ERR_INVALID, "cap_height=%d should be at most log2(leaves.len())=%d", cap_height, lg);
  if (n_leaves > ((size_t)1 << 31)) return etp_fail(ctx, ETP_ERR_INVALID, "too many leaves");
  etp_tree* t = new etp_tree();
  t->ctx = ctx; t->n_leaves = n_leaves; t->leaf_len = leaf_len; t->cap_height = cap_height;
  t->cap.resize((size_t)4 << cap_height);
  int rc = dev_alloc(ctx, levels_words(n_leaves, cap_height) * 8, (void**)&t->levels);
  if (rc != ETP_OK) { delete t; return rc; }
  DevBuf<uint64_t> d(ctx);
  rc = d.alloc(n_leaves * leaf_len);
  if (rc != ETP_OK) { etp_tree_free(t); return rc; }
  cudaError_t e = cudaMemcpyAsync(d.p, leaves, n_leaves * leaf_len * 8, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) { etp_tree_free(t); return etp_fail(ctx, ETP_ERR_CUDA, "upload failed: %s", cudaGetErrorString(e)); }
  rc = launch_leaf_hash_rowmajor(ctx, d.p, (int)leaf_len, n_leaves, t->levels);
  if (rc != ETP_OK) { etp_tree_free(t); return rc; }
  rc = merkle_build_levels(ctx, t->levels, n_leaves, cap_height, t->cap.data());
  if (rc != ETP_OK) { etp_tree_free(t); return rc; }
  *out = t;
  return ETP_OK;
}
extern "C" void etp_tree_free(etp_tree* t) {
  etp_bind(t ? t->ctx : nullptr);
  if (!t) return;
  dev_free(t->ctx, t->levels);
  delete t;
}
extern "C" size_t etp_tree_num_digests(const etp_tree* t) { return t ? 2 * (t->n_leaves - ((size_t)1 << t->cap_height)) : 0; }
extern "C" int etp_tree_cap(etp_tree* t, uint64_t* cap_out) {
  etp_bind(t ? t->ctx : nullptr);
  if (!t || !cap_out) return ETP_ERR_INVALID;
  memcpy(cap_out, t->cap.data(), t->cap.size() * 8);
  return ETP_OK;
}
extern "C" int etp_tree_digests(etp_tree* t, uint64_t* out) {
  etp_bind(t ? t->ctx : nullptr);
  if (!t || (!out && etp_tree_num_digests(t))) return ETP_ERR_INVALID;
  return merkle_download_digests(t->ctx, t->levels, t->n_leaves, t->cap_height, out);
}
extern "C" int etp_tree_prove(etp_tree* t, size_t leaf_index, uint64_t* out) {
  etp_bind(t ? t->ctx : nullptr);
  if (!t) return ETP_ERR_INVALID;
  return merkle_prove_from_levels(t->ctx, t->levels, t->n_leaves, t->cap_height, leaf_index, out);
}

// =================================================================================================
// PolynomialBatch
// =================================================================================================
int batch_create(etp_ctx* ctx, size_t n_cols, int log_n, int rate_bits, int blinding, int cap_height, etp_batch** out) {
  *out = nullptr;
  if (blinding) return etp_fail(ctx, ETP_ERR_INVALID, "blinding is not supported (the STARK path commits with blinding=false)");
  if (log_n < 0 || rate_bits < 0 || log_n + rate_bits > 31) return etp_fail(ctx, ETP_ERR_INVALID, "bad degree / rate");
  if (cap_height < 0 || cap_height > log_n + rate_bits)
    return etp_fail(ctx, ETP_ERR_INVALID, "cap_height=%d should be at most log2(leaves.len())=%d", cap_height, log_n + rate_bits);
  if (n_cols > 65535) return etp_fail(ctx, ETP_ERR_INVALID, "too many polynomials");
  etp_batch* b = new etp_batch();
  b->ctx = ctx; b->n_cols = n_cols; b->log_n = log_n; b->rate_bits = rate_bits; b->cap_height = cap_height;
  b->cap.resize((size_t)4 << cap_height);
  int rc = dev_alloc(ctx, n_cols * b->n() * 8, (void**)&b->coeffs);
  if (rc == ETP_OK) rc = dev_alloc(ctx, n_cols * b->lde_n() * 8, (void**)&b->lde);
  if (rc == ETP_OK) rc = dev_alloc(ctx, levels_words(b->lde_n(), cap_height) * 8, (void**)&b->levels);
  if (rc != ETP_OK) { etp_batch_free(b); return rc; }
  for (auto& e : b->ev) cudaEventCreate(&e);
  *out = b;
  return ETP_OK;
}

int launch_leaf_hash(etp_ctx* ctx, const merkle::LeafSrc& src, int c_begin, int c_end, int n_cols_total, uint32_t row0,
                     uint32_t n_rows, uint64_t* digests) {
  if (n_rows == 0) return ETP_OK;
  static const bool no_coop = getenv("ETP_NO_COOP_LEAVES") != nullptr;  // A/B switch for tools/: thread-per-row kernel at every size
  if (n_cols_total > 4 && n_rows <= merkle::COOP_LEAF_MAX_ROWS && !no_coop) {
    const unsigned threads = n_rows * poseidon::COOP_GROUP;
    merkle::hash_leaves_colmajor_coop<<<(threads + merkle::COOP_THREADS - 1) / merkle::COOP_THREADS, merkle::COOP_THREADS, 0, ctx->stream>>>(
        src, c_begin, c_end, n_cols_total, row0, n_rows, digests);
  } else {
    merkle::hash_leaves_colmajor<<<(n_rows + merkle::HASH_THREADS - 1) / merkle::HASH_THREADS, merkle::HASH_THREADS, 0, ctx->stream>>>(
        src, c_begin, c_end, n_cols_total, row0, n_rows, digests);
  }
  ETP_LAUNCH_CHECK(ctx);
  return ETP_OK;
}

int get_sync_event(etp_ctx* ctx, size_t i, cudaEvent_t* out) {
  while (ctx->sync_events.size() <= i) {
    cudaEvent_t e;
    ETP_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->sync_events.push_back(e);
  }
  *out = ctx->sync_events[i];
  return ETP_OK;
}

static int batch_finish_levels(etp_batch* b);

int batch_commit_from_coeffs(etp_batch* b) {
  etp_ctx* ctx = b->ctx;
  if (!b->timed_ifft) cudaEventRecord(b->ev[0], ctx->stream);
  cudaEventRecord(b->ev[1], ctx->stream);
  // "FFT + blinding": zero-pad to n << rate_bits, coset_fft(7); bit-reversed order comes for free
  NttArgs args;
  args.in = b->coeffs; args.in_stride = b->n(); args.n_in = (uint32_t)b->n();
  args.out = b->lde; args.out_stride = b->lde_n();
  args.log_n = b->log_n + b->rate_bits; args.n_cols = b->n_cols;
  args.coset_shift = gl::GENERATOR;
  ETP_TRY(ntt_run(ctx, args));
  cudaEventRecord(b->ev[2], ctx->stream);
  // "build Merkle tree"
  ETP_TRY(launch_leaf_hash(ctx, merkle::single_src(b->lde, b->lde_n(), (int)b->n_cols), 0, (int)b->n_cols, (int)b->n_cols, 0,
                           (uint32_t)b->lde_n(), b->levels));
  cudaEventRecord(b->ev[3], ctx->stream);
  b->timed_ifft = false;
  return batch_finish_levels(b);
}

// inner levels + cap of a batch whose leaf digests are in levels[0]
static int batch_finish_levels(etp_batch* b) {
  etp_ctx* ctx = b->ctx;
  size_t n = b->lde_n();
  uint64_t* cur = b->levels;
  while (n > ((size_t)1 << b->cap_height)) {
    uint64_t* nxt = cur + 4 * n;
    const uint32_t parents = (uint32_t)(n >> 1);
    ETP_TRY(launch_hash_level(ctx, cur, parents, nxt));
    cur = nxt;
    n >>= 1;
  }
  cudaEventRecord(b->ev[4], ctx->stream);
  ETP_CUDA(ctx, cudaMemcpyAsync(b->cap.data(), cur, ((size_t)32) << b->cap_height, cudaMemcpyDeviceToHost, ctx->stream));
  ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ETP_OK;
}

int batch_commit_from_values(etp_batch* b, const uint64_t* values_dev, size_t col_stride) {
  // "IFFT": natural -> natural; the LDE buffer doubles as scratch for the multi-pass transform
  cudaEventRecord(b->ev[0], b->ctx->stream);
  b->timed_ifft = true;
  NttArgs args;
  args.in = values_dev; args.in_stride = col_stride; args.n_in = (uint32_t)b->n();
  args.out = b->coeffs; args.out_stride = b->n();
  args.scratch = b->lde; args.scratch_stride = b->lde_n();
  args.log_n = b->log_n; args.n_cols = b->n_cols; args.inverse = true; args.natural_out = true;
  ETP_TRY(ntt_run(b->ctx, args));
  return batch_commit_from_coeffs(b);
}

// ---- streamed host commit -----------------------------------------------------------------------------
// Columns are independent until the leaf sponge, and the sponge absorbs them in order, 8 at a time: so the
// batch is cut into groups of G columns (G a multiple of 8) and group k runs iFFT -> coset LDE -> sponge
// absorb on the compute stream while group k+1 is copied on the copy stream.  Only 2 G n words of staging
// are needed (values) or none (coefficients land in b->coeffs directly).
static size_t stream_group_cols(const etp_batch* b) {
  const size_t bytes = b->n_cols * b->n() * 8;
  if (b->n_cols <= 8 || bytes < ((size_t)32 << 20)) return b->n_cols ? b->n_cols : 1;  // too small to pipeline
  size_t target = 8;                // aim at 8 groups
  if (const char* e = getenv("ETP_STREAM_GROUPS")) target = (size_t)atoi(e) > 0 ? (size_t)atoi(e) : 8;
  size_t g = (b->n_cols + target - 1) / target;
  g = (g + 7) / 8 * 8;
  return g;
}

__global__ void k_canon_copy(const uint64_t* src, size_t src_stride, uint64_t* dst, size_t n, size_t n_cols);

// keep_values (device, n_cols x n, values only): the columns land there and stay (a prover needs the trace itself for its
// helper columns) instead of passing through the two staging slots.
static int batch_commit_from_host_streamed_impl(etp_batch* b, const uint64_t* const* cols, bool is_values, uint64_t* keep_values);
int batch_commit_from_host_streamed(etp_batch* b, const uint64_t* const* cols, bool is_values, uint64_t* keep_values) {
  const int rc = batch_commit_from_host_streamed_impl(b, cols, is_values, keep_values);
  if (rc != ETP_OK) {
    // H2D copies of later groups may still be in flight on copy_stream while the caller hands the staging /
    // coefficient blocks back to the cache: drain both streams first
    cudaStreamSynchronize(b->ctx->copy_stream);
    cudaStreamSynchronize(b->ctx->stream);
    cudaGetLastError();
  }
  return rc;
}
static int batch_commit_from_host_streamed_impl(etp_batch* b, const uint64_t* const* cols, bool is_values, uint64_t* keep_values) {
  etp_ctx* ctx = b->ctx;
  const size_t n = b->n(), C = b->n_cols, G = stream_group_cols(b);
  // group boundaries (multiples of 8 columns).  The first group is a single sponge chunk: its copy is the only one
  // nothing can hide.  A last group of fewer than 8 columns joins its predecessor: its ragged sponge chunk keeps
  // rate lanes of the previous permutation, and only the capacity lanes are carried from one launch to the next.
  // The pipeline ramps up: a group's copy (0.6 ms per 2^22-row column) must not outlast the previous group's compute
  // (1 ms per column), so the first two groups are single sponge chunks and only then come groups of G columns.
  std::vector<size_t> bounds{0};
  if (G < C && G > 8) {
    bounds.push_back(8);
    if (C > 24) bounds.push_back(16);
  }
  while (bounds.back() < C) bounds.push_back(bounds.back() + G < C ? bounds.back() + G : C);
  if (bounds.size() > 2 && C - bounds[bounds.size() - 2] < 8) bounds.erase(bounds.end() - 2);
  const size_t n_groups = bounds.size() - 1;
  for (size_t c = 0; c < C; c++)
    if (!cols[c]) return etp_fail(ctx, ETP_ERR_INVALID, "null column %zu", c);
  DevBuf<uint64_t> stage(ctx);
  const size_t slot = (G + 8) * n;
  const bool staged = is_values && !keep_values;
  if (staged) ETP_TRY(stage.alloc(2 * slot));
  cudaEvent_t ready;
  ETP_TRY(get_sync_event(ctx, 0, &ready));
  cudaEventRecord(b->ev[0], ctx->stream);
  cudaEventRecord(b->ev[1], ctx->stream);
  cudaEventRecord(b->ev[2], ctx->stream);
  ETP_CUDA(ctx, cudaEventRecord(ready, ctx->stream));  // buffers allocated (stream-ordered) before the copies start
  ETP_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ready, 0));
  for (size_t k = 0; k < n_groups; k++) {
    const size_t c0 = bounds[k], gc = bounds[k + 1] - c0;
    uint64_t* land = staged ? stage.p + (k & 1) * slot : (is_values ? keep_values + c0 * n : b->coeffs + c0 * n);
    cudaEvent_t h2d_done, slot_free;
    ETP_TRY(get_sync_event(ctx, 1 + 2 * k, &h2d_done));
    ETP_TRY(get_sync_event(ctx, 2 + 2 * k, &slot_free));
    if (staged && k >= 2) {  // the staging slot is free once the iFFT of group k-2 has consumed it
      cudaEvent_t prev;
      ETP_TRY(get_sync_event(ctx, 2 + 2 * (k - 2), &prev));
      ETP_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, prev, 0));
    }
    // transforms of columns [t0, t0 + tc) of this group, once `arrived` (recorded on the copy stream) has fired
    auto transform = [&](size_t t0, size_t tc, cudaEvent_t arrived, bool last_of_group) -> int {
      ETP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, arrived, 0));
      uint64_t* src = land + (t0 - c0) * n;
      NttArgs args;
      if (is_values) {  // "IFFT": natural -> natural; this group's slice of the LDE buffer is the scratch
        args.in = src; args.in_stride = n; args.n_in = (uint32_t)n;
        args.out = b->coeffs + t0 * n; args.out_stride = n;
        args.scratch = b->lde + t0 * b->lde_n(); args.scratch_stride = b->lde_n();
        args.log_n = b->log_n; args.n_cols = tc; args.inverse = true; args.natural_out = true;
        ETP_TRY(ntt_run(ctx, args));
        if (last_of_group) ETP_CUDA(ctx, cudaEventRecord(slot_free, ctx->stream));
      } else {
        k_canon_copy<<<(unsigned)((tc * n + 255) / 256), 256, 0, ctx->stream>>>(src, n, src, n, tc);
        ETP_LAUNCH_CHECK(ctx);
      }
      args = NttArgs();  // "FFT + blinding"
      args.in = b->coeffs + t0 * n; args.in_stride = n; args.n_in = (uint32_t)n;
      args.out = b->lde + t0 * b->lde_n(); args.out_stride = b->lde_n();
      args.log_n = b->log_n + b->rate_bits; args.n_cols = tc;
      args.coset_shift = gl::GENERATOR;
      return ntt_run(ctx, args);
    };
    if (k == 0 && n_groups > 1) {
      // nothing hides the first group's copy, so its columns are transformed one by one as they land: the exposed
      // time is one column's copy instead of eight
      for (size_t c = 0; c < gc; c++) {
        cudaEvent_t col_done;
        ETP_TRY(get_sync_event(ctx, 1 + 2 * n_groups + c, &col_done));
        ETP_CUDA(ctx, cudaMemcpyAsync(land + c * n, cols[c0 + c], n * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
        ETP_CUDA(ctx, cudaEventRecord(col_done, ctx->copy_stream));
        ETP_TRY(transform(c0 + c, 1, col_done, c + 1 == gc));
      }
    } else {
      for (size_t c = 0; c < gc; c++)
        ETP_CUDA(ctx, cudaMemcpyAsync(land + c * n, cols[c0 + c], n * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
      ETP_CUDA(ctx, cudaEventRecord(h2d_done, ctx->copy_stream));
      ETP_TRY(transform(c0, gc, h2d_done, true));
    }
    ETP_TRY(launch_leaf_hash(ctx, merkle::single_src(b->lde, b->lde_n(), (int)C), (int)c0, (int)(c0 + gc), (int)C, 0,
                             (uint32_t)b->lde_n(), b->levels));
  }
  if (C == 0)
    ETP_TRY(launch_leaf_hash(ctx, merkle::single_src(b->lde, b->lde_n(), 0), 0, 0, 0, 0, (uint32_t)b->lde_n(), b->levels));
  cudaEventRecord(b->ev[3], ctx->stream);
  b->timed_ifft = false;
  return batch_finish_levels(b);
}

__global__ void k_canon_copy(const uint64_t* src, size_t src_stride, uint64_t* dst, size_t n, size_t n_cols) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= n * n_cols) return;
  const size_t c = t / n, i = t % n;
  dst[t] = gl::canon(src[c * src_stride + i]);
}

extern "C" int etp_batch_from_values_host(etp_ctx* ctx, const uint64_t* const* cols, size_t n_cols, int log_n, int rate_bits,
                                          int blinding, int cap_height, etp_batch** out) {
  etp_bind(ctx);
  if (!ctx || !out || (!cols && n_cols)) return ETP_ERR_INVALID;
  etp_batch* b;
  ETP_TRY(batch_create(ctx, n_cols, log_n, rate_bits, blinding, cap_height, &b));
  int rc = batch_commit_from_host_streamed(b, cols, true, nullptr);
  if (rc != ETP_OK) { etp_batch_free(b); return rc; }
  *out = b;
  return ETP_OK;
}

extern "C" int etp_batch_from_coeffs_host(etp_ctx* ctx, const uint64_t* const* cols, size_t n_cols, int log_n, int rate_bits,
                                          int blinding, int cap_height, etp_batch** out) {
  etp_bind(ctx);
  if (!ctx || !out || (!cols && n_cols)) return ETP_ERR_INVALID;
  etp_batch* b;
  ETP_TRY(batch_create(ctx, n_cols, log_n, rate_bits, blinding, cap_height, &b));
  int rc = batch_commit_from_host_streamed(b, cols, false, nullptr);
  if (rc != ETP_OK) { etp_batch_free(b); return rc; }
  *out = b;
  return ETP_OK;
}

extern "C" int etp_batch_from_values_dev(etp_ctx* ctx, const uint64_t* values_dev, size_t col_stride, size_t n_cols, int log_n,
                                         int rate_bits, int blinding, int cap_height, etp_batch** out) {
  etp_bind(ctx);
  if (!ctx || !out || (!values_dev && n_cols)) return ETP_ERR_INVALID;
  etp_batch* b;
  ETP_TRY(batch_create(ctx, n_cols, log_n, rate_bits, blinding, cap_height, &b));
  int rc = batch_commit_from_values(b, values_dev, col_stride);
  if (rc != ETP_OK) { etp_batch_free(b); return rc; }
  *out = b;
  return ETP_OK;
}

extern "C" int etp_batch_from_coeffs_dev(etp_ctx* ctx, const uint64_t* coeffs_dev, size_t col_stride, size_t n_cols, int log_n,
                                         int rate_bits, int blinding, int cap_height, etp_batch** out) {
  etp_bind(ctx);
  if (!ctx || !out || (!coeffs_dev && n_cols)) return ETP_ERR_INVALID;
  etp_batch* b;
  ETP_TRY(batch_create(ctx, n_cols, log_n, rate_bits, blinding, cap_height, &b));
  int rc = ETP_OK;
  const size_t tot = n_cols * b->n();
  if (tot) {
    k_canon_copy<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(coeffs_dev, col_stride, b->coeffs, b->n(), n_cols);
    ctx->launches++;
  }
  if (rc == ETP_OK) rc = batch_commit_from_coeffs(b);
  if (rc != ETP_OK) { etp_batch_free(b); return rc; }
  *out = b;
  return ETP_OK;
}

extern "C" int etp_batch_recommit_values_dev(etp_batch* b, const uint64_t* values_dev, size_t col_stride) {
  etp_bind(b ? b->ctx : nullptr);
  if (!b || !values_dev) return ETP_ERR_INVALID;
  return batch_commit_from_values(b, values_dev, col_stride);
}

extern "C" int etp_batch_last_commit_timings(const etp_batch* b, float ms_out[4]) {
  if (!b || !ms_out) return ETP_ERR_INVALID;
  for (int i = 0; i < 4; i++)
    if (cudaEventElapsedTime(&ms_out[i], b->ev[i], b->ev[i + 1]) != cudaSuccess) return ETP_ERR_STATE;
  return ETP_OK;
}

extern "C" void etp_batch_free(etp_batch* b) {
  etp_bind(b ? b->ctx : nullptr);
  if (!b) return;
  for (auto& e : b->ev)
    if (e) cudaEventDestroy(e);
  dev_free(b->ctx, b->coeffs);
  dev_free(b->ctx, b->lde);
  dev_free(b->ctx, b->levels);
  delete b;
}
extern "C" size_t etp_batch_num_cols(const etp_batch* b) { return b ? b->n_cols : 0; }
extern "C" int etp_batch_degree_log(const etp_batch* b) { return b ? b->log_n : -1; }
extern "C" size_t etp_batch_num_digests(const etp_batch* b) { return b ? 2 * (b->lde_n() - ((size_t)1 << b->cap_height)) : 0; }
extern "C" int etp_batch_cap(etp_batch* b, uint64_t* cap_out) {
  etp_bind(b ? b->ctx : nullptr);
  if (!b || !cap_out) return ETP_ERR_INVALID;
  memcpy(cap_out, b->cap.data(), b->cap.size() * 8);
  return ETP_OK;
}
extern "C" int etp_batch_download_coeffs(etp_batch* b, uint64_t* out) {
  etp_bind(b ? b->ctx : nullptr);
  if (!b || (!out && b->n_cols)) return ETP_ERR_INVALID;
  if (b->n_cols == 0) return ETP_OK;
  ETP_CUDA(b->ctx, cudaMemcpyAsync(out, b->coeffs, b->n_cols * b->n() * 8, cudaMemcpyDeviceToHost, b->ctx->stream));
  ETP_CUDA(b->ctx, cudaStreamSynchronize(b->ctx->stream));
  return ETP_OK;
}

// rows[i][c] = lde[c][i]  (tiled transpose through shared memory)
__global__ void k_transpose_to_rows(const uint64_t* __restrict__ lde, size_t col_stride, int n_cols, size_t n_rows,
                                    uint64_t* __restrict__ rows) {
  __shared__ uint64_t tile[32][33];
  const size_t r0 = (size_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int k = threadIdx.y; k < 32; k += blockDim.y) {
    const int c = c0 + k;
    const size_t r = r0 + threadIdx.x;
    if (c < n_cols && r < n_rows) tile[k][threadIdx.x] = lde[(size_t)c * col_stride + r];
  }
  __syncthreads();
  for (int k = threadIdx.y; k < 32; k += blockDim.y) {
    const size_t r = r0 + k;
    const int c = c0 + threadIdx.x;
    if (c < n_cols && r < n_rows) rows[r * n_cols + c] = gl::canon(tile[threadIdx.x][k]);
  }
}

extern "C" int etp_batch_download_leaves(etp_batch* b, uint64_t* out) {
  etp_bind(b ? b->ctx : nullptr);
  if (!b || (!out && b->n_cols)) return ETP_ERR_INVALID;
  if (b->n_cols == 0) return ETP_OK;
  etp_ctx* ctx = b->ctx;
  DevBuf<uint64_t> rows(ctx);
  ETP_TRY(rows.alloc(b->n_cols * b->lde_n()));
  dim3 grid((unsigned)((b->lde_n() + 31) / 32), (unsigned)((b->n_cols + 31) / 32)), block(32, 8);
  k_transpose_to_rows<<<grid, block, 0, ctx->stream>>>(b->lde, b->lde_n(), (int)b->n_cols, b->lde_n(), rows.p);
  ETP_LAUNCH_CHECK(ctx);
  ETP_CUDA(ctx, cudaMemcpyAsync(out, rows.p, b->n_cols * b->lde_n() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ETP_OK;
}
extern "C" int etp_batch_download_digests(etp_batch* b, uint64_t* out) {
  etp_bind(b ? b->ctx : nullptr);
  if (!b) return ETP_ERR_INVALID;
  return merkle_download_digests(b->ctx, b->levels, b->lde_n(), b->cap_height, out);
}
extern "C" int etp_batch_leaves_at(etp_batch* b, const uint64_t* idx, size_t n_idx, uint64_t* rows_out) {
  etp_bind(b ? b->ctx : nullptr);
  if (!b || (!idx && n_idx)) return ETP_ERR_INVALID;
  if (n_idx == 0 || b->n_cols == 0) return ETP_OK;
  etp_ctx* ctx = b->ctx;
  for (size_t q = 0; q < n_idx; q++)
    if (idx[q] >= b->lde_n()) return etp_fail(ctx, ETP_ERR_INVALID, "leaf index out of range");
  DevBuf<uint64_t> d_idx(ctx), d_rows(ctx);
  ETP_TRY(d_idx.alloc(n_idx));
  ETP_TRY(d_rows.alloc(n_idx * b->n_cols));
  ETP_CUDA(ctx, cudaMemcpyAsync(d_idx.p, idx, n_idx * 8, cudaMemcpyHostToDevice, ctx->stream));
  const size_t tot = n_idx * b->n_cols;
  merkle::gather_rows_colmajor<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(b->lde, b->lde_n(), (int)b->n_cols, d_idx.p,
                                                                                      (int)n_idx, d_rows.p);
  ETP_LAUNCH_CHECK(ctx);
  ETP_CUDA(ctx, cudaMemcpyAsync(rows_out, d_rows.p, tot * 8, cudaMemcpyDeviceToHost, ctx->stream));
  ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ETP_OK;
}
extern "C" int etp_batch_get_lde_values(etp_batch* b, size_t index, size_t step, uint64_t* row_out) {
  etp_bind(b ? b->ctx : nullptr);
  if (!b) return ETP_ERR_INVALID;
  const size_t i = index * step;
  if (i >= b->lde_n()) return etp_fail(b->ctx, ETP_ERR_INVALID, "get_lde_values: index out of range");
  uint64_t pos = gl::bitrev32((uint32_t)i, b->log_n + b->rate_bits);
  return etp_batch_leaves_at(b, &pos, 1, row_out);
}
extern "C" int etp_batch_prove(etp_batch* b, size_t leaf_index, uint64_t* out) {
  etp_bind(b ? b->ctx : nullptr);
  if (!b) return ETP_ERR_INVALID;
  return merkle_prove_from_levels(b->ctx, b->levels, b->lde_n(), b->cap_height, leaf_index, out);
}
extern "C" const uint64_t* etp_batch_lde_dev(const etp_batch* b, size_t* col_stride) {
  if (!b) return nullptr;
  if (col_stride) *col_stride = b->lde_n();
  return b->lde;
}
extern "C" const uint64_t* etp_batch_coeffs_dev(const etp_batch* b, size_t* col_stride) {
  if (!b) return nullptr;
  if (col_stride) *col_stride = b->n();
  return b->coeffs;
}

// =================================================================================================
// pipe-rate micro-benchmark (bench.py's integer-pipe roofline denominator)
// =================================================================================================
// Each kernel issues PIPE_ITERS x PIPE_CHAINS instructions of one kind per thread on independent dependency chains
// (8 chains x 8 warps per SM sub-partition hide the pipe latency), so elapsed time / instruction count is the
// sustained issue rate of that pipe.  Timed with CUDA events on the context's stream.
namespace pipes {
constexpr int ITERS = 2048, CHAINS = 8, THREADS = 256, BLOCKS_PER_SM = 4;
__global__ void __launch_bounds__(THREADS) k_imad_wide_zero(uint64_t* out, uint32_t b) {  // IMAD.WIDE.U32 d = a*b + RZ
  uint64_t acc[CHAINS];
  for (int i = 0; i < CHAINS; i++) acc[i] = i + threadIdx.x;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(acc[i]) : "r"((uint32_t)acc[i]), "r"(b));
  }
  uint64_t s = 0;
  for (int i = 0; i < CHAINS; i++) s += acc[i];
  out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(THREADS) k_imad_wide_acc(uint64_t* out, uint32_t b) {  // IMAD.WIDE.U32 d = a*b + d (64-bit addend)
  uint64_t acc[CHAINS];
  for (int i = 0; i < CHAINS; i++) acc[i] = i + threadIdx.x;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"((uint32_t)acc[(i + 1) % CHAINS]), "r"(b));
  }
  uint64_t s = 0;
  for (int i = 0; i < CHAINS; i++) s += acc[i];
  out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(THREADS) k_dfma(uint64_t* out, uint32_t b) {  // FP64 pipe (shares the ALU issue port)
  double acc[CHAINS];
  const double m = 1.0 + 1e-9 * b, c = 1e-3;
  for (int i = 0; i < CHAINS; i++) acc[i] = i + threadIdx.x;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(acc[i]) : "d"(m), "d"(c));
  }
  double s = 0;
  for (int i = 0; i < CHAINS; i++) s += acc[i];
  out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = (uint64_t)s;
}
}  // namespace pipes

// rates_out[3]: sustained thread-instructions per second of the whole device for
// [0] IMAD.WIDE.U32 with a zero addend, [1] mad.wide.u32 accumulating into a 64-bit addend (ptxas emits IMAD.WIDE + a 64-bit
// IADD3 pair for it), [2] DFMA.  (A plain integer-add chain is not reported: ptxas rewrites it — IADD3 fusion, IMAD moves —
// so its "rate" says nothing about the ALU port.)
extern "C" int etp_bench_pipe_rates(etp_ctx* ctx, double rates_out[3]) {
  etp_bind(ctx);
  if (!ctx || !rates_out) return ETP_ERR_INVALID;
  cudaDeviceProp prop;
  ETP_CUDA(ctx, cudaGetDeviceProperties(&prop, ctx->device));
  const unsigned blocks = (unsigned)prop.multiProcessorCount * pipes::BLOCKS_PER_SM;
  DevBuf<uint64_t> out(ctx);
  ETP_TRY(out.alloc((size_t)blocks * pipes::THREADS));
  cudaEvent_t e0, e1;
  ETP_CUDA(ctx, cudaEventCreate(&e0));
  ETP_CUDA(ctx, cudaEventCreate(&e1));
  typedef void (*kern_t)(uint64_t*, uint32_t);
  const kern_t kerns[3] = {pipes::k_imad_wide_zero, pipes::k_imad_wide_acc, pipes::k_dfma};
  int rc = ETP_OK;
  for (int k = 0; k < 3 && rc == ETP_OK; k++) {
    float best = 0;
    for (int rep = 0; rep < 4; rep++) {  // first repetition warms up
      cudaEventRecord(e0, ctx->stream);
      kerns[k]<<<blocks, pipes::THREADS, 0, ctx->stream>>>(out.p, 3u);
      ctx->launches++;
      cudaEventRecord(e1, ctx->stream);
      if (cudaEventSynchronize(e1) != cudaSuccess || cudaGetLastError() != cudaSuccess) { rc = etp_fail(ctx, ETP_ERR_CUDA, "pipe benchmark kernel failed"); break; }
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep && (best == 0 || ms < best)) best = ms;
    }
    const double instr = (double)blocks * pipes::THREADS * pipes::ITERS * pipes::CHAINS;
    rates_out[k] = best > 0 ? instr / (best * 1e-3) : 0;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return rc;
}
