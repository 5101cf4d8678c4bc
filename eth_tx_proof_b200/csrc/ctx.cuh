// Context, error plumbing, power tables and the NTT / Merkle host drivers shared by the C ABI files.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "../../include/etp_b200.h"
#include "gl.cuh"
#include "merkle.cuh"
#include "ntt.cuh"

struct DevPowTable {
  uint64_t* lo = nullptr;
  uint64_t* hi = nullptr;
  int lo_bits = 0;
};

// a kernel compiled at run time from a constraint program (etp_jit.cu)
struct JitKernel {
  void* library = nullptr;  // cudaLibrary_t
  void* kernel = nullptr;   // cudaKernel_t
};
struct RegisteredTable;  // etp_stark.cu

struct etp_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // H2D of column groups, overlapped with the transforms / hashing on `stream`
  std::vector<cudaEvent_t> sync_events;  // timing-disabled events reused by the streamed host commits
  std::string err;
  uint64_t launches = 0;
  // (base, bits, scale) -> device tables
  std::map<std::tuple<uint64_t, int, uint64_t>, DevPowTable> pow_tables;
  // full-size tables: (log_n | -1, shift, s | n_in, B, inverse) -> device array
  std::map<std::tuple<int, uint64_t, int, int, int>, uint64_t*> full_tables;
  // last prove timings
  std::vector<std::pair<const char*, float>> timings;
  uint64_t* d_pow_result = nullptr;  // PoW grind result slot
  std::vector<RegisteredTable*> tables;  // program-defined tables, id = ETP_TABLE_FIRST_REGISTERED + index
  // Block cache behind dev_alloc / dev_free.  Everything a context does is ordered on `stream`, so a block released
  // by dev_free may be handed out again at once: its next use is enqueued behind its last one.  Proofs and commits of
  // a shape seen before therefore allocate nothing (measured: with cudaMallocAsync pools single proofs at 2^22 rows
  // jittered between 55 and 900 ms).  Blocks are matched by size (<= 12.5 % slack); on out-of-memory the cache is
  // emptied and the allocation retried; etp_ctx_trim() empties it on request.
  std::mutex cache_mutex;  // objects may be released from another thread (a garbage collector) than the one proving
  std::multimap<size_t, void*> cache_free;
  std::unordered_map<void*, size_t> cache_live;
  size_t cache_free_bytes = 0;
};
int dev_cache_trim(etp_ctx* ctx);
void free_registered_tables(etp_ctx* ctx);  // etp_stark.cu
int jit_compile(etp_ctx* ctx, const std::string& source, std::vector<char>* cubin_out, std::string* log_out);
int jit_load(etp_ctx* ctx, const std::vector<char>& cubin, const char* entry, JitKernel* out);
void jit_unload(JitKernel* k);

// Every C-ABI entry point binds the calling thread to its context's device first: callers are arbitrary host threads
// (tokio workers in the reference, parallel.ProverPool here) whose current device is whatever the runtime defaults to.
inline void etp_bind(const etp_ctx* ctx) {
  if (ctx) cudaSetDevice(ctx->device);
}

inline int etp_fail(etp_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  return code;
}

#define ETP_CUDA(ctx, call)                                                                                  \
  do {                                                                                                       \
    cudaError_t e__ = (call);                                                                                \
    if (e__ != cudaSuccess) {                                                                                \
      cudaGetLastError(); /* reported here: must not resurface in a later launch check */                   \
      return etp_fail((ctx), ETP_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                      __LINE__);                                                                             \
    }                                                                                                        \
  } while (0)
#define ETP_TRY(expr)            \
  do {                           \
    int rc__ = (expr);           \
    if (rc__ != ETP_OK) return rc__; \
  } while (0)
#define ETP_LAUNCH_CHECK(ctx)    \
  do {                           \
    (ctx)->launches++;           \
    ETP_CUDA((ctx), cudaGetLastError()); \
  } while (0)

inline int dev_alloc(etp_ctx* ctx, size_t bytes, void** out) {
  bytes = (bytes + 511) & ~(size_t)511;
  if (bytes == 0) bytes = 512;
  std::unique_lock<std::mutex> lock(ctx->cache_mutex);
  auto it = ctx->cache_free.lower_bound(bytes);
  if (it != ctx->cache_free.end() && it->first <= bytes + bytes / 8) {
    *out = it->second;
    ctx->cache_live[*out] = it->first;
    ctx->cache_free_bytes -= it->first;
    ctx->cache_free.erase(it);
    return ETP_OK;
  }
  ETP_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaError_t e = cudaMalloc(out, bytes);
  if (e == cudaErrorMemoryAllocation) {  // give the cached blocks back and try once more
    cudaGetLastError();
    lock.unlock();
    ETP_TRY(dev_cache_trim(ctx));
    lock.lock();
    e = cudaMalloc(out, bytes);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    return etp_fail(ctx, ETP_ERR_CUDA, "cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
  }
  ctx->cache_live[*out] = bytes;
  return ETP_OK;
}
inline void dev_free(etp_ctx* ctx, void* p) {
  if (!p) return;
  std::lock_guard<std::mutex> lock(ctx->cache_mutex);
  auto it = ctx->cache_live.find(p);
  if (it == ctx->cache_live.end()) {  // not ours (never happens for dev_alloc'ed memory): hand it to the runtime
    cudaFree(p);
    return;
  }
  ctx->cache_free.emplace(it->second, p);
  ctx->cache_free_bytes += it->second;
  ctx->cache_live.erase(it);
}
template <class T>
struct DevBuf {  // RAII scratch on the context's stream-ordered pool
  etp_ctx* ctx;
  T* p = nullptr;
  explicit DevBuf(etp_ctx* c) : ctx(c) {}
  int alloc(size_t n) { return dev_alloc(ctx, n * sizeof(T), (void**)&p); }
  ~DevBuf() { dev_free(ctx, p); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};

// ---- power tables --------------------------------------------------------------------------------
// lookup(e) = scale * base^e for e < 2^bits
int get_pow_table(etp_ctx* ctx, uint64_t base, int bits, uint64_t scale, ntt::PowTable* out);

// ---- NTT driver ------------------------------------------------------------------------------------
struct NttArgs {
  const uint64_t* in = nullptr;
  size_t in_stride = 0;
  uint32_t n_in = 0;           // number of (non-zero) inputs, <= 2^log_n
  uint64_t* out = nullptr;
  size_t out_stride = 0;
  uint64_t* scratch = nullptr; // needed when natural_out and more than one pass (size n_cols x 2^log_n)
  size_t scratch_stride = 0;
  int log_n = 0;
  size_t n_cols = 0;
  bool inverse = false;
  bool natural_out = false;    // false: position p holds out[bitrev(p)]
  uint64_t coset_shift = 0;    // 0: plain; forward: in[j] *= shift^j; inverse: out[k] *= shift^-k
};
int ntt_num_passes(int log_n);
int ntt_run(etp_ctx* ctx, const NttArgs& a);

// ---- Merkle driver -----------------------------------------------------------------------------------
inline size_t levels_words(size_t n_leaves, int cap_height) {
  size_t w = 0;
  for (size_t n = n_leaves; n >= ((size_t)1 << cap_height); n >>= 1) { w += 4 * n; if (n == 1) break; }
  return w;
}
inline size_t level_offset(size_t n_leaves, int level) {
  size_t w = 0;
  for (int i = 0; i < level; i++) w += 4 * (n_leaves >> i);
  return w;
}
// levels[0] must already hold the leaf digests; hashes up to the cap level and copies the cap to host
int merkle_build_levels(etp_ctx* ctx, uint64_t* levels, size_t n_leaves, int cap_height, uint64_t* cap_host);
int launch_hash_level(etp_ctx* ctx, const uint64_t* child, uint32_t parents, uint64_t* parent);
int launch_leaf_hash_rowmajor(etp_ctx* ctx, const uint64_t* rows, int leaf_len, size_t n_leaves, uint64_t* digests);
int merkle_prove_from_levels(etp_ctx* ctx, const uint64_t* levels, size_t n_leaves, int cap_height, size_t leaf_index,
                             uint64_t* siblings_out_host);
int merkle_download_digests(etp_ctx* ctx, const uint64_t* levels, size_t n_leaves, int cap_height, uint64_t* out_host);

inline int log2_exact(size_t n) {
  if (n == 0 || (n & (n - 1))) return -1;
  int l = 0;
  while (((size_t)1 << l) < n) l++;
  return l;
}

struct etp_tree {
  etp_ctx* ctx;
  size_t n_leaves, leaf_len;
  int cap_height;
  uint64_t* levels = nullptr;
  std::vector<uint64_t> cap;
};

struct etp_batch {
  etp_ctx* ctx;
  size_t n_cols;
  int log_n, rate_bits, cap_height;
  uint64_t* coeffs = nullptr;  // n_cols x n
  uint64_t* lde = nullptr;     // n_cols x (n << rate_bits), bit-reversed row order
  uint64_t* levels = nullptr;
  std::vector<uint64_t> cap;
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // phase boundaries of the last commit
  bool timed_ifft = false;
  size_t n() const { return (size_t)1 << log_n; }
  size_t lde_n() const { return (size_t)1 << (log_n + rate_bits); }
};
int batch_create(etp_ctx* ctx, size_t n_cols, int log_n, int rate_bits, int blinding, int cap_height, etp_batch** out);
// coeffs already in b->coeffs: LDE + leaf hashing + tree
int batch_commit_from_coeffs(etp_batch* b);
// host columns -> commit, column group by column group: group k is transformed and absorbed by the leaf
// sponges while group k+1 crosses PCIe.  is_values: run the iFFT first (from_values) or take the columns as
// coefficients (from_coeffs).
int batch_commit_from_host_streamed(etp_batch* b, const uint64_t* const* cols, bool is_values, uint64_t* keep_values);
int launch_leaf_hash(etp_ctx* ctx, const merkle::LeafSrc& src, int c_begin, int c_end, int n_cols_total, uint32_t row0,
                     uint32_t n_rows, uint64_t* digests);
int get_sync_event(etp_ctx* ctx, size_t i, cudaEvent_t* out);
// values (device, natural order) -> coeffs -> commit
int batch_commit_from_values(etp_batch* b, const uint64_t* values_dev, size_t col_stride);
