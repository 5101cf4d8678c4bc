// Column-split commit of ONE oversized table across the GPUs of a box (BASELINE.json
// north_star: "NVLink peer copies or NCCL are used only ... when a single oversized table is column-split
// across GPUs").  One process per GPU; rank g of G owns
//   * columns [g*cps, min((g+1)*cps, C))           (cps = columns per shard, a multiple of 8), and
//   * leaf rows [g*L/G, (g+1)*L/G), L = n << rate_bits = the cap subtrees [g*2^h/G, (g+1)*2^h/G).
// Phase 1 (local): iFFT + coset LDE of the own columns — columns are independent, no traffic.
// Exchange: every rank exports its LDE matrix as a CUDA IPC handle; peers map it (NVLink P2P).
// Phase 2: the leaf-hash kernel of rank g walks its row range and reads the 8-column sponge chunks of
//   every shard straight from the owner's HBM over NVLink (merkle::LeafSrc) — the all-gather of row
//   tiles is fused into the hashing kernel, which is integer-bound (≈90 GB/s of input per GPU), far
//   below NVLink bandwidth.  Each rank then builds its own cap subtrees; the cap is the concatenation
//   of the G parts (a 2^h*32-byte all-gather done by the caller over torch.distributed).
// Same field elements as PolynomialBatch::from_values on the whole table (plonky2/src/fri/oracle.rs):
// tests compare the assembled cap and the Merkle paths with the unsplit commit.
#include "shard.cuh"


extern "C" size_t etp_shard_cols_per_rank(size_t n_cols_total, int world) {
  if (world <= 0) return 0;
  size_t cps = (n_cols_total + world - 1) / world;
  return (cps + 7) / 8 * 8;
}

extern "C" int etp_shard_create(etp_ctx* ctx, size_t n_cols_total, int log_n, int rate_bits, int cap_height, int rank, int world,
                                etp_shard** out) {
  etp_bind(ctx);
  if (!ctx || !out) return ETP_ERR_INVALID;
  *out = nullptr;
  if (world < 1 || world > merkle::MAX_SRC || (world & (world - 1)) || rank < 0 || rank >= world)
    return etp_fail(ctx, ETP_ERR_INVALID, "shard: world must be a power of two <= %d and 0 <= rank < world", merkle::MAX_SRC);
  if (log_n < 0 || rate_bits < 0 || log_n + rate_bits > 31) return etp_fail(ctx, ETP_ERR_INVALID, "bad degree / rate");
  if (cap_height < 0 || cap_height > log_n + rate_bits)
    return etp_fail(ctx, ETP_ERR_INVALID, "cap_height=%d should be at most log2(leaves.len())=%d", cap_height, log_n + rate_bits);
  if ((1 << cap_height) < world)
    return etp_fail(ctx, ETP_ERR_INVALID, "shard: 2^cap_height=%d must be >= world=%d (every rank owns whole cap subtrees)",
                    1 << cap_height, world);
  if (n_cols_total == 0 || n_cols_total > 65535) return etp_fail(ctx, ETP_ERR_INVALID, "shard: bad column count");
  etp_shard* s = new etp_shard();
  s->ctx = ctx; s->n_cols_total = n_cols_total; s->log_n = log_n; s->rate_bits = rate_bits; s->cap_height = cap_height;
  s->rank = rank; s->world = world;
  s->cps = etp_shard_cols_per_rank(n_cols_total, world);
  s->c0 = s->cps * rank;
  s->local_cols = s->c0 >= n_cols_total ? 0 : (s->c0 + s->cps <= n_cols_total ? s->cps : n_cols_total - s->c0);
  cudaSetDevice(ctx->device);
  const size_t lc = s->local_cols ? s->local_cols : 1;
  cudaError_t e = cudaMalloc((void**)&s->coeffs, lc * s->n() * 8);
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->lde, lc * s->lde_n() * 8);
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->levels, levels_words(s->rows(), s->local_cap_height()) * 8);
  if (e != cudaSuccess) {
    etp_shard_free(s);
    return etp_fail(ctx, ETP_ERR_CUDA, "shard: allocation failed: %s", cudaGetErrorString(e));
  }
  s->peer[rank] = s->lde;
  *out = s;
  return ETP_OK;
}

extern "C" void etp_shard_free(etp_shard* s) {
  etp_bind(s ? s->ctx : nullptr);
  if (!s) return;
  cudaSetDevice(s->ctx->device);
  cudaStreamSynchronize(s->ctx->stream);
  cudaFree(s->coeffs);
  cudaFree(s->lde);
  cudaFree(s->levels);
  delete s;
}

extern "C" size_t etp_shard_first_col(const etp_shard* s) { return s ? s->c0 : 0; }
extern "C" size_t etp_shard_num_local_cols(const etp_shard* s) { return s ? s->local_cols : 0; }
extern "C" size_t etp_shard_first_row(const etp_shard* s) { return s ? s->row0() : 0; }
extern "C" size_t etp_shard_num_rows(const etp_shard* s) { return s ? s->rows() : 0; }
extern "C" const uint64_t* etp_shard_lde_dev(const etp_shard* s) { return s ? s->lde : nullptr; }

static int shard_transform(etp_shard* s, const uint64_t* values_dev, size_t col_stride, DevBuf<uint64_t>* scratch) {
  etp_ctx* ctx = s->ctx;
  if (s->local_cols == 0) return ETP_OK;
  NttArgs a;  // "IFFT"; the LDE buffer doubles as scratch
  a.in = values_dev; a.in_stride = col_stride; a.n_in = (uint32_t)s->n();
  a.out = s->coeffs; a.out_stride = s->n();
  a.scratch = s->lde; a.scratch_stride = s->lde_n();
  a.log_n = s->log_n; a.n_cols = s->local_cols; a.inverse = true; a.natural_out = true;
  ETP_TRY(ntt_run(ctx, a));
  a = NttArgs();  // "FFT + blinding"
  a.in = s->coeffs; a.in_stride = s->n(); a.n_in = (uint32_t)s->n();
  a.out = s->lde; a.out_stride = s->lde_n();
  a.log_n = s->log_n + s->rate_bits; a.n_cols = s->local_cols;
  a.coset_shift = gl::GENERATOR;
  ETP_TRY(ntt_run(ctx, a));
  (void)scratch;
  return ETP_OK;
}

extern "C" int etp_shard_transform_values_dev(etp_shard* s, const uint64_t* values_dev, size_t col_stride) {
  etp_bind(s ? s->ctx : nullptr);
  if (!s || (!values_dev && s->local_cols)) return ETP_ERR_INVALID;
  ETP_TRY(shard_transform(s, values_dev, col_stride, nullptr));
  ETP_CUDA(s->ctx, cudaStreamSynchronize(s->ctx->stream));  // peers may read the LDE once this returns (+ a barrier)
  s->committed = false;
  return ETP_OK;
}

extern "C" int etp_shard_transform_values_host(etp_shard* s, const uint64_t* const* local_cols) {
  etp_bind(s ? s->ctx : nullptr);
  if (!s || (!local_cols && s->local_cols)) return ETP_ERR_INVALID;
  etp_ctx* ctx = s->ctx;
  DevBuf<uint64_t> stage(ctx);
  ETP_TRY(stage.alloc(s->local_cols * s->n()));
  for (size_t c = 0; c < s->local_cols; c++) {
    if (!local_cols[c]) return etp_fail(ctx, ETP_ERR_INVALID, "null column %zu", c);
    ETP_CUDA(ctx, cudaMemcpyAsync(stage.p + c * s->n(), local_cols[c], s->n() * 8, cudaMemcpyHostToDevice, ctx->stream));
  }
  ETP_TRY(shard_transform(s, stage.p, s->n(), nullptr));
  ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  s->committed = false;
  return ETP_OK;
}

// ---- peer mapping ---------------------------------------------------------------------------------------
extern "C" int etp_ipc_export(etp_ctx* ctx, const void* dev_ptr, unsigned char handle_out[ETP_IPC_HANDLE_BYTES]) {
  etp_bind(ctx);
  if (!ctx || !dev_ptr || !handle_out) return ETP_ERR_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == ETP_IPC_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  ETP_CUDA(ctx, cudaSetDevice(ctx->device));
  ETP_CUDA(ctx, cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)));
  memcpy(handle_out, &h, sizeof h);
  return ETP_OK;
}
extern "C" int etp_ipc_open(etp_ctx* ctx, const unsigned char handle[ETP_IPC_HANDLE_BYTES], void** dev_ptr_out) {
  etp_bind(ctx);
  if (!ctx || !handle || !dev_ptr_out) return ETP_ERR_INVALID;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof h);
  ETP_CUDA(ctx, cudaSetDevice(ctx->device));
  ETP_CUDA(ctx, cudaIpcOpenMemHandle(dev_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return ETP_OK;
}
extern "C" int etp_ipc_close(etp_ctx* ctx, void* dev_ptr) {
  etp_bind(ctx);
  if (!ctx) return ETP_ERR_INVALID;
  ETP_CUDA(ctx, cudaSetDevice(ctx->device));
  ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ETP_CUDA(ctx, cudaIpcCloseMemHandle(dev_ptr));
  return ETP_OK;
}
extern "C" int etp_shard_set_peer(etp_shard* s, int peer_rank, const uint64_t* peer_lde) {
  etp_bind(s ? s->ctx : nullptr);
  if (!s) return ETP_ERR_INVALID;
  if (peer_rank < 0 || peer_rank >= s->world || peer_rank == s->rank || !peer_lde)
    return etp_fail(s->ctx, ETP_ERR_INVALID, "shard: bad peer rank %d", peer_rank);
  s->peer[peer_rank] = peer_lde;
  return ETP_OK;
}

static int shard_src(etp_shard* s, merkle::LeafSrc* src) {
  *src = merkle::LeafSrc{};
  const int n_src = (int)((s->n_cols_total + s->cps - 1) / s->cps);
  for (int g = 0; g < n_src; g++) {
    if (!s->peer[g]) return etp_fail(s->ctx, ETP_ERR_STATE, "shard: LDE of rank %d not mapped (etp_shard_set_peer)", g);
    src->base[g] = s->peer[g];
  }
  src->col_stride = s->lde_n();
  src->cols_per_src = (int)s->cps;
  return ETP_OK;
}

// ---- phase 2 ------------------------------------------------------------------------------------------------
extern "C" int etp_shard_commit_rows(etp_shard* s, uint64_t* cap_part_out) {
  etp_bind(s ? s->ctx : nullptr);
  if (!s || !cap_part_out) return ETP_ERR_INVALID;
  etp_ctx* ctx = s->ctx;
  merkle::LeafSrc src;
  ETP_TRY(shard_src(s, &src));
  ETP_TRY(launch_leaf_hash(ctx, src, 0, (int)s->n_cols_total, (int)s->n_cols_total, (uint32_t)s->row0(), (uint32_t)s->rows(),
                           s->levels));
  ETP_TRY(merkle_build_levels(ctx, s->levels, s->rows(), s->local_cap_height(), cap_part_out));
  s->committed = true;
  return ETP_OK;
}

extern "C" int etp_shard_prove(etp_shard* s, size_t leaf_index, uint64_t* siblings_out) {
  etp_bind(s ? s->ctx : nullptr);
  if (!s) return ETP_ERR_INVALID;
  if (!s->committed) return etp_fail(s->ctx, ETP_ERR_STATE, "shard: not committed");
  if (leaf_index < s->row0() || leaf_index >= s->row0() + s->rows())
    return etp_fail(s->ctx, ETP_ERR_INVALID, "shard: leaf %zu is owned by another rank", leaf_index);
  return merkle_prove_from_levels(s->ctx, s->levels, s->rows(), s->local_cap_height(), leaf_index - s->row0(), siblings_out);
}

// rows[q][c] = column c (any owner) at leaf idx[q]
__global__ void k_gather_rows_src(const __grid_constant__ merkle::LeafSrc src, int n_cols, const uint64_t* __restrict__ idx, int n_idx,
                                  uint64_t* __restrict__ rows) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_idx * n_cols) return;
  const int q = t / n_cols, c = t % n_cols, si = c / src.cols_per_src;
  rows[t] = gl::canon(src.base[si][(size_t)(c - si * src.cols_per_src) * src.col_stride + idx[q]]);
}

extern "C" int etp_shard_leaves_at(etp_shard* s, const uint64_t* idx, size_t n_idx, uint64_t* rows_out) {
  etp_bind(s ? s->ctx : nullptr);
  if (!s || (!idx && n_idx)) return ETP_ERR_INVALID;
  if (n_idx == 0) return ETP_OK;
  etp_ctx* ctx = s->ctx;
  for (size_t q = 0; q < n_idx; q++)
    if (idx[q] >= s->lde_n()) return etp_fail(ctx, ETP_ERR_INVALID, "leaf index out of range");
  merkle::LeafSrc src;
  ETP_TRY(shard_src(s, &src));
  DevBuf<uint64_t> d_idx(ctx), d_rows(ctx);
  ETP_TRY(d_idx.alloc(n_idx));
  ETP_TRY(d_rows.alloc(n_idx * s->n_cols_total));
  ETP_CUDA(ctx, cudaMemcpyAsync(d_idx.p, idx, n_idx * 8, cudaMemcpyHostToDevice, ctx->stream));
  const size_t tot = n_idx * s->n_cols_total;
  k_gather_rows_src<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(src, (int)s->n_cols_total, d_idx.p, (int)n_idx, d_rows.p);
  ETP_LAUNCH_CHECK(ctx);
  ETP_CUDA(ctx, cudaMemcpyAsync(rows_out, d_rows.p, tot * 8, cudaMemcpyDeviceToHost, ctx->stream));
  ETP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ETP_OK;
}

extern "C" int etp_shard_download_coeffs(etp_shard* s, uint64_t* out) {
  etp_bind(s ? s->ctx : nullptr);
  if (!s || (!out && s->local_cols)) return ETP_ERR_INVALID;
  if (s->local_cols == 0) return ETP_OK;
  ETP_CUDA(s->ctx, cudaMemcpyAsync(out, s->coeffs, s->local_cols * s->n() * 8, cudaMemcpyDeviceToHost, s->ctx->stream));
  ETP_CUDA(s->ctx, cudaStreamSynchronize(s->ctx->stream));
  return ETP_OK;
}
