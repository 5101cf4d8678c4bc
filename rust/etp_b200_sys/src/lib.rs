//! Raw bindings of `include/etp_b200.h` + the safe wrappers a patched `plonky2` / `starky` would call.
//! SOURCE ONLY — not compiled here (no Rust toolchain in the build image). See INTEGRATION.md.
#![allow(non_camel_case_types)]
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct etp_ctx { _p: [u8; 0] }
#[repr(C)] pub struct etp_batch { _p: [u8; 0] }
#[repr(C)] pub struct etp_tree { _p: [u8; 0] }

#[repr(C)] pub struct etp_shard { _p: [u8; 0] }
pub mod recorder;

extern "C" {
    pub fn etp_ctx_create(device: c_int, out: *mut *mut etp_ctx) -> c_int;
    pub fn etp_ctx_destroy(ctx: *mut etp_ctx);
    pub fn etp_ctx_trim(ctx: *mut etp_ctx) -> c_int;
    pub fn etp_host_pin(ctx: *mut etp_ctx, ptr: *mut c_void, bytes: usize) -> c_int;
    pub fn etp_host_unpin(ctx: *mut etp_ctx, ptr: *mut c_void) -> c_int;
    pub fn etp_host_poseidon_permute(state: *mut u64);
    pub fn etp_ctx_cached_bytes(ctx: *const etp_ctx) -> usize;
    pub fn etp_last_error(ctx: *const etp_ctx) -> *const c_char;
    pub fn etp_batch_from_values_host(ctx: *mut etp_ctx, cols: *const *const u64, n_cols: usize, log_n: c_int,
        rate_bits: c_int, blinding: c_int, cap_height: c_int, out: *mut *mut etp_batch) -> c_int;
    pub fn etp_batch_from_coeffs_host(ctx: *mut etp_ctx, cols: *const *const u64, n_cols: usize, log_n: c_int,
        rate_bits: c_int, blinding: c_int, cap_height: c_int, out: *mut *mut etp_batch) -> c_int;
    pub fn etp_batch_free(b: *mut etp_batch);
    pub fn etp_batch_cap(b: *mut etp_batch, cap_out: *mut u64) -> c_int;
    pub fn etp_batch_download_coeffs(b: *mut etp_batch, out: *mut u64) -> c_int;
    pub fn etp_batch_download_leaves(b: *mut etp_batch, out: *mut u64) -> c_int;
    pub fn etp_batch_download_digests(b: *mut etp_batch, out: *mut u64) -> c_int;
    pub fn etp_batch_num_digests(b: *const etp_batch) -> usize;
    pub fn etp_batch_leaves_at(b: *mut etp_batch, idx: *const u64, n_idx: usize, rows_out: *mut u64) -> c_int;
    pub fn etp_batch_prove(b: *mut etp_batch, leaf_index: usize, siblings_out: *mut u64) -> c_int;
    pub fn etp_merkle_new_host(ctx: *mut etp_ctx, leaves: *const u64, n_leaves: usize, leaf_len: usize,
        cap_height: c_int, out: *mut *mut etp_tree) -> c_int;
    pub fn etp_tree_cap(t: *mut etp_tree, cap_out: *mut u64) -> c_int;
    pub fn etp_tree_digests(t: *mut etp_tree, digests_out: *mut u64) -> c_int;
    pub fn etp_tree_num_digests(t: *const etp_tree) -> usize;
    pub fn etp_tree_free(t: *mut etp_tree);
    pub fn etp_compute_quotient_polys_dev(ctx: *mut etp_ctx, table: c_int, trace: *mut etp_batch, aux: *mut etp_batch,
        lookup_challenges: *const u64, n_lookup_challenges: c_int, public_inputs: *const u64,
        alphas: *const u64, n_alphas: c_int, out_dev: *mut u64) -> c_int;
    pub fn etp_stark_proof_words(ctx: *const etp_ctx, table: c_int, log_n: c_int) -> usize;
    // program-defined tables (csrc/cprog.h): what `recorder::record_table` produces
    pub fn etp_table_register(ctx: *mut etp_ctx, program: *const u64, n_words: usize, lookups: *const i32,
        n_lookup_words: usize, table_id_out: *mut c_int) -> c_int;
    pub fn etp_cprog_compile_check(program: *const u64, n_words: usize, cubin_bytes_out: *mut usize, err: *mut c_char,
        err_len: usize) -> c_int;
    // column-split commit of one oversized table across the GPUs of a box (one worker process per GPU)
    pub fn etp_shard_create(ctx: *mut etp_ctx, n_cols_total: usize, log_n: c_int, rate_bits: c_int, cap_height: c_int,
        rank: c_int, world: c_int, out: *mut *mut etp_shard) -> c_int;
    pub fn etp_shard_free(s: *mut etp_shard);
    pub fn etp_shard_transform_values_host(s: *mut etp_shard, local_cols: *const *const u64) -> c_int;
    pub fn etp_shard_lde_dev(s: *const etp_shard) -> *const u64;
    pub fn etp_ipc_export(ctx: *mut etp_ctx, dev_ptr: *const c_void, handle_out: *mut u8) -> c_int;
    pub fn etp_ipc_open(ctx: *mut etp_ctx, handle: *const u8, dev_ptr_out: *mut *mut c_void) -> c_int;
    pub fn etp_ipc_close(ctx: *mut etp_ctx, dev_ptr: *mut c_void) -> c_int;
    pub fn etp_shard_set_peer(s: *mut etp_shard, peer_rank: c_int, peer_lde: *const u64) -> c_int;
    pub fn etp_shard_commit_rows(s: *mut etp_shard, cap_part_out: *mut u64) -> c_int;
    pub fn etp_shard_prove(s: *mut etp_shard, leaf_index: usize, siblings_out: *mut u64) -> c_int;
    pub fn etp_shard_leaves_at(s: *mut etp_shard, idx: *const u64, n_idx: usize, rows_out: *mut u64) -> c_int;
    pub fn etp_stark_prove_host(ctx: *mut etp_ctx, table: c_int, log_n: c_int, trace: *const u64,
        public_inputs: *const u64, proof_out: *mut u64) -> c_int;
    pub fn etp_pow_grind(ctx: *mut etp_ctx, state: *const u64, pos: c_int, bits: c_int, witness_out: *mut u64) -> c_int;
    pub fn etp_dev_alloc(ctx: *mut etp_ctx, bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn etp_dev_free(ctx: *mut etp_ctx, ptr: *mut c_void) -> c_int;
}

/// One context per worker thread (one Paladin worker <-> one GPU: CUDA_VISIBLE_DEVICES=%i).
pub struct Ctx(pub *mut etp_ctx);
unsafe impl Send for Ctx {}
impl Ctx {
    pub fn new(device: i32) -> Self {
        let mut p = std::ptr::null_mut();
        let rc = unsafe { etp_ctx_create(device, &mut p) };
        assert_eq!(rc, 0, "etp_b200: no usable CUDA device {device} (there is no CPU fallback)");
        Ctx(p)
    }
    /// Upstream's functions are infallible by signature and panic on misuse; keep that contract.
    pub fn check(&self, rc: c_int) {
        if rc != 0 {
            let msg = unsafe { CStr::from_ptr(etp_last_error(self.0)) }.to_string_lossy().into_owned();
            panic!("etp_b200 error {rc}: {msg}");
        }
    }
}
impl Drop for Ctx { fn drop(&mut self) { unsafe { etp_ctx_destroy(self.0) } } }

/// Device-resident `PolynomialBatch<GoldilocksField, PoseidonGoldilocksConfig, 2>`.
pub struct DeviceBatch { pub raw: *mut etp_batch, pub n_cols: usize, pub degree_log: usize, pub rate_bits: usize, pub cap_height: usize }
impl Drop for DeviceBatch { fn drop(&mut self) { unsafe { etp_batch_free(self.raw) } } }

impl DeviceBatch {
    /// `PolynomialBatch::from_values(values, rate_bits, blinding=false, cap_height, ..)`; `values[c]` is the
    /// `Vec<u64>` behind `PolynomialValues<GoldilocksField>` (GoldilocksField is `repr(transparent)` over u64).
    pub fn from_values(ctx: &Ctx, values: &[&[u64]], rate_bits: usize, blinding: bool, cap_height: usize) -> Self {
        let n = values.first().map_or(1, |v| v.len());
        assert!(values.iter().all(|v| v.len() == n), "Polynomial degrees inconsistent");
        let ptrs: Vec<*const u64> = values.iter().map(|v| v.as_ptr()).collect();
        let mut raw = std::ptr::null_mut();
        ctx.check(unsafe { etp_batch_from_values_host(ctx.0, ptrs.as_ptr(), ptrs.len(), n.trailing_zeros() as c_int,
            rate_bits as c_int, blinding as c_int, cap_height as c_int, &mut raw) });
        DeviceBatch { raw, n_cols: values.len(), degree_log: n.trailing_zeros() as usize, rate_bits, cap_height }
    }
    /// `merkle_tree.cap` as 2^cap_height digests of 4 u64.
    pub fn cap(&self, ctx: &Ctx) -> Vec<[u64; 4]> {
        let mut out = vec![[0u64; 4]; 1 << self.cap_height];
        ctx.check(unsafe { etp_batch_cap(self.raw, out.as_mut_ptr() as *mut u64) });
        out
    }
    /// `merkle_tree.leaves[i]` for the query rounds (only queried rows ever cross PCIe).
    pub fn leaves_at(&self, ctx: &Ctx, idx: &[u64]) -> Vec<Vec<u64>> {
        let mut flat = vec![0u64; idx.len() * self.n_cols];
        ctx.check(unsafe { etp_batch_leaves_at(self.raw, idx.as_ptr(), idx.len(), flat.as_mut_ptr()) });
        flat.chunks(self.n_cols.max(1)).map(|r| r.to_vec()).collect()
    }
    /// `merkle_tree.prove(leaf_index).siblings`
    pub fn prove(&self, ctx: &Ctx, leaf_index: usize) -> Vec<[u64; 4]> {
        let mut out = vec![[0u64; 4]; self.degree_log + self.rate_bits - self.cap_height];
        ctx.check(unsafe { etp_batch_prove(self.raw, leaf_index, out.as_mut_ptr() as *mut u64) });
        out
    }
}
