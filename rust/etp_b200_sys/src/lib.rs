//! `etp_b200_sys` — Rust side of libetp_b200: the generated raw bindings (`sys`, every function of `include/etp_b200.h`),
//! safe wrappers shaped like the plonky2 / starky items they stand behind, and the constraint-program recorder.
//! SOURCE ONLY — not compiled in this repository (no Rust toolchain in the build image); see INTEGRATION.md for how a
//! plonky2 0.2.2 / starky 0.4.0 fork uses it (/root/reference/Cargo.toml `[patch.crates-io]`).
#![allow(non_camel_case_types)]
pub mod recorder;
pub mod sys;
pub use sys::*;

use std::ffi::CStr;
use std::os::raw::{c_char, c_int};
use std::ptr;

/// One CUDA device + stream + scratch cache.  One per worker thread (`thread_local!`), like upstream's one tokio task per op
/// (/root/reference/ops/src/lib.rs:29-61).
pub struct Ctx(pub *mut etp_ctx);
unsafe impl Send for Ctx {}

impl Ctx {
    /// `device`: index inside CUDA_VISIBLE_DEVICES (a Paladin worker pinned to one GPU passes 0).
    pub fn new(device: i32) -> Self {
        let mut p = ptr::null_mut();
        let rc = unsafe { etp_ctx_create(device as c_int, &mut p) };
        assert!(rc == ETP_OK, "etp_ctx_create({device}) failed ({rc}): no usable CUDA device; there is no CPU fallback");
        Ctx(p)
    }
    /// Upstream's functions are infallible by signature and panic on misuse; so does the glue.
    pub fn check(&self, rc: c_int) {
        if rc != ETP_OK {
            let msg = unsafe { CStr::from_ptr(etp_last_error(self.0)) }.to_string_lossy().into_owned();
            panic!("etp_b200 error {rc}: {msg}");
        }
    }
}
impl Drop for Ctx {
    fn drop(&mut self) { unsafe { etp_ctx_destroy(self.0) } }
}

/// Device-resident `PolynomialBatch<GoldilocksField, PoseidonGoldilocksConfig, 2>` (plonky2/src/fri/oracle.rs).
pub struct DeviceBatch { pub raw: *mut etp_batch, pub n_cols: usize, pub degree_log: usize, pub rate_bits: usize, pub cap_height: usize }
unsafe impl Send for DeviceBatch {}

impl DeviceBatch {
    /// `PolynomialBatch::from_values(values, rate_bits, blinding, cap_height, ..)`; `cols[c]` = the u64 view of column c
    /// (`GoldilocksField` is `repr(transparent)` over u64).
    pub fn from_values(ctx: &Ctx, cols: &[&[u64]], rate_bits: usize, blinding: bool, cap_height: usize) -> Self {
        Self::build(ctx, cols, rate_bits, blinding, cap_height, true)
    }
    /// `PolynomialBatch::from_coeffs`
    pub fn from_coeffs(ctx: &Ctx, cols: &[&[u64]], rate_bits: usize, blinding: bool, cap_height: usize) -> Self {
        Self::build(ctx, cols, rate_bits, blinding, cap_height, false)
    }
    fn build(ctx: &Ctx, cols: &[&[u64]], rate_bits: usize, blinding: bool, cap_height: usize, values: bool) -> Self {
        let n = cols.first().map_or(1, |c| c.len());
        assert!(n.is_power_of_two() && cols.iter().all(|c| c.len() == n), "all polynomials must have the same power-of-two length");
        let ptrs: Vec<*const u64> = cols.iter().map(|c| c.as_ptr()).collect();
        let mut raw = ptr::null_mut();
        let log_n = n.trailing_zeros() as c_int;
        let rc = unsafe {
            if values {
                etp_batch_from_values_host(ctx.0, ptrs.as_ptr(), cols.len(), log_n, rate_bits as c_int, blinding as c_int, cap_height as c_int, &mut raw)
            } else {
                etp_batch_from_coeffs_host(ctx.0, ptrs.as_ptr(), cols.len(), log_n, rate_bits as c_int, blinding as c_int, cap_height as c_int, &mut raw)
            }
        };
        ctx.check(rc);
        DeviceBatch { raw, n_cols: cols.len(), degree_log: log_n as usize, rate_bits, cap_height }
    }
    /// `merkle_tree.cap` as 2^cap_height x 4 words
    pub fn cap(&self) -> Vec<[u64; 4]> {
        let mut out = vec![[0u64; 4]; 1 << self.cap_height];
        let rc = unsafe { etp_batch_cap(self.raw, out.as_mut_ptr() as *mut u64) };
        assert!(rc == ETP_OK);
        out
    }
    /// `polynomials` (lazy public field): n_cols x 2^degree_log, column-major
    pub fn polynomials(&self) -> Vec<u64> {
        let mut out = vec![0u64; self.n_cols << self.degree_log];
        let rc = unsafe { etp_batch_download_coeffs(self.raw, out.as_mut_ptr()) };
        assert!(rc == ETP_OK);
        out
    }
    /// `merkle_tree.leaves` (lazy public field): (2^degree_log << rate_bits) rows x n_cols
    pub fn leaves(&self) -> Vec<u64> {
        let mut out = vec![0u64; (self.n_cols << self.degree_log) << self.rate_bits];
        let rc = unsafe { etp_batch_download_leaves(self.raw, out.as_mut_ptr()) };
        assert!(rc == ETP_OK);
        out
    }
    /// `merkle_tree.digests` in plonky2's layout, 4 words per digest
    pub fn digests(&self) -> Vec<u64> {
        let nd = unsafe { etp_batch_num_digests(self.raw) };
        let mut out = vec![0u64; 4 * nd];
        let rc = unsafe { etp_batch_download_digests(self.raw, out.as_mut_ptr()) };
        assert!(rc == ETP_OK);
        out
    }
    /// `merkle_tree.prove(leaf_index)`
    pub fn prove(&self, leaf_index: usize) -> Vec<[u64; 4]> {
        let mut out = vec![[0u64; 4]; self.degree_log + self.rate_bits - self.cap_height];
        let rc = unsafe { etp_batch_prove(self.raw, leaf_index, out.as_mut_ptr() as *mut u64) };
        assert!(rc == ETP_OK);
        out
    }
    /// `StarkOpeningSet::new`'s eval_commitment: every polynomial at the extension point z = (c0, c1)
    pub fn eval_at_ext_point(&self, z: [u64; 2]) -> Vec<[u64; 2]> {
        let mut out = vec![[0u64; 2]; self.n_cols];
        let rc = unsafe { etp_batch_eval_at_ext_point(self.raw, z.as_ptr(), out.as_mut_ptr() as *mut u64) };
        assert!(rc == ETP_OK);
        out
    }
}
impl Drop for DeviceBatch {
    fn drop(&mut self) { unsafe { etp_batch_free(self.raw) } }
}

/// The transcript as the C ABI sees it; convert from / to plonky2's `Challenger` field by field
/// (`sponge_state`, `input_buffer`, `output_buffer` are public in the fork).
impl etp_challenger {
    pub fn new() -> Self {
        let mut c = etp_challenger { sponge_state: [0; 12], input_buffer: [0; 8], output_buffer: [0; 8], input_len: 0, output_len: 0 };
        unsafe { etp_challenger_init(&mut c) };
        c
    }
    pub fn from_parts(sponge_state: [u64; 12], input_buffer: &[u64], output_buffer: &[u64]) -> Self {
        assert!(input_buffer.len() < 8 && output_buffer.len() <= 8);
        let mut c = Self::new();
        c.sponge_state = sponge_state;
        c.input_buffer[..input_buffer.len()].copy_from_slice(input_buffer);
        c.output_buffer[..output_buffer.len()].copy_from_slice(output_buffer);
        c.input_len = input_buffer.len() as u32;
        c.output_len = output_buffer.len() as u32;
        c
    }
    pub fn observe_elements(&mut self, e: &[u64]) { unsafe { etp_challenger_observe(self, e.as_ptr(), e.len()) } }
    pub fn get_challenge(&mut self) -> u64 { unsafe { etp_challenger_get_challenge(self) } }
    pub fn get_n_challenges(&mut self, n: usize) -> Vec<u64> {
        let mut out = vec![0u64; n];
        unsafe { etp_challenger_get_n_challenges(self, n, out.as_mut_ptr()) };
        out
    }
    pub fn compact(&mut self) -> [u64; 12] {
        unsafe { etp_challenger_compact(self) };
        self.sponge_state
    }
}

/// `starky::prover::prove_with_commitment` for a registered table: returns the flat "B200STK2" proof words that
/// `StarkProof::from_flat` in the fork re-shapes into `StarkProofWithPublicInputs`; the challenger is advanced in place.
pub fn prove_with_commitment(ctx: &Ctx, table: i32, trace_commitment: &DeviceBatch, trace_dev: *const u64, col_stride: usize,
                             ctl_challenges: Option<&[u64; 4]>, challenger: &mut etp_challenger, public_inputs: &[u64]) -> Vec<u64> {
    let words = unsafe { etp_stark_proof_words(ctx.0, table as c_int, trace_commitment.degree_log as c_int) };
    assert!(words > 0, "unknown table {table}");
    let mut proof = vec![0u64; words];
    let mut pi = public_inputs.to_vec();
    pi.push(0);
    let rc = unsafe {
        etp_prove_with_commitment(ctx.0, table as c_int, trace_commitment.raw, trace_dev, col_stride,
                                  ctl_challenges.map_or(ptr::null(), |c| c.as_ptr()), challenger, pi.as_ptr(), proof.as_mut_ptr())
    };
    ctx.check(rc);
    proof
}

/// `PolynomialBatch::prove_openings(instance, oracles, challenger, fri_params, timing)` -> flat FriProof words.
pub fn prove_openings(ctx: &Ctx, batches: &[([u64; 2], Vec<etp_fri_poly>)], oracles: &[&DeviceBatch], challenger: &mut etp_challenger,
                      params: &etp_fri_params) -> Vec<u64> {
    let raw_batches: Vec<etp_fri_batch> =
        batches.iter().map(|(point, polys)| etp_fri_batch { point: *point, polynomials: polys.as_ptr(), n_polynomials: polys.len() }).collect();
    let handles: Vec<*mut etp_batch> = oracles.iter().map(|o| o.raw).collect();
    let n_cols: Vec<usize> = oracles.iter().map(|o| o.n_cols).collect();
    let words = unsafe { etp_fri_proof_words(n_cols.as_ptr(), n_cols.len(), params) };
    let mut out = vec![0u64; words];
    let rc = unsafe {
        etp_prove_openings(ctx.0, raw_batches.as_ptr(), raw_batches.len(), handles.as_ptr(), handles.len(), challenger, params, out.as_mut_ptr())
    };
    ctx.check(rc);
    out
}

/// Registers a table from a recorded program + auxiliary-column spec (`recorder::finish`, `recorder::AuxSpecBuilder`).
pub fn register_table(ctx: &Ctx, program: &[u64], aux_spec: &[u64]) -> i32 {
    let mut id: c_int = 0;
    let rc = unsafe { etp_table_register_ex(ctx.0, program.as_ptr(), program.len(), aux_spec.as_ptr(), aux_spec.len(), &mut id) };
    ctx.check(rc);
    id as i32
}

/// Witness of one PoseidonGate row (`plonky2::gates::poseidon::PoseidonGenerator::run_once`): the 135 wires for the given inputs.
/// The recursion layers' circuits are mostly PoseidonGate rows, so a generator in the fork can call this instead of the scalar
/// permutation.
pub fn poseidon_gate_wires(inputs: &[u64; 12], swap: bool) -> [u64; 135] {
    let mut wires = [0u64; 135];
    unsafe { etp_host_poseidon_gate_wires(inputs.as_ptr(), swap as c_int, wires.as_mut_ptr()) };
    wires
}

/// The CUDA source the library generates for a constraint program (inspection / CI): one kernel, or segment functions above 16384 ops.
pub fn generated_cuda(program: &[u64]) -> Option<String> {
    let n = unsafe { etp_cprog_generate_cuda(program.as_ptr(), program.len(), std::ptr::null_mut(), 0) };
    if n < 0 { return None; }
    let mut buf = vec![0u8; n as usize + 1];
    unsafe { etp_cprog_generate_cuda(program.as_ptr(), program.len(), buf.as_mut_ptr() as *mut c_char, buf.len()) };
    buf.truncate(n as usize);
    String::from_utf8(buf).ok()
}


/// plonky2's circuit prover on the device (`plonk::prover::prove` after witness generation): built once per circuit from the
/// recorded vanishing polynomial (run `eval_vanishing_poly` on `recorder::Sym` over the virtual columns constants, sigmas,
/// wires, Zs, partial products, X — terms emitted in reverse — and take `recorder::finish()`), the constants and sigma
/// polynomials' values and the coset shifts; `prove` is one call per proof.
pub struct DeviceCircuit { pub raw: *mut etp_circuit, pub proof_words: usize }
unsafe impl Send for DeviceCircuit {}

impl DeviceCircuit {
    /// `constants` / `sigmas`: values on the subgroup, column-major (`num_constants x n`, `num_routed_wires x n`).
    #[allow(clippy::too_many_arguments)]
    pub fn new(ctx: &Ctx, vanishing_program: &[u64], constants: &[u64], num_constants: usize, sigmas: &[u64], k_is: &[u64], num_wires: usize,
               degree_bits: usize, quotient_degree_factor: usize, num_challenges: usize, fri_params: &etp_fri_params,
               circuit_digest: Option<[u64; 4]>) -> Self {
        let n = 1usize << degree_bits;
        assert_eq!(constants.len(), num_constants * n);
        assert_eq!(sigmas.len(), k_is.len() * n);
        let mut raw: *mut etp_circuit = std::ptr::null_mut();
        let digest_ptr = circuit_digest.as_ref().map_or(std::ptr::null(), |d| d.as_ptr());
        let rc = unsafe {
            etp_circuit_create(ctx.0, vanishing_program.as_ptr(), vanishing_program.len(), constants.as_ptr(), num_constants as c_int, sigmas.as_ptr(),
                               k_is.as_ptr(), k_is.len() as c_int, num_wires as c_int, degree_bits as c_int, quotient_degree_factor as c_int,
                               num_challenges as c_int, fri_params, digest_ptr, &mut raw)
        };
        ctx.check(rc);
        let proof_words = unsafe { etp_circuit_proof_words(raw) };
        DeviceCircuit { raw, proof_words }
    }

    /// `wires`: the witness, `num_wires x n` values column-major.  Returns the flat "B200PLK1" proof words
    /// (include/etp_b200.h; `ProofWithPublicInputs` is rebuilt from them field by field).
    pub fn prove(&self, ctx: &Ctx, wires: &[u64], public_inputs_hash: [u64; 4]) -> Vec<u64> {
        let mut out = vec![0u64; self.proof_words];
        let rc = unsafe { etp_circuit_prove_host(self.raw, wires.as_ptr(), public_inputs_hash.as_ptr(), out.as_mut_ptr()) };
        ctx.check(rc);
        out
    }

    pub fn digest(&self) -> [u64; 4] {
        let mut d = [0u64; 4];
        unsafe { etp_circuit_digest(self.raw, d.as_mut_ptr()) };
        d
    }
}
impl Drop for DeviceCircuit {
    fn drop(&mut self) {
        unsafe { etp_circuit_free(self.raw) }
    }
}
