//! See Cargo.toml.  Mirrors tools/gen_golden.py case by case: COMMITS (PolynomialBatch::from_values on SplitMix64 columns) and
//! the Fibonacci PROOFS (starky::prover::prove under StarkConfig::standard_fast_config), hashing exactly the arrays the Python
//! side hashes (little-endian u64, C order).  The flat proof layout is "B200STK2" (DESIGN.md).
use std::marker::PhantomData;

use plonky2::field::extension::{Extendable, FieldExtension};
use plonky2::field::goldilocks_field::GoldilocksField;
use plonky2::field::packed::PackedField;
use plonky2::field::polynomial::PolynomialValues;
use plonky2::field::types::{Field, PrimeField64};
use plonky2::fri::oracle::PolynomialBatch;
use plonky2::hash::hash_types::{HashOut, RichField};
use plonky2::iop::ext_target::ExtensionTarget;
use plonky2::plonk::circuit_builder::CircuitBuilder;
use plonky2::plonk::config::{GenericConfig, PoseidonGoldilocksConfig};
use plonky2::util::timing::TimingTree;
use sha2::{Digest, Sha256};
use starky::config::StarkConfig;
use starky::constraint_consumer::{ConstraintConsumer, RecursiveConstraintConsumer};
use starky::evaluation_frame::{StarkEvaluationFrame, StarkFrame};
use starky::proof::StarkProofWithPublicInputs;
use starky::prover::prove;
use starky::stark::Stark;
use starky::verifier::verify_stark_proof;

const D: usize = 2;
type C = PoseidonGoldilocksConfig;
type F = <C as GenericConfig<D>>::F;
const P: u64 = 0xFFFF_FFFF_0000_0001;

// ---- eth_tx_proof_b200/synthetic.py: splitmix64, _rand, random_columns, fibonacci_trace -------------------------------------
fn splitmix64(x: u64) -> u64 {
    let mut z = x.wrapping_add(0x9E37_79B9_7F4A_7C15);
    z = (z ^ (z >> 30)).wrapping_mul(0xBF58_476D_1CE4_E5B9);
    z = (z ^ (z >> 27)).wrapping_mul(0x94D0_49BB_1331_11EB);
    z ^ (z >> 31)
}
fn rand_stream(seed: u64, stream: u64, n: usize) -> Vec<u64> {
    let base = splitmix64(seed.wrapping_mul(0x100_0003).wrapping_add(stream));
    (0..n as u64).map(|i| splitmix64(i.wrapping_mul(0x2545_F491_4F6C_DD1D).wrapping_add(base))).collect()
}
fn random_columns(n_cols: usize, log_n: usize, seed: u64) -> Vec<Vec<u64>> {
    (0..n_cols).map(|c| rand_stream(seed + c as u64, 0, 1 << log_n).into_iter().map(|x| if x >= P { x - P } else { x }).collect()).collect()
}
fn fibonacci_trace(log_n: usize, seed: u64) -> (Vec<Vec<u64>>, Vec<u64>) {
    let n = 1usize << log_n;
    let (mut x0, mut x1) = ((seed * 3 + 1) % P, (seed * 5 + 2) % P);
    let mut pi = vec![x0, x1];
    let mut t = vec![vec![0u64; n], vec![0u64; n]];
    for i in 0..n {
        t[0][i] = x0;
        t[1][i] = x1;
        let s = ((x0 as u128 + x1 as u128) % P as u128) as u64;
        x0 = x1;
        x1 = s;
    }
    pi.push(t[1][n - 1]);
    (t, pi)
}

fn sha(words: &[u64]) -> String {
    let mut h = Sha256::new();
    for w in words { h.update(w.to_le_bytes()); }
    h.finalize().iter().map(|b| format!("{b:02x}")).collect()
}
fn hash_words(h: &HashOut<F>) -> [u64; 4] { [0, 1, 2, 3].map(|i| h.elements[i].to_canonical_u64()) }

// ---- tools/gen_golden.py: commit_case ----------------------------------------------------------------------------------------
fn commit_case(n_cols: usize, log_n: usize, rate_bits: usize, cap_height: usize, seed: u64) -> serde_json::Value {
    let cols = random_columns(n_cols, log_n, seed);
    let values: Vec<PolynomialValues<F>> =
        cols.iter().map(|c| PolynomialValues::new(c.iter().map(|&x| F::from_canonical_u64(x)).collect())).collect();
    let mut timing = TimingTree::default();
    let b = PolynomialBatch::<F, C, D>::from_values(values, rate_bits, false, cap_height, &mut timing, None);
    let n_leaves = 1usize << (log_n + rate_bits);
    let cap: Vec<u64> = b.merkle_tree.cap.0.iter().flat_map(hash_words).collect();
    let coeffs: Vec<u64> = b.polynomials.iter().flat_map(|p| p.coeffs.iter().map(|x| x.to_canonical_u64())).collect();
    let digests: Vec<u64> = b.merkle_tree.digests.iter().flat_map(hash_words).collect();
    let mut idx = vec![0usize, 1, n_leaves / 3, n_leaves - 1];
    idx.sort_unstable();
    idx.dedup();
    let rows: Vec<u64> = idx.iter().flat_map(|&i| b.merkle_tree.leaves[i].iter().map(|x| x.to_canonical_u64())).collect();
    let paths: Vec<u64> = if n_leaves > (1 << cap_height) {
        idx.iter().flat_map(|&i| b.merkle_tree.prove(i).siblings.iter().flat_map(hash_words).collect::<Vec<u64>>()).collect()
    } else {
        vec![]
    };
    serde_json::json!({
        "shape": [n_cols, log_n, rate_bits, cap_height, seed], "cap_sha256": sha(&cap),
        "cap_first": cap[..4].iter().map(|x| format!("{x:016x}")).collect::<Vec<_>>(),
        "coeffs_sha256": sha(&coeffs), "digests_sha256": sha(&digests), "leaf_rows": idx, "leaf_rows_sha256": sha(&rows),
        "paths_sha256": sha(&paths),
    })
}

// ---- starky/src/fibonacci_stark.rs (cfg(test) upstream, restated here): 2 columns, public inputs [x0, x1, x1_last] -------------
#[derive(Copy, Clone)]
struct FibonacciStark<F: RichField + Extendable<D>, const D: usize> { _p: PhantomData<F> }
impl<F: RichField + Extendable<D>, const D: usize> Stark<F, D> for FibonacciStark<F, D> {
    type EvaluationFrame<FE, P, const D2: usize> = StarkFrame<P, P::Scalar, 2, 3> where FE: FieldExtension<D2, BaseField = F>, P: PackedField<Scalar = FE>;
    type EvaluationFrameTarget = StarkFrame<ExtensionTarget<D>, ExtensionTarget<D>, 2, 3>;
    fn eval_packed_generic<FE, P, const D2: usize>(&self, vars: &Self::EvaluationFrame<FE, P, D2>, yield_constr: &mut ConstraintConsumer<P>)
    where FE: FieldExtension<D2, BaseField = F>, P: PackedField<Scalar = FE> {
        let (lv, nv, pi) = (vars.get_local_values(), vars.get_next_values(), vars.get_public_inputs());
        yield_constr.constraint_first_row(lv[0] - pi[0]);
        yield_constr.constraint_first_row(lv[1] - pi[1]);
        yield_constr.constraint_last_row(lv[1] - pi[2]);
        yield_constr.constraint_transition(nv[0] - lv[1]);
        yield_constr.constraint_transition(nv[1] - lv[0] - lv[1]);
    }
    fn eval_ext_circuit(&self, _b: &mut CircuitBuilder<F, D>, _v: &Self::EvaluationFrameTarget, _y: &mut RecursiveConstraintConsumer<F, D>) {
        unimplemented!("not needed by the prover")
    }
    fn constraint_degree(&self) -> usize { 2 }
}

fn ext_words(e: &<F as Extendable<D>>::Extension) -> [u64; 2] {
    let a: [F; 2] = e.to_basefield_array();
    [a[0].to_canonical_u64(), a[1].to_canonical_u64()]
}
/// StarkProofWithPublicInputs -> the flat "B200STK2" words (no auxiliary polynomials / CTL in these cases)
fn flatten(table: u64, degree_bits: usize, config: &StarkConfig, p: &StarkProofWithPublicInputs<F, C, D>) -> Vec<u64> {
    let pr = &p.proof;
    let fri = &pr.opening_proof;
    let n_trace = pr.openings.local_values.len();
    let n_quot = pr.openings.quotient_polys.as_ref().map_or(0, |q| q.len());
    let mut w = vec![0u64; 24];
    let push_cap = |w: &mut Vec<u64>, cap: &plonky2::hash::merkle_tree::MerkleCap<F, <C as GenericConfig<D>>::Hasher>| {
        for h in &cap.0 { w.extend(hash_words(h)); }
    };
    push_cap(&mut w, &pr.trace_cap);
    assert!(pr.auxiliary_polys_cap.is_none());
    push_cap(&mut w, pr.quotient_polys_cap.as_ref().unwrap());
    for e in &pr.openings.local_values { w.extend(ext_words(e)); }
    for e in &pr.openings.next_values { w.extend(ext_words(e)); }
    for e in pr.openings.quotient_polys.as_ref().unwrap() { w.extend(ext_words(e)); }
    for cap in &fri.commit_phase_merkle_caps { push_cap(&mut w, cap); }
    for round in &fri.query_round_proofs {
        for (evals, proof) in &round.initial_trees_proof.evals_proofs {
            w.extend(evals.iter().map(|x| x.to_canonical_u64()));
            for h in &proof.siblings { w.extend(hash_words(h)); }
        }
        for step in &round.steps {
            for e in &step.evals { w.extend(ext_words(e)); }
            for h in &step.merkle_proof.siblings { w.extend(hash_words(h)); }
        }
    }
    for e in &fri.final_poly.coeffs { w.extend(ext_words(e)); }
    w.push(fri.pow_witness.to_canonical_u64());
    w.extend(p.public_inputs.iter().map(|x| x.to_canonical_u64()));
    let fc = &config.fri_config;
    let hdr = [0x4232_3030_5354_4B32u64, table, degree_bits as u64, n_trace as u64, 0, n_quot as u64, fc.cap_height as u64,
               fri.commit_phase_merkle_caps.len() as u64, 4, fri.final_poly.coeffs.len() as u64, fc.num_query_rounds as u64,
               p.public_inputs.len() as u64, fc.rate_bits as u64, fc.proof_of_work_bits as u64, config.num_challenges as u64, w.len() as u64];
    w[..16].copy_from_slice(&hdr);
    w
}

fn fibonacci_case(log_n: usize, seed: u64) -> anyhow::Result<serde_json::Value> {
    let (t, pi) = fibonacci_trace(log_n, seed);
    let trace: Vec<PolynomialValues<F>> = t.iter().map(|c| PolynomialValues::new(c.iter().map(|&x| F::from_canonical_u64(x)).collect())).collect();
    let pis: Vec<F> = pi.iter().map(|&x| F::from_canonical_u64(x)).collect();
    let config = StarkConfig::standard_fast_config();
    let stark = FibonacciStark::<F, D> { _p: PhantomData };
    let proof = prove::<F, C, _, D>(stark, &config, trace, &pis, &mut TimingTree::default())?;
    verify_stark_proof(stark, proof.clone(), &config)?;
    let words = flatten(0, log_n, &config, &proof);
    let mut without_id = words.clone();
    without_id.remove(1);
    Ok(serde_json::json!({
        "table": "fibonacci", "log_n": log_n, "seed": seed, "words": words.len(), "proof_sha256_without_table_id": sha(&without_id),
        "pow_witness": format!("{:016x}", proof.proof.opening_proof.pow_witness.to_canonical_u64()),
    }))
}

fn main() -> anyhow::Result<()> {
    // tools/gen_golden.py: COMMITS and the fibonacci entries of PROOFS
    let commits = [(3, 4, 1, 2, 101u64), (9, 5, 1, 4, 102), (4, 6, 1, 4, 103), (21, 8, 1, 4, 104), (128, 10, 1, 4, 105), (17, 7, 2, 3, 106), (135, 6, 1, 0, 107)];
    let proofs = [(5usize, 1u64), (9, 2)];
    let out = serde_json::json!({
        "note": "REAL plonky2 0.2.2 / starky 0.4.0 outputs on the seeded inputs of tools/gen_golden.py (rust/parity_dump)",
        "commits": commits.iter().map(|&(c, l, r, h, s)| commit_case(c, l, r, h, s)).collect::<Vec<_>>(),
        "proofs": proofs.iter().map(|&(l, s)| fibonacci_case(l, s)).collect::<anyhow::Result<Vec<_>>>()?,
    });
    println!("{}", serde_json::to_string_pretty(&out)?);
    Ok(())
}
