//! Constraint-program recorder: runs a starky table's UNMODIFIED `eval_packed_generic` (+ `eval_packed_lookups_generic`,
//! `eval_cross_table_lookup_checks`) ONCE on symbolic values and writes the straight-line program that
//! `etp_table_register_ex` compiles for sm_100a (wire format: eth_tx_proof_b200/csrc/cprog.h; Python twin used by this
//! repo's tests: eth_tx_proof_b200/cprog.py), plus the auxiliary-column spec (lookups, CTL Z descriptors).
//!
//! How it plugs into starky 0.4.0 without touching the tables (evm_arithmetization's seven STARKs included): every table
//! is generic over `FE: FieldExtension<D2, BaseField = F>, P: PackedField<Scalar = FE>`.  `Sym` below is a `Copy` handle
//! into a thread-local hash-consing arena that implements `Field` and `FieldExtension<2, BaseField = GoldilocksField>`
//! (it never needs the extension's structure: evaluators only use ring operations and constants), so plonky2_field's
//! blanket `unsafe impl<F: Field> PackedField for F` makes it a valid `P` with `WIDTH = 1`.  The stock
//! `ConstraintConsumer<Sym>` is fed ONE symbolic alpha and marker values for `z_last` / the Lagrange selectors; the
//! emission order and kinds are then read back from the shape of its accumulator, acc_k = acc_{k-1} * ALPHA + c_k * marker.
//!
//! NOT compiled in this repo (no Rust toolchain in the build image).  Written against plonky2_field 0.2.2 / starky 0.4.0
//! (/root/reference/Cargo.lock:3466,4529) as published; every trait item those versions require is spelled out below.
use std::cell::RefCell;
use std::collections::HashMap;
use std::fmt;
use std::iter::{Product, Sum};
use std::ops::{Add, AddAssign, Div, DivAssign, Mul, MulAssign, Neg, Sub, SubAssign};

use num::BigUint;
use plonky2_field::extension::FieldExtension;
use plonky2_field::goldilocks_field::GoldilocksField;
use plonky2_field::ops::Square;
use plonky2_field::types::{Field, PrimeField64, Sample};
use serde::{Deserialize, Serialize};

pub const MAGIC: u64 = 0x3147525043505445; // "ETPCPRG1"
pub const AUXSPEC_MAGIC: u64 = 0x3153585541505445; // "ETPAUXS1"
const P: u64 = 0xFFFF_FFFF_0000_0001;

#[repr(u8)]
#[derive(Clone, Copy, PartialEq, Eq, Hash, Debug)]
pub enum Op { Const = 0, Lv, Nv, La, Na, Pi, Ch, Add, Sub, Mul, Emit, EmitTransition, EmitFirstRow, EmitLastRow }

/// Arena node: (opcode, a, b, immediate).  Ids 0..RESERVED are pre-seeded so that `Field`'s associated constants and the
/// consumer markers are `const`-constructible handles.
type Node = (u8, u32, u32, u64);
const ID_ZERO: u32 = 0;
const ID_ONE: u32 = 1;
const ID_TWO: u32 = 2;
const ID_NEG_ONE: u32 = 3;
const ID_SEVEN: u32 = 4; // MULTIPLICATIVE_GROUP_GENERATOR (never used by an evaluator; present for completeness)
const ID_ALPHA: u32 = 5; // markers: opaque values handed to ConstraintConsumer::new
const ID_Z_LAST: u32 = 6;
const ID_L_FIRST: u32 = 7;
const ID_L_LAST: u32 = 8;
const MARKER_OP: u8 = 0xFF; // not a program opcode: markers never reach the output

struct Arena { ops: Vec<Node>, memo: HashMap<Node, u32> }
impl Arena {
    fn new() -> Self {
        let mut a = Arena { ops: Vec::new(), memo: HashMap::new() };
        for (id, node) in [(ID_ZERO, (Op::Const as u8, 0, 0, 0u64)), (ID_ONE, (Op::Const as u8, 0, 0, 1)), (ID_TWO, (Op::Const as u8, 0, 0, 2)),
                           (ID_NEG_ONE, (Op::Const as u8, 0, 0, P - 1)), (ID_SEVEN, (Op::Const as u8, 0, 0, 7)),
                           (ID_ALPHA, (MARKER_OP, 0, 0, 0)), (ID_Z_LAST, (MARKER_OP, 1, 0, 0)), (ID_L_FIRST, (MARKER_OP, 2, 0, 0)),
                           (ID_L_LAST, (MARKER_OP, 3, 0, 0))] {
            debug_assert_eq!(a.ops.len() as u32, id);
            a.ops.push(node);
            a.memo.insert(node, id);
        }
        a
    }
}
thread_local!(static ARENA: RefCell<Arena> = RefCell::new(Arena::new()));

/// Start a new recording on this thread.
pub fn reset() { ARENA.with(|a| *a.borrow_mut() = Arena::new()); }

fn push(op: u8, a: u32, b: u32, imm: u64) -> Sym {
    ARENA.with(|ar| {
        let mut ar = ar.borrow_mut();
        let key = (op, a, b, imm);
        if let Some(&id) = ar.memo.get(&key) { return Sym(id); }
        ar.ops.push(key);
        let id = (ar.ops.len() - 1) as u32;
        ar.memo.insert(key, id);
        Sym(id)
    })
}
fn node(id: u32) -> Node { ARENA.with(|ar| ar.borrow().ops[id as usize]) }

#[derive(Clone, Copy, PartialEq, Eq, Hash, Default, Serialize, Deserialize)]
pub struct Sym(pub u32);

impl Sym {
    pub fn constant(v: u64) -> Sym { push(Op::Const as u8, 0, 0, v % P) }
    pub fn local(col: usize) -> Sym { push(Op::Lv as u8, col as u32, 0, 0) }
    pub fn next(col: usize) -> Sym { push(Op::Nv as u8, col as u32, 0, 0) }
    pub fn aux_local(col: usize) -> Sym { push(Op::La as u8, col as u32, 0, 0) }
    pub fn aux_next(col: usize) -> Sym { push(Op::Na as u8, col as u32, 0, 0) }
    pub fn public_input(i: usize) -> Sym { push(Op::Pi as u8, i as u32, 0, 0) }
    /// challenge scalars: 0..num_challenges = lookup challenges, then CTL (beta_k, gamma_k) at num_challenges + 2k, + 2k + 1
    pub fn challenge(i: usize) -> Sym { push(Op::Ch as u8, i as u32, 0, 0) }
    pub const ALPHA: Sym = Sym(ID_ALPHA);
    pub const Z_LAST: Sym = Sym(ID_Z_LAST);
    pub const LAGRANGE_FIRST: Sym = Sym(ID_L_FIRST);
    pub const LAGRANGE_LAST: Sym = Sym(ID_L_LAST);
    fn const_value(self) -> Option<u64> { let n = node(self.0); if n.0 == Op::Const as u8 { Some(n.3) } else { None } }
}
impl fmt::Debug for Sym { fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result { write!(f, "v{}", self.0) } }
impl fmt::Display for Sym { fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result { write!(f, "v{}", self.0) } }

impl Add for Sym { type Output = Sym; fn add(self, o: Sym) -> Sym { push(Op::Add as u8, self.0, o.0, 0) } }
impl Sub for Sym { type Output = Sym; fn sub(self, o: Sym) -> Sym { push(Op::Sub as u8, self.0, o.0, 0) } }
impl Mul for Sym { type Output = Sym; fn mul(self, o: Sym) -> Sym { push(Op::Mul as u8, self.0, o.0, 0) } }
impl Neg for Sym { type Output = Sym; fn neg(self) -> Sym { Sym(ID_ZERO) - self } }
/// Division only by constants (evaluators divide by small integers at most: multiply by the constant's inverse).
impl Div for Sym {
    type Output = Sym;
    fn div(self, o: Sym) -> Sym {
        let c = o.const_value().expect("symbolic recording: division by a non-constant");
        self * Sym::constant(GoldilocksField::from_canonical_u64(c).inverse().to_canonical_u64())
    }
}
impl AddAssign for Sym { fn add_assign(&mut self, o: Sym) { *self = *self + o; } }
impl SubAssign for Sym { fn sub_assign(&mut self, o: Sym) { *self = *self - o; } }
impl MulAssign for Sym { fn mul_assign(&mut self, o: Sym) { *self = *self * o; } }
impl DivAssign for Sym { fn div_assign(&mut self, o: Sym) { *self = *self / o; } }
impl Sum for Sym { fn sum<I: Iterator<Item = Sym>>(iter: I) -> Sym { iter.fold(Sym(ID_ZERO), |a, b| a + b) } }
impl Product for Sym { fn product<I: Iterator<Item = Sym>>(iter: I) -> Sym { iter.fold(Sym(ID_ONE), |a, b| a * b) } }
impl Square for Sym { fn square(&self) -> Sym { *self * *self } }
impl Sample for Sym {
    fn sample<R>(_rng: &mut R) -> Self where R: rand::RngCore + ?Sized { unimplemented!("symbolic values cannot be sampled") }
}

impl Field for Sym {
    const ZERO: Self = Sym(ID_ZERO);
    const ONE: Self = Sym(ID_ONE);
    const TWO: Self = Sym(ID_TWO);
    const NEG_ONE: Self = Sym(ID_NEG_ONE);
    const TWO_ADICITY: usize = 32;
    const CHARACTERISTIC_TWO_ADICITY: usize = 32;
    const MULTIPLICATIVE_GROUP_GENERATOR: Self = Sym(ID_SEVEN);
    const POWER_OF_TWO_GENERATOR: Self = Sym(ID_SEVEN); // never read while recording
    const BITS: usize = 64;
    fn order() -> BigUint { GoldilocksField::order() }
    fn characteristic() -> BigUint { GoldilocksField::characteristic() }
    fn try_inverse(&self) -> Option<Self> {
        self.const_value().and_then(|c| GoldilocksField::from_canonical_u64(c).try_inverse()).map(|x| Sym::constant(x.to_canonical_u64()))
    }
    fn from_noncanonical_biguint(n: BigUint) -> Self { Sym::constant(GoldilocksField::from_noncanonical_biguint(n).to_canonical_u64()) }
    fn from_canonical_u64(n: u64) -> Self { Sym::constant(n) }
    fn from_noncanonical_u128(n: u128) -> Self { Sym::constant(GoldilocksField::from_noncanonical_u128(n).to_canonical_u64()) }
    fn from_noncanonical_u64(n: u64) -> Self { Sym::constant(n % P) }
    fn from_noncanonical_i64(n: i64) -> Self { Sym::constant(GoldilocksField::from_noncanonical_i64(n).to_canonical_u64()) }
    fn from_noncanonical_u96(n: (u64, u32)) -> Self { Sym::constant(GoldilocksField::from_noncanonical_u96(n).to_canonical_u64()) }
}

/// "Extension of degree 2 over Goldilocks" as far as the type system is concerned; the structure is never used.
impl FieldExtension<2> for Sym {
    type BaseField = GoldilocksField;
    fn to_basefield_array(&self) -> [GoldilocksField; 2] {
        [GoldilocksField::from_canonical_u64(self.const_value().expect("symbolic value has no base-field coordinates")), GoldilocksField::ZERO]
    }
    fn from_basefield_array(arr: [GoldilocksField; 2]) -> Self {
        assert!(arr[1] == GoldilocksField::ZERO, "recorded programs are over the base field");
        Sym::constant(arr[0].to_canonical_u64())
    }
    fn from_basefield(x: GoldilocksField) -> Self { Sym::constant(x.to_canonical_u64()) }
    fn scalar_mul(&self, scalar: GoldilocksField) -> Self { *self * Sym::constant(scalar.to_canonical_u64()) }
}

/// What the stock consumer is constructed with: `ConstraintConsumer::new(vec![Sym::ALPHA], Sym::Z_LAST, Sym::LAGRANGE_FIRST,
/// Sym::LAGRANGE_LAST)`; after the evaluators ran, pass `consumer.accumulators()[0]` to `finish`.
pub fn consumer_args() -> (Vec<Sym>, Sym, Sym, Sym) { (vec![Sym::ALPHA], Sym::Z_LAST, Sym::LAGRANGE_FIRST, Sym::LAGRANGE_LAST) }

/// Unwinds acc = (..((0 * A + c_0 m_0) * A + c_1 m_1) ..) * A + c_{n-1} m_{n-1} into [(kind, constraint)] in emission order.
fn unwind(acc: Sym) -> Vec<(Op, Sym)> {
    let mut out = Vec::new();
    let mut cur = acc;
    while cur.0 != ID_ZERO {
        let (op, a, b, _) = node(cur.0);
        assert_eq!(op, Op::Add as u8, "accumulator is not of the form acc * alpha + constraint");
        let (mop, ma, mb, _) = node(a);
        assert!(mop == Op::Mul as u8 && mb == ID_ALPHA, "accumulator is not of the form acc * alpha + constraint");
        // the emitted value: c (plain), c * Z_LAST, c * LAGRANGE_FIRST or c * LAGRANGE_LAST
        let (cop, ca, cb, _) = node(b);
        let emitted = if cop == Op::Mul as u8 && cb == ID_Z_LAST { (Op::EmitTransition, Sym(ca)) }
            else if cop == Op::Mul as u8 && cb == ID_L_FIRST { (Op::EmitFirstRow, Sym(ca)) }
            else if cop == Op::Mul as u8 && cb == ID_L_LAST { (Op::EmitLastRow, Sym(ca)) }
            else { (Op::Emit, Sym(b)) };
        out.push(emitted);
        cur = Sym(ma);
    }
    out.reverse();
    out
}

/// Serialises the recording: the values the constraints depend on, in topological (= arena) order, then the EMIT ops in
/// emission order.  `acc` = the consumer's accumulator for the single symbolic alpha.
pub fn finish(acc: Sym, n_trace_cols: usize, n_aux_cols: usize, n_public_inputs: usize, n_challenge_scalars: usize,
              constraint_degree: usize) -> Vec<u64> {
    finish_constraints(unwind(acc), n_trace_cols, n_aux_cols, n_public_inputs, n_challenge_scalars, constraint_degree)
}

/// The same for an explicit constraint list in emission order — the plonky2 circuit prover's vanishing polynomial
/// (`plonk/vanishing_poly.rs`), whose terms are a `Vec` reduced with `reduce_with_powers` rather than folded by a consumer:
/// pass them REVERSED (term i is weighted by alpha^i there, emission i of N by alpha^(N-1-i) here), the `L_0(x)(Z - 1)` terms
/// as `(Op::EmitFirstRow, Z - 1)`.
pub fn finish_constraints(constraints: Vec<(Op, Sym)>, n_trace_cols: usize, n_aux_cols: usize, n_public_inputs: usize,
                          n_challenge_scalars: usize, constraint_degree: usize) -> Vec<u64> {
    // mark the sub-DAG reachable from the constraints (markers and accumulator nodes are not part of it)
    let n_nodes = ARENA.with(|a| a.borrow().ops.len());
    let mut live = vec![false; n_nodes];
    let mut stack: Vec<u32> = constraints.iter().map(|(_, c)| c.0).collect();
    while let Some(id) = stack.pop() {
        if live[id as usize] { continue; }
        live[id as usize] = true;
        let (op, a, b, _) = node(id);
        assert_ne!(op, MARKER_OP, "a consumer marker leaked into a constraint");
        if op == Op::Add as u8 || op == Op::Sub as u8 || op == Op::Mul as u8 { stack.push(a); stack.push(b); }
    }
    let mut remap = vec![u32::MAX; n_nodes];
    let mut ops: Vec<(u8, u32, u32, u64)> = Vec::new();
    for id in 0..n_nodes {
        if !live[id] { continue; }
        let (op, a, b, imm) = node(id as u32);
        let (a, b) = if op == Op::Add as u8 || op == Op::Sub as u8 || op == Op::Mul as u8 { (remap[a as usize], remap[b as usize]) } else { (a, b) };
        remap[id] = ops.len() as u32;
        ops.push((op, a, b, imm));
    }
    for (kind, c) in &constraints { ops.push((*kind as u8, remap[c.0 as usize], 0, 0)); }
    let mut w = vec![MAGIC, ops.len() as u64, n_trace_cols as u64, n_aux_cols as u64, n_public_inputs as u64, n_challenge_scalars as u64,
                     constraint_degree as u64, constraints.len() as u64];
    for (op, a, b, imm) in ops {
        w.push(op as u64 | (a as u64) << 8 | (b as u64) << 36);
        w.push(imm);
    }
    w
}

/// starky::lookup::Column as data: (local terms, next-row terms, constant).
#[derive(Clone, Debug, Default)]
pub struct ColumnSpec { pub local: Vec<(usize, u64)>, pub next_row: Vec<(usize, u64)>, pub constant: u64 }
/// starky::lookup::Filter as data.
#[derive(Clone, Debug, Default)]
pub struct FilterSpec { pub products: Vec<(ColumnSpec, ColumnSpec)>, pub constants: Vec<ColumnSpec> }
impl FilterSpec {
    /// Filter::default(): evaluates to 1
    pub fn one() -> Self { FilterSpec { products: vec![], constants: vec![ColumnSpec { local: vec![], next_row: vec![], constant: 1 }] } }
}

/// Builder of the auxiliary-column spec words of `etp_table_register_ex` (layout: include/etp_b200.h).
#[derive(Default)]
pub struct AuxSpecBuilder { lookups: Vec<Vec<u64>>, zs: Vec<Vec<u64>> }
impl AuxSpecBuilder {
    fn column(w: &mut Vec<u64>, c: &ColumnSpec) {
        w.push(c.local.len() as u64);
        for (col, f) in &c.local { w.push(*col as u64); w.push(*f); }
        w.push(c.next_row.len() as u64);
        for (col, f) in &c.next_row { w.push(*col as u64); w.push(*f); }
        w.push(c.constant);
    }
    fn filter(w: &mut Vec<u64>, f: &FilterSpec) {
        w.push(f.products.len() as u64);
        for (a, b) in &f.products { Self::column(w, a); Self::column(w, b); }
        w.push(f.constants.len() as u64);
        for c in &f.constants { Self::column(w, c); }
    }
    /// one `Lookup { columns, table_column, frequencies_column, filter_columns }`
    pub fn lookup(&mut self, columns: &[ColumnSpec], filters: &[FilterSpec], table: &ColumnSpec, frequencies: &ColumnSpec) {
        assert_eq!(columns.len(), filters.len());
        let mut w = vec![columns.len() as u64];
        for c in columns { Self::column(&mut w, c); }
        for f in filters { Self::filter(&mut w, f); }
        Self::column(&mut w, table);
        Self::column(&mut w, frequencies);
        self.lookups.push(w);
    }
    /// one `CtlZData` of this table, in `CtlData.zs_columns` order; `colsets` = its (columns, filter) pairs
    pub fn ctl_z(&mut self, challenge_index: usize, colsets: &[(Vec<ColumnSpec>, FilterSpec)]) {
        let mut w = vec![challenge_index as u64, colsets.len() as u64];
        for (cols, filt) in colsets {
            w.push(cols.len() as u64);
            for c in cols { Self::column(&mut w, c); }
            Self::filter(&mut w, filt);
        }
        self.zs.push(w);
    }
    pub fn finish(self) -> Vec<u64> {
        let mut w = vec![AUXSPEC_MAGIC, self.lookups.len() as u64, self.zs.len() as u64];
        for l in self.lookups { w.extend(l); }
        for z in self.zs { w.extend(z); }
        w
    }
}

/// In the starky fork (starky/src/vanishing_poly.rs is `pub(crate)`, so this function lives there):
///
/// ```ignore
/// pub fn record_program<S: Stark<GoldilocksField, 2>>(stark: &S, config: &StarkConfig, ctl_zs: &[CtlZData<GoldilocksField>]) -> Vec<u64> {
///     use etp_b200_sys::recorder::{self as rec, Sym};
///     rec::reset();
///     let lv: Vec<Sym> = (0..S::COLUMNS).map(Sym::local).collect();
///     let nv: Vec<Sym> = (0..S::COLUMNS).map(Sym::next).collect();
///     let pis: Vec<Sym> = (0..S::PUBLIC_INPUTS).map(Sym::public_input).collect();
///     let vars = S::EvaluationFrame::<Sym, Sym, 2>::from_values(&lv, &nv, &pis);
///     let lookups = stark.lookups();
///     let n_lookup = stark.num_lookup_helper_columns(config);
///     let n_helpers: usize = ctl_zs.iter().map(|z| z.helper_columns.len()).sum();
///     let n_aux = n_lookup + n_helpers + ctl_zs.len();
///     let lookup_vars = stark.uses_lookups().then(|| LookupCheckVars {
///         local_values: (0..n_lookup).map(Sym::aux_local).collect(),
///         next_values: (0..n_lookup).map(Sym::aux_next).collect(),
///         challenges: (0..config.num_challenges).map(|k| /* F-typed in 0.4.0: the fork widens this field to FE */ Sym::challenge(k)).collect(),
///     });
///     let mut h = n_lookup;
///     let ctl_vars: Vec<CtlCheckVars<_, Sym, Sym, 2>> = ctl_zs.iter().enumerate().map(|(i, z)| {
///         let k = /* index of z.challenge in the GrandProductChallengeSet */ i % config.num_challenges;
///         let helpers = (h..h + z.helper_columns.len()).map(Sym::aux_local).collect(); h += z.helper_columns.len();
///         CtlCheckVars { helper_columns: helpers, local_z: Sym::aux_local(n_lookup + n_helpers + i), next_z: Sym::aux_next(n_lookup + n_helpers + i),
///                        challenges: GrandProductChallenge { beta: Sym::challenge(config.num_challenges + 2 * k), gamma: Sym::challenge(config.num_challenges + 2 * k + 1) },
///                        columns: z.columns.clone(), filter: z.filter.clone() }
///     }).collect();
///     let (alphas, z_last, l_first, l_last) = rec::consumer_args();
///     let mut consumer = ConstraintConsumer::<Sym>::new(alphas, z_last, l_first, l_last);
///     eval_vanishing_poly::<GoldilocksField, Sym, Sym, S, 2, 2>(stark, &vars, &lookups, lookup_vars, (!ctl_vars.is_empty()).then_some(&ctl_vars[..]), &mut consumer);
///     rec::finish(consumer.accumulators()[0], S::COLUMNS, n_aux, S::PUBLIC_INPUTS,
///                 if ctl_zs.is_empty() { config.num_challenges } else { 3 * config.num_challenges }, stark.constraint_degree())
/// }
/// ```
/// (`GrandProductChallenge<F>` / `LookupCheckVars.challenges: Vec<F>` carry base-field challenges in 0.4.0; the fork makes those
/// two fields generic over the evaluation field so that symbolic challenges can flow through `combine` — a two-line change in
/// starky/src/lookup.rs.)
pub const FORK_RECORDING_RECIPE: () = ();
