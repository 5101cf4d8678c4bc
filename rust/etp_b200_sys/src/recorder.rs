//! Constraint-program recorder: runs a starky table's `eval_packed_generic` ONCE on symbolic values and writes
//! the straight-line program that `etp_table_register` compiles for sm_100a (wire format:
//! eth_tx_proof_b200/csrc/cprog.h; Python twin used by this repo's tests: eth_tx_proof_b200/cprog.py).
//!
//! NOT compiled in this repo (no Rust toolchain in the build image) — shipped as the source a maintainer adds
//! to the starky fork.  `Sym` is a `Copy` handle (u32 id) into a thread-local arena, so it satisfies the bounds
//! starky puts on `P: PackedField<Scalar = FE>`; the arena hash-conses, like the Python builder.
use std::cell::RefCell;
use std::collections::HashMap;
use std::ops::{Add, Mul, Neg, Sub};

pub const MAGIC: u64 = 0x3147525043505445; // "ETPCPRG1"
#[repr(u8)]
#[derive(Clone, Copy, PartialEq, Eq, Hash)]
pub enum Op { Const = 0, Lv, Nv, La, Na, Pi, Ch, Add, Sub, Mul, Emit, EmitTransition, EmitFirstRow, EmitLastRow }

#[derive(Default)]
struct Arena { ops: Vec<(u8, u32, u32, u64)>, memo: HashMap<(u8, u32, u32, u64), u32>, n_constraints: u32 }
thread_local!(static ARENA: RefCell<Arena> = RefCell::new(Arena::default()));

#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub struct Sym(pub u32);

fn push(op: Op, a: u32, b: u32, imm: u64, memoise: bool) -> Sym {
    ARENA.with(|ar| {
        let mut ar = ar.borrow_mut();
        let key = (op as u8, a, b, imm);
        if memoise { if let Some(&id) = ar.memo.get(&key) { return Sym(id); } }
        ar.ops.push(key);
        let id = (ar.ops.len() - 1) as u32;
        if memoise { ar.memo.insert(key, id); }
        Sym(id)
    })
}
impl Sym {
    pub fn constant(v: u64) -> Sym { push(Op::Const, 0, 0, v, true) }   // v canonical (< p)
    pub fn local(col: usize) -> Sym { push(Op::Lv, col as u32, 0, 0, true) }
    pub fn next(col: usize) -> Sym { push(Op::Nv, col as u32, 0, 0, true) }
    pub fn aux_local(col: usize) -> Sym { push(Op::La, col as u32, 0, 0, true) }
    pub fn aux_next(col: usize) -> Sym { push(Op::Na, col as u32, 0, 0, true) }
    pub fn public_input(i: usize) -> Sym { push(Op::Pi, i as u32, 0, 0, true) }
    pub fn challenge(i: usize) -> Sym { push(Op::Ch, i as u32, 0, 0, true) }
}
impl Add for Sym { type Output = Sym; fn add(self, o: Sym) -> Sym { push(Op::Add, self.0, o.0, 0, true) } }
impl Sub for Sym { type Output = Sym; fn sub(self, o: Sym) -> Sym { push(Op::Sub, self.0, o.0, 0, true) } }
impl Mul for Sym { type Output = Sym; fn mul(self, o: Sym) -> Sym { push(Op::Mul, self.0, o.0, 0, true) } }
impl Neg for Sym { type Output = Sym; fn neg(self) -> Sym { Sym::constant(0) - self } }
// ... AddAssign / SubAssign / MulAssign / Sum / Product / Mul<GoldilocksField> / From<GoldilocksField> follow the
// same pattern; `PackedField for Sym` sets WIDTH = 1 and maps `Self::ZEROS / ONES` to constants.

/// The recording `ConstraintConsumer`: same four methods as starky/src/constraint_consumer.rs, emission order kept.
pub struct RecordingConsumer;
impl RecordingConsumer {
    fn emit(op: Op, c: Sym) { push(op, c.0, 0, 0, false); ARENA.with(|a| a.borrow_mut().n_constraints += 1); }
    pub fn constraint(&mut self, c: Sym) { Self::emit(Op::Emit, c) }
    pub fn constraint_transition(&mut self, c: Sym) { Self::emit(Op::EmitTransition, c) }
    pub fn constraint_first_row(&mut self, c: Sym) { Self::emit(Op::EmitFirstRow, c) }
    pub fn constraint_last_row(&mut self, c: Sym) { Self::emit(Op::EmitLastRow, c) }
}

/// Serialises the arena.  Call after `stark.eval_packed_generic(&vars, &mut RecordingConsumer)` followed by
/// `eval_packed_lookups_generic` on symbolic `LookupCheckVars` (aux_local / aux_next / challenge handles).
pub fn finish(n_trace_cols: usize, n_aux_cols: usize, n_public_inputs: usize, n_challenges: usize,
              constraint_degree: usize) -> Vec<u64> {
    ARENA.with(|ar| {
        let ar = std::mem::take(&mut *ar.borrow_mut());
        let mut w = vec![MAGIC, ar.ops.len() as u64, n_trace_cols as u64, n_aux_cols as u64, n_public_inputs as u64,
                         n_challenges as u64, constraint_degree as u64, ar.n_constraints as u64];
        for (op, a, b, imm) in ar.ops {
            w.push(op as u64 | (a as u64) << 8 | (b as u64) << 36);
            w.push(imm);
        }
        w
    })
}

/// `Lookup { columns, table_column, frequencies_column, .. }` (single-column `Column`s, no filters) -> the flat
/// i32 description `etp_table_register` takes.
pub fn flatten_lookups(lookups: &[(Vec<usize>, usize, usize)]) -> Vec<i32> {
    let mut out = vec![lookups.len() as i32];
    for (looking, table, freq) in lookups {
        out.extend([*table as i32, *freq as i32, looking.len() as i32]);
        out.extend(looking.iter().map(|&c| c as i32));
    }
    out
}
