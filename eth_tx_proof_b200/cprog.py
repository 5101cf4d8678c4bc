"""Constraint programs: a starky table's ``eval_packed_generic`` + lookup checks as straight-line SSA code.

Wire format: eth_tx_proof_b200/csrc/cprog.h.  In production the patched starky records a program per table by
running the table's evaluator on a symbolic PackedField (rust/etp_b200_sys); this module is the same recorder
for Python-described tables (tests, benches, synthetic tables of the evm_arithmetization shapes the reference
proves through /root/reference/ops/src/lib.rs:52).  ``ProgramBuilder`` hash-conses expressions, so shared
sub-expressions are evaluated once; constraints are emitted in call order, which is the order the
ConstraintConsumer folds them with the alphas.

Host side only: nothing here touches the GPU or the oracle.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

P = 0xFFFFFFFF00000001
MAGIC = 0x3147525043505445  # "ETPCPRG1"
(CONST, LV, NV, LA, NA, PI, CH, ADD, SUB, MUL, EMIT, EMIT_TRANSITION, EMIT_FIRST_ROW, EMIT_LAST_ROW) = range(14)
NUM_CHALLENGES = 2  # StarkConfig::standard_fast_config().num_challenges


class Expr:
    __slots__ = ("b", "id")

    def __init__(self, b, i):
        self.b, self.id = b, i

    def _lift(self, o):
        return o if isinstance(o, Expr) else self.b.const(o)

    def __add__(self, o):
        return self.b._op(ADD, self.id, self._lift(o).id)

    __radd__ = __add__

    def __sub__(self, o):
        return self.b._op(SUB, self.id, self._lift(o).id)

    def __rsub__(self, o):
        return self.b._op(SUB, self._lift(o).id, self.id)

    def __mul__(self, o):
        return self.b._op(MUL, self.id, self._lift(o).id)

    __rmul__ = __mul__


class ProgramBuilder:
    def __init__(self, n_trace_cols: int, n_public_inputs: int = 0, constraint_degree: int = 3):
        self.n_trace, self.n_pi, self.degree = n_trace_cols, n_public_inputs, constraint_degree
        self.n_aux = 0
        self.n_ch = 0
        self.ops: List[Tuple[int, int, int, int]] = []  # (opcode, a, b, imm)
        self._memo = {}
        self.n_constraints = 0
        self.lookups: List[Tuple[List[int], int, int]] = []

    def _op(self, op, a=0, b=0, imm=0) -> Expr:
        key = (op, a, b, imm)
        if key in self._memo:
            return Expr(self, self._memo[key])
        self.ops.append(key)
        self._memo[key] = len(self.ops) - 1
        return Expr(self, len(self.ops) - 1)

    def const(self, v: int) -> Expr:
        return self._op(CONST, imm=int(v) % P)

    def lv(self, c: int) -> Expr:
        assert 0 <= c < self.n_trace
        return self._op(LV, c)

    def nv(self, c: int) -> Expr:
        assert 0 <= c < self.n_trace
        return self._op(NV, c)

    def la(self, c: int) -> Expr:
        self.n_aux = max(self.n_aux, c + 1)
        return self._op(LA, c)

    def na(self, c: int) -> Expr:
        self.n_aux = max(self.n_aux, c + 1)
        return self._op(NA, c)

    def pi(self, i: int) -> Expr:
        assert 0 <= i < self.n_pi
        return self._op(PI, i)

    def challenge(self, i: int) -> Expr:
        self.n_ch = max(self.n_ch, i + 1)
        return self._op(CH, i)

    def _emit(self, op, e: Expr):
        self.ops.append((op, e.id, 0, 0))  # never memoised: every constraint is folded, duplicates included
        self.n_constraints += 1

    def constraint(self, e: Expr):
        self._emit(EMIT, e)

    def transition(self, e: Expr):
        self._emit(EMIT_TRANSITION, e)

    def first_row(self, e: Expr):
        self._emit(EMIT_FIRST_ROW, e)

    def last_row(self, e: Expr):
        self._emit(EMIT_LAST_ROW, e)

    # ---- starky::lookup::eval_packed_lookups_generic (no filters) ------------------------------------------
    def add_lookup(self, looking: Sequence[int], table_col: int, freq_col: int):
        self.lookups.append((list(looking), table_col, freq_col))

    def emit_lookup_constraints(self):
        """Call once, after the table's own constraints.  Auxiliary columns: per lookup, per challenge: one helper
        column per chunk of (degree - 1) looking columns, then Z."""
        chunk = max(1, self.degree - 1)
        start = 0
        for looking, table_col, freq_col in self.lookups:
            n_help = -(-len(looking) // chunk)
            for k in range(NUM_CHALLENGES):
                ch = self.challenge(k)
                helpers = []
                for c in range(n_help):
                    cols = [self.lv(j) + ch for j in looking[c * chunk:(c + 1) * chunk]]
                    h = self.la(start + c)
                    helpers.append(h)
                    # eval_helper_columns: h * prod(col_j + ch) - sum_j prod_{i != j}(col_i + ch)
                    prod = cols[0]
                    for x in cols[1:]:
                        prod = prod * x
                    if len(cols) == 1:
                        rhs = self.const(1)
                    else:
                        rhs = None
                        for j in range(len(cols)):
                            t = None
                            for i, x in enumerate(cols):
                                if i != j:
                                    t = x if t is None else t * x
                            rhs = t if rhs is None else rhs + t
                    self.constraint(h * prod - rhs)
                z, next_z = self.la(start + n_help), self.na(start + n_help)
                twc = self.lv(table_col) + ch
                hs = helpers[0]
                for h in helpers[1:]:
                    hs = hs + h
                y = hs * twc - self.lv(freq_col)
                self.first_row(z)
                self.constraint((next_z - z) * twc - y)
                start += n_help + 1
        self.n_aux = max(self.n_aux, start)

    def num_aux_columns(self) -> int:
        chunk = max(1, self.degree - 1)
        return sum(-(-len(l[0]) // chunk) + 1 for l in self.lookups) * NUM_CHALLENGES

    def build(self) -> "Program":
        return Program(self)


class Program:
    def __init__(self, b: ProgramBuilder):
        self.n_trace, self.n_aux, self.n_pi, self.n_ch = b.n_trace, b.n_aux, b.n_pi, b.n_ch
        self.degree, self.n_constraints = b.degree, b.n_constraints
        self.lookups = [(list(l), t, f) for l, t, f in b.lookups]
        self.ops = list(b.ops)
        w = np.zeros(8 + 2 * len(self.ops), dtype=np.uint64)
        w[:8] = [MAGIC, len(self.ops), self.n_trace, self.n_aux, self.n_pi, self.n_ch, self.degree, self.n_constraints]
        for k, (op, a, bb, imm) in enumerate(self.ops):
            w[8 + 2 * k] = op | (a << 8) | (bb << 36)
            w[9 + 2 * k] = imm
        self.words = w

    def evaluate(self, lv, nv, la=(), na=(), pi=(), ch=(), add=None, sub=None, mul=None, lift=None):
        """Interprets the program over any ring (default: Python ints mod p).  Returns [(kind, value)] in emission
        order, kind in {EMIT, EMIT_TRANSITION, EMIT_FIRST_ROW, EMIT_LAST_ROW}.  Used by the Python verifier over the
        extension field and by trace self-checks."""
        add = add or (lambda x, y: (x + y) % P)
        sub = sub or (lambda x, y: (x - y) % P)
        mul = mul or (lambda x, y: (x * y) % P)
        lift = lift or (lambda x: int(x) % P)
        v = [None] * len(self.ops)
        out = []
        for k, (op, a, b, imm) in enumerate(self.ops):
            if op == CONST:
                v[k] = lift(imm)
            elif op == LV:
                v[k] = lv[a]
            elif op == NV:
                v[k] = nv[a]
            elif op == LA:
                v[k] = la[a]
            elif op == NA:
                v[k] = na[a]
            elif op == PI:
                v[k] = lift(pi[a])
            elif op == CH:
                v[k] = lift(ch[a])
            elif op == ADD:
                v[k] = add(v[a], v[b])
            elif op == SUB:
                v[k] = sub(v[a], v[b])
            elif op == MUL:
                v[k] = mul(v[a], v[b])
            else:
                out.append((op, v[a]))
        return out

    def check_trace(self, trace: np.ndarray, public_inputs=()) -> int:
        """check_constraints analogue on the trace domain (own constraints only, i.e. programs without aux reads):
        -1 if every row satisfies every constraint, else row * 1000 + constraint index.  Pure Python: small traces."""
        n = trace.shape[1]
        t = [[int(x) for x in col] for col in trace]
        for i in range(n):
            lv = [c[i] for c in t]
            nv = [c[(i + 1) % n] for c in t]
            for idx, (kind, val) in enumerate(self.evaluate(lv, nv, pi=public_inputs)):
                if kind == EMIT_TRANSITION and i == n - 1:
                    continue
                if kind == EMIT_FIRST_ROW and i != 0:
                    continue
                if kind == EMIT_LAST_ROW and i != n - 1:
                    continue
                if val % P:
                    return i * 1000 + idx
        return -1


# ---- tables of the shapes the reference proves -------------------------------------------------------------
def memory_program() -> Program:
    """The memory-shaped table (built in as ETP_TABLE_MEMORY) as a constraint program, in
    the same constraint order: proofs of the registered table must equal the built-in table's word for word
    (only the table id in the header differs)."""
    from .synthetic import (M_CFC, M_COUNTER, M_CTX, M_FILTER, M_FREQ, M_INIT_AUX, M_IS_READ, M_RANGE_CHECK, M_SEG, M_SFC,
                            M_TIMESTAMP, M_VALUE0, M_VFC, M_VIRT, MEMORY_COLUMNS)

    b = ProgramBuilder(MEMORY_COLUMNS, 0, 3)
    lv, nv, one = b.lv, b.nv, b.const(1)
    f = lv(M_FILTER)
    b.constraint(f * (f - one))
    b.constraint((one - f) * (one - lv(M_IS_READ)))
    cfc, sfc, vfc = lv(M_CFC), lv(M_SFC), lv(M_VFC)
    unchanged = one - cfc - sfc - vfc
    for x in (cfc, sfc, vfc, unchanged):
        b.constraint(x * (one - x))
    d_ctx, d_seg = nv(M_CTX) - lv(M_CTX), nv(M_SEG) - lv(M_SEG)
    d_virt, d_ts = nv(M_VIRT) - lv(M_VIRT), nv(M_TIMESTAMP) - lv(M_TIMESTAMP)
    for x in (sfc * d_ctx, vfc * d_ctx, vfc * d_seg, unchanged * d_ctx, unchanged * d_seg, unchanged * d_virt):
        b.transition(x)
    computed = (cfc * (d_ctx - one) + sfc * (d_seg - one)) + (vfc * (d_virt - one) + unchanged * d_ts)
    b.transition(lv(M_RANGE_CHECK) - computed)
    init_aux = lv(M_INIT_AUX)
    b.transition(init_aux - nv(M_SEG) * (one - unchanged) * nv(M_IS_READ))
    read_unchanged = nv(M_IS_READ) * unchanged
    ctx_init = nv(M_CTX) * init_aux
    seg_init = (nv(M_SEG) - 13) * init_aux
    for i in range(8):
        v, nvv = lv(M_VALUE0 + i), nv(M_VALUE0 + i)
        b.transition(read_unchanged * (nvv - v))
        b.transition(ctx_init * nvv)
        b.transition(seg_init * nvv)
    b.first_row(lv(M_COUNTER))
    b.transition(nv(M_COUNTER) - lv(M_COUNTER) - one)
    b.add_lookup([M_RANGE_CHECK], M_COUNTER, M_FREQ)
    b.emit_lookup_constraints()
    return b.build()


def fibonacci_program() -> Program:
    b = ProgramBuilder(2, 3, 2)
    b.first_row(b.lv(0) - b.pi(0))
    b.first_row(b.lv(1) - b.pi(1))
    b.last_row(b.lv(1) - b.pi(2))
    b.transition(b.nv(0) - b.lv(1))
    b.transition(b.nv(1) - b.lv(0) - b.lv(1))
    return b.build()


# A logic-shaped table (evm_arithmetization/src/logic.rs, recalled shape — SURVEY.md Appendix B): three
# operation flags, two operands and a result of `limbs` 32-bit limbs, the operands bit-decomposed (one column
# per bit).  For AND / OR / XOR the result limb is  sum_bit 2^bit * (s * (x + y) + a * x * y)  with
# (s, a) = (0, 1) AND, (1, -1) OR, (1, -2) XOR.  Columns: [is_and, is_or, is_xor, x bits..., y bits..., result limbs].
def logic_layout(limbs: int = 8):
    bits = 32 * limbs
    return {"IS_AND": 0, "IS_OR": 1, "IS_XOR": 2, "X": 3, "Y": 3 + bits, "RES": 3 + 2 * bits, "cols": 3 + 2 * bits + limbs,
            "bits": bits, "limbs": limbs}


def logic_program(limbs: int = 8) -> Program:
    L = logic_layout(limbs)
    b = ProgramBuilder(L["cols"], 0, 3)
    lv, one = b.lv, b.const(1)
    is_and, is_or, is_xor = lv(L["IS_AND"]), lv(L["IS_OR"]), lv(L["IS_XOR"])
    for fl in (is_and, is_or, is_xor):
        b.constraint(fl * (fl - one))
    flag_sum = is_and + is_or + is_xor
    b.constraint(flag_sum * (flag_sum - one))
    sum_coeff = is_or + is_xor
    and_coeff = is_and - is_or - is_xor * 2
    for i in range(L["bits"]):
        for base in (L["X"], L["Y"]):
            bit = lv(base + i)
            b.constraint(bit * (bit - one))
    for limb in range(limbs):
        x_lin = y_lin = xy = None
        for k in range(32):
            x, y = lv(L["X"] + 32 * limb + k), lv(L["Y"] + 32 * limb + k)
            w = b.const(1 << k)
            tx, ty, txy = x * w, y * w, x * y * w
            x_lin = tx if x_lin is None else x_lin + tx
            y_lin = ty if y_lin is None else y_lin + ty
            xy = txy if xy is None else xy + txy
        b.constraint(lv(L["RES"] + limb) - (sum_coeff * (x_lin + y_lin) + and_coeff * xy))
    return b.build()


def logic_trace(log_n: int, limbs: int = 8, seed: int = 11) -> np.ndarray:
    from .synthetic import _rand

    L = logic_layout(limbs)
    n = 1 << log_n
    t = np.zeros((L["cols"], n), dtype=np.uint64)
    op = (_rand(seed, 0, n) % np.uint64(4)).astype(np.int64)  # 3 = padding row (no flag set)
    t[L["IS_AND"]] = (op == 0)
    t[L["IS_OR"]] = (op == 1)
    t[L["IS_XOR"]] = (op == 2)
    for limb in range(limbs):
        x = _rand(seed, 1 + 2 * limb, n) & np.uint64(0xFFFFFFFF)
        y = _rand(seed, 2 + 2 * limb, n) & np.uint64(0xFFFFFFFF)
        for k in range(32):
            t[L["X"] + 32 * limb + k] = (x >> np.uint64(k)) & np.uint64(1)
            t[L["Y"] + 32 * limb + k] = (y >> np.uint64(k)) & np.uint64(1)
        res = np.where(op == 0, x & y, np.where(op == 1, x | y, np.where(op == 2, x ^ y, np.uint64(0))))
        t[L["RES"] + limb] = res
    return t


# A range-checked table in the style of the arithmetic table's 16-bit limb checks (evm_arithmetization/src/
# arithmetic/arithmetic_stark.rs, recalled shape): `n_limbs` limb columns that must hold values < 2^log_n, all of
# them looked up in one counter column (logUp with multiplicities), plus a toy relation c = a * b over the first limbs.
def rangecheck_layout(n_limbs: int = 7):
    return {"LIMB": 0, "COUNTER": n_limbs, "FREQ": n_limbs + 1, "PROD": n_limbs + 2, "cols": n_limbs + 3, "n_limbs": n_limbs}


def rangecheck_program(n_limbs: int = 7) -> Program:
    L = rangecheck_layout(n_limbs)
    b = ProgramBuilder(L["cols"], 0, 3)
    b.constraint(b.lv(L["PROD"]) - b.lv(0) * b.lv(1))
    b.first_row(b.lv(L["COUNTER"]))
    b.transition(b.nv(L["COUNTER"]) - b.lv(L["COUNTER"]) - 1)
    b.add_lookup(list(range(n_limbs)), L["COUNTER"], L["FREQ"])
    b.emit_lookup_constraints()
    return b.build()


def rangecheck_trace(log_n: int, n_limbs: int = 7, seed: int = 5) -> np.ndarray:
    from .synthetic import _rand

    L = rangecheck_layout(n_limbs)
    n = 1 << log_n
    t = np.zeros((L["cols"], n), dtype=np.uint64)
    freq = np.zeros(n, dtype=np.int64)
    for j in range(n_limbs):
        v = (_rand(seed, j, n) % np.uint64(n)).astype(np.int64)
        t[j] = v.astype(np.uint64)
        freq += np.bincount(v, minlength=n)
    t[L["COUNTER"]] = np.arange(n, dtype=np.uint64)
    t[L["FREQ"]] = freq.astype(np.uint64)
    t[L["PROD"]] = (t[0].astype(object) * t[1].astype(object) % P).astype(np.uint64)
    return t


# ---- shape-only stand-ins for the evm_arithmetization tables ---------------------------------------------------
# The sources of the seven tables the reference proves (/root/reference/ops/src/lib.rs:52 -> evm_arithmetization::
# {arithmetic, byte_packing, cpu, keccak, keccak_sponge, logic, memory}) are not available offline (hard
# part 1).  What a prover's cost depends on is their SHAPE: trace width, number and degree of the constraints, logUp
# lookups.  `shape_program` builds a table of a given width whose constraints have that profile and that a generated
# trace satisfies:  column 0 is a row counter (first-row + transition constraint); with `n_lookup` > 0, column 1 holds
# multiplicities and columns 2 .. 2 + n_lookup are limbs range-checked against the counter (logUp, chunked helpers);
# the remaining columns come in groups (a, b, d, c) with c = a*b + d, every fifth group c = a*b*d (degree 3), and up to
# three trailing boolean flags.  The widths below are the approximate upstream ones.
EVM_TABLE_SHAPES = {          # name: (trace columns, range-checked limbs)
    "arithmetic": (112, 16),
    "byte_packing": (80, 8),
    "cpu": (128, 0),
    "keccak": (2400, 0),
    "keccak_sponge": (424, 8),
}


def shape_layout(n_cols: int, n_lookup: int = 0):
    first = 1 + (1 + n_lookup if n_lookup else 0)
    assert n_cols >= first + 4
    n_groups = (n_cols - first) // 4
    return {"COUNTER": 0, "FREQ": 1 if n_lookup else None, "LIMB": 2 if n_lookup else None, "n_lookup": n_lookup,
            "GROUP": first, "n_groups": n_groups, "FLAG": first + 4 * n_groups, "n_flags": n_cols - first - 4 * n_groups,
            "cols": n_cols}


def shape_program(n_cols: int, n_lookup: int = 0, emit_lookups: bool = True) -> Program:
    """emit_lookups=False leaves the logUp constraints out (for Program.check_trace, which has no auxiliary columns)."""
    L = shape_layout(n_cols, n_lookup)
    b = ProgramBuilder(n_cols, 0, 3)
    lv, one = b.lv, b.const(1)
    b.first_row(lv(0))
    b.transition(b.nv(0) - lv(0) - 1)
    for g in range(L["n_groups"]):
        a, bb, d, c = (lv(L["GROUP"] + 4 * g + k) for k in range(4))
        b.constraint(c - a * bb * d if g % 5 == 4 else c - (a * bb + d))
    for f in range(L["n_flags"]):
        fl = lv(L["FLAG"] + f)
        b.constraint(fl * (fl - one))
    if n_lookup and emit_lookups:
        b.add_lookup(list(range(L["LIMB"], L["LIMB"] + n_lookup)), L["COUNTER"], L["FREQ"])
        b.emit_lookup_constraints()
    return b.build()


def shape_trace(log_n: int, n_cols: int, n_lookup: int = 0, seed: int = 23) -> np.ndarray:
    """A trace that satisfies shape_program(n_cols, n_lookup): operands below 2^20, so every product fits 64 bits."""
    from .synthetic import _rand

    L = shape_layout(n_cols, n_lookup)
    n = 1 << log_n
    t = np.zeros((n_cols, n), dtype=np.uint64)
    t[0] = np.arange(n, dtype=np.uint64)
    if n_lookup:
        freq = np.zeros(n, dtype=np.int64)
        for j in range(n_lookup):
            v = (_rand(seed, 1000 + j, n) % np.uint64(n)).astype(np.int64)
            t[L["LIMB"] + j] = v.astype(np.uint64)
            freq += np.bincount(v, minlength=n)
        t[L["FREQ"]] = freq.astype(np.uint64)
    m = np.uint64((1 << 20) - 1)
    for g in range(L["n_groups"]):
        c0 = L["GROUP"] + 4 * g
        a, bb, d = (_rand(seed, 3 * g + k, n) & m for k in range(3))
        t[c0], t[c0 + 1], t[c0 + 2] = a, bb, d
        t[c0 + 3] = a * bb * d if g % 5 == 4 else a * bb + d
    for f in range(L["n_flags"]):
        t[L["FLAG"] + f] = _rand(seed, 5000 + f, n) & np.uint64(1)
    return t


# degree bits of the seven tables in a small transaction: the low end of the reference's circuit ranges
# (/root/reference/README.md:53-59: arithmetic 15.., byte packing 9.., cpu 12.., keccak 14.., keccak sponge 9..,
# logic 12.., memory 17..), one notch up for the tables an ETH transfer actually exercises
TX_TABLE_DEGREE_BITS = {"arithmetic": 16, "byte_packing": 10, "cpu": 14, "keccak": 14, "keccak_sponge": 10, "logic": 12, "memory": 18}
