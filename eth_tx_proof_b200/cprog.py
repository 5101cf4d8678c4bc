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


class Column:
    """starky::lookup::Column: a linear combination of the local row, of the next row, plus a constant."""

    def __init__(self, local=(), next_row=(), constant=0):
        self.local = [(int(c), int(f) % P) for c, f in local]
        self.next_row = [(int(c), int(f) % P) for c, f in next_row]
        self.constant = int(constant) % P

    @classmethod
    def single(cls, c):
        return cls([(c, 1)])

    @classmethod
    def single_next_row(cls, c):
        return cls((), [(c, 1)])

    @classmethod
    def constant_(cls, v):
        return cls((), (), v)

    @classmethod
    def le_bits(cls, cols):
        return cls([(c, 1 << i) for i, c in enumerate(cols)])

    @classmethod
    def lift(cls, c):
        return c if isinstance(c, Column) else cls.single(c)

    def is_single(self):
        return len(self.local) == 1 and self.local[0][1] == 1 and not self.next_row and self.constant == 0

    def words(self):
        w = [len(self.local)]
        for c, f in self.local:
            w += [c, f]
        w.append(len(self.next_row))
        for c, f in self.next_row:
            w += [c, f]
        w.append(self.constant)
        return w

    def expr(self, b):
        """Column::eval_with_next as a program expression."""
        acc = None
        for c, f in self.local:
            t = b.lv(c) if f == 1 else b.lv(c) * f
            acc = t if acc is None else acc + t
        for c, f in self.next_row:
            t = b.nv(c) if f == 1 else b.nv(c) * f
            acc = t if acc is None else acc + t
        if acc is None:
            return b.const(self.constant)
        return acc + self.constant if self.constant else acc

    def eval_table(self, trace, i):
        """Column::eval_table: row i of a (cols, n) trace of Python ints / numpy u64, next row cyclic."""
        n = len(trace[0])
        acc = self.constant
        for c, f in self.local:
            acc += int(trace[c][i]) * f
        for c, f in self.next_row:
            acc += int(trace[c][(i + 1) % n]) * f
        return acc % P


class Filter:
    """starky::lookup::Filter: sum of products of two Columns plus a sum of Columns; Default = the constant 1."""

    def __init__(self, products=(), constants=None):
        self.products = [(Column.lift(a), Column.lift(b)) for a, b in products]
        self.constants = [Column.constant_(1)] if constants is None and not self.products else [Column.lift(c) for c in (constants or ())]

    def is_default(self):
        return not self.products and len(self.constants) == 1 and not self.constants[0].local and \
            not self.constants[0].next_row and self.constants[0].constant == 1

    def words(self):
        w = [len(self.products)]
        for a, b in self.products:
            w += a.words() + b.words()
        w.append(len(self.constants))
        for c in self.constants:
            w += c.words()
        return w

    def expr(self, b):
        acc = None
        for x, y in self.products:
            t = x.expr(b) * y.expr(b)
            acc = t if acc is None else acc + t
        for c in self.constants:
            t = c.expr(b)
            acc = t if acc is None else acc + t
        return acc if acc is not None else b.const(0)

    def eval_table(self, trace, i):
        acc = 0
        for x, y in self.products:
            acc += x.eval_table(trace, i) * y.eval_table(trace, i)
        for c in self.constants:
            acc += c.eval_table(trace, i)
        return acc % P


AUXSPEC_MAGIC = 0x3153585541505445  # "ETPAUXS1"
CH_CTL_BASE = NUM_CHALLENGES  # challenge scalars: lookup challenges [0, 2), then CTL (beta_k, gamma_k) at 2 + 2k, 3 + 2k


class ProgramBuilder:
    def __init__(self, n_trace_cols: int, n_public_inputs: int = 0, constraint_degree: int = 3):
        self.n_trace, self.n_pi, self.degree = n_trace_cols, n_public_inputs, constraint_degree
        self.n_aux = 0
        self.n_ch = 0
        self.ops: List[Tuple[int, int, int, int]] = []  # (opcode, a, b, imm)
        self._memo = {}
        self.n_constraints = 0
        self.lookups: List[Tuple[List[int], int, int]] = []  # (columns, table column, frequencies column[, filters])
        self.ctl_zs: List[Tuple[int, list]] = []  # (challenge index, [(columns, filter)]) in the table's CtlData order

    def _op(self, op, a=0, b=0, imm=0) -> Expr:
        if op == MUL:  # peephole: x * 1 (default filters) costs nothing
            for x, y in ((a, b), (b, a)):
                if self.ops[x][0] == CONST and self.ops[x][3] == 1:
                    return Expr(self, y)
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

    # ---- starky::lookup::eval_packed_lookups_generic / eval_helper_columns --------------------------------------
    def add_lookup(self, looking: Sequence, table_col, freq_col, filters: Sequence = None):
        """Lookup { columns, table_column, frequencies_column, filter_columns }: columns / table / frequencies are
        trace column indices or `Column`s, filters default to Filter::default() (the constant 1)."""
        looking = list(looking)
        if filters is None:
            self.lookups.append((looking, table_col, freq_col))
        else:
            assert len(filters) == len(looking)
            self.lookups.append((looking, table_col, freq_col, list(filters)))

    def _helper_constraints(self, combos, filts, helpers):
        """eval_helper_columns: chunks of (degree - 1) = 1 or 2 (combin, filter) pairs against one helper column each."""
        chunk = max(1, self.degree - 1)
        assert chunk <= 2, "eval_helper_columns: todo!(\"Allow other constraint degrees\") upstream"
        for c, h in enumerate(helpers):
            cs, fs = combos[c * chunk:(c + 1) * chunk], filts[c * chunk:(c + 1) * chunk]
            if len(cs) == 2:
                self.constraint(cs[1] * cs[0] * h - fs[0] * cs[1] - fs[1] * cs[0])
            else:
                self.constraint(cs[0] * h - fs[0])

    def emit_lookup_constraints(self):
        """Call once, after the table's own constraints.  Auxiliary columns: per lookup, per challenge: one helper
        column per chunk of (degree - 1) looking columns, then Z."""
        chunk = max(1, self.degree - 1)
        start = 0
        for lk in self.lookups:
            looking, table_col, freq_col = lk[0], lk[1], lk[2]
            filters = lk[3] if len(lk) > 3 else [Filter()] * len(looking)
            n_help = -(-len(looking) // chunk)
            for k in range(NUM_CHALLENGES):
                ch = self.challenge(k)
                combos = [Column.lift(j).expr(self) + ch for j in looking]
                filts = [f.expr(self) for f in filters]
                helpers = [self.la(start + c) for c in range(n_help)]
                self._helper_constraints(combos, filts, helpers)
                z, next_z = self.la(start + n_help), self.na(start + n_help)
                twc = Column.lift(table_col).expr(self) + ch
                hs = helpers[0]
                for h in helpers[1:]:
                    hs = hs + h
                y = hs * twc - Column.lift(freq_col).expr(self)
                self.first_row(z)
                self.constraint((next_z - z) * twc - y)
                start += n_help + 1
        self.n_aux = max(self.n_aux, start)
        self._n_lookup_cols = start

    # ---- starky::cross_table_lookup: CtlZData of this table + eval_cross_table_lookup_checks ---------------------
    def add_ctl_z(self, challenge: int, colsets: Sequence):
        """One CtlZData of this table, in the order cross_table_lookup_data pushes them: `challenge` = index of the
        (beta, gamma) pair, `colsets` = [(columns, filter)] — one pair when the table appears once in the CTL (looked
        table, or a single looking entry), several when it is looking more than once."""
        assert 0 <= challenge < NUM_CHALLENGES
        self.ctl_zs.append((challenge, [([Column.lift(c) for c in cols], f if f is not None else Filter()) for cols, f in colsets]))

    def num_ctl_helper_columns(self) -> int:
        chunk = max(1, self.degree - 1)
        return sum(-(-len(sets) // chunk) if len(sets) > 1 else 0 for _, sets in self.ctl_zs)

    def emit_ctl_constraints(self):
        """Call once, after emit_lookup_constraints.  Auxiliary columns after the lookup columns: all CTL helper columns
        (Z by Z), then all CTL Z columns."""
        n_lookup = self.num_lookup_columns()
        n_helpers = self.num_ctl_helper_columns()
        chunk = max(1, self.degree - 1)
        h_start = n_lookup
        for zi, (k, sets) in enumerate(self.ctl_zs):
            beta, gamma = self.challenge(CH_CTL_BASE + 2 * k), self.challenge(CH_CTL_BASE + 2 * k + 1)
            combos, filts = [], []
            for cols, filt in sets:
                acc = None  # challenges.combine = reduce_with_powers(evals, beta) + gamma
                for c in reversed(cols):
                    e = c.expr(self)
                    acc = e if acc is None else acc * beta + e
                combos.append(acc + gamma)
                filts.append(filt.expr(self))
            z_col = n_lookup + n_helpers + zi
            local_z, next_z = self.la(z_col), self.na(z_col)
            if len(sets) > 1:
                n_h = -(-len(sets) // chunk)
                helpers = [self.la(h_start + c) for c in range(n_h)]
                h_start += n_h
                self._helper_constraints(combos, filts, helpers)
                h_sum = helpers[0]
                for h in helpers[1:]:
                    h_sum = h_sum + h
                self.last_row(local_z - h_sum)
                self.transition(local_z - next_z - h_sum)
            else:
                self.last_row(combos[0] * local_z - filts[0])
                self.transition(combos[0] * (local_z - next_z) - filts[0])
        self.n_aux = max(self.n_aux, n_lookup + n_helpers + len(self.ctl_zs))

    def num_lookup_columns(self) -> int:
        chunk = max(1, self.degree - 1)
        return sum(-(-len(l[0]) // chunk) + 1 for l in self.lookups) * NUM_CHALLENGES

    def num_aux_columns(self) -> int:
        return self.num_lookup_columns() + self.num_ctl_helper_columns() + len(self.ctl_zs)

    def build(self) -> "Program":
        return Program(self)


class Program:
    def __init__(self, b: ProgramBuilder):
        self.n_trace, self.n_aux, self.n_pi, self.n_ch = b.n_trace, b.n_aux, b.n_pi, b.n_ch
        self.degree, self.n_constraints = b.degree, b.n_constraints
        self.lookups = [tuple(l) for l in b.lookups]
        self.ctl_zs = list(b.ctl_zs)
        self.n_lookup_cols = b.num_lookup_columns()
        self.n_ctl_helper_cols = b.num_ctl_helper_columns()
        self.ops = list(b.ops)
        w = np.zeros(8 + 2 * len(self.ops), dtype=np.uint64)
        w[:8] = [MAGIC, len(self.ops), self.n_trace, self.n_aux, self.n_pi, self.n_ch, self.degree, self.n_constraints]
        for k, (op, a, bb, imm) in enumerate(self.ops):
            w[8 + 2 * k] = op | (a << 8) | (bb << 36)
            w[9 + 2 * k] = imm
        self.words = w

    def simple_lookups(self) -> bool:
        """True when every lookup uses plain columns and default filters and there is no CTL: the round-1 flat i32 lookup
        description of etp_table_register is enough."""
        if self.ctl_zs:
            return False
        for lk in self.lookups:
            if len(lk) > 3 and not all(f.is_default() for f in lk[3]):
                return False
            if not all(isinstance(c, int) or c.is_single() for c in list(lk[0]) + [lk[1], lk[2]]):
                return False
        return True

    def flat_lookups(self):
        """[(looking columns, table column, frequencies column)] with plain column indices (simple_lookups() only)."""
        ix = lambda c: c if isinstance(c, int) else c.local[0][0]
        return [([ix(c) for c in lk[0]], ix(lk[1]), ix(lk[2])) for lk in self.lookups]

    @property
    def aux_spec(self) -> np.ndarray:
        """Auxiliary-column description for etp_table_register_ex (u64 words): every Lookup with its Columns and Filters,
        then the table's CtlZData descriptors (include/etp_b200.h)."""
        w = [AUXSPEC_MAGIC, len(self.lookups), len(self.ctl_zs)]
        for lk in self.lookups:
            looking = [Column.lift(c) for c in lk[0]]
            filters = lk[3] if len(lk) > 3 else [Filter()] * len(looking)
            w.append(len(looking))
            for c in looking:
                w += c.words()
            for f in filters:
                w += f.words()
            w += Column.lift(lk[1]).words() + Column.lift(lk[2]).words()
        for k, sets in self.ctl_zs:
            w += [k, len(sets)]
            for cols, filt in sets:
                w.append(len(cols))
                for c in cols:
                    w += c.words()
                w += filt.words()
        return np.array(w, dtype=np.uint64)

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
    b = memory_builder()
    b.emit_lookup_constraints()
    return b.build()


def memory_builder(extra_cols: int = 0) -> ProgramBuilder:
    """The memory-shaped table's own constraints + its range-check Lookup on a builder with `extra_cols` more trace columns
    (CTL ports); the caller emits the lookup / CTL checks."""
    from .synthetic import (M_CFC, M_COUNTER, M_CTX, M_FILTER, M_FREQ, M_INIT_AUX, M_IS_READ, M_RANGE_CHECK, M_SEG, M_SFC,
                            M_TIMESTAMP, M_VALUE0, M_VFC, M_VIRT, MEMORY_COLUMNS)

    b = ProgramBuilder(MEMORY_COLUMNS + extra_cols, 0, 3)
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
    return b


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
    return logic_builder(limbs).build()


def logic_builder(limbs: int = 8, extra_cols: int = 0) -> ProgramBuilder:
    L = logic_layout(limbs)
    b = ProgramBuilder(L["cols"] + extra_cols, 0, 3)
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
    return b


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
    b = shape_builder(n_cols, n_lookup, with_lookup=emit_lookups)
    if n_lookup and emit_lookups:
        b.emit_lookup_constraints()
    return b.build()


def shape_builder(n_cols: int, n_lookup: int = 0, extra_cols: int = 0, with_lookup: bool = True) -> ProgramBuilder:
    L = shape_layout(n_cols, n_lookup)
    b = ProgramBuilder(n_cols + extra_cols, 0, 3)
    lv, one = b.lv, b.const(1)
    b.first_row(lv(0))
    b.transition(b.nv(0) - lv(0) - 1)
    for g in range(L["n_groups"]):
        a, bb, d, c = (lv(L["GROUP"] + 4 * g + k) for k in range(4))
        b.constraint(c - a * bb * d if g % 5 == 4 else c - (a * bb + d))
    for f in range(L["n_flags"]):
        fl = lv(L["FLAG"] + f)
        b.constraint(fl * (fl - one))
    if n_lookup and with_lookup:
        b.add_lookup(list(range(L["LIMB"], L["LIMB"] + n_lookup)), L["COUNTER"], L["FREQ"])
    return b


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


# ---- a small multi-table system linked by a cross-table lookup --------------------------------------------------------
# The CTL mechanism of starky 0.4.0 / evm_arithmetization (cross_table_lookup.rs: CtlData, partial_sums,
# eval_cross_table_lookup_checks, verify_cross_table_lookups; reached from /root/reference/ops/src/lib.rs:52 through
# prove_with_traces) on synthetic tables — the seven EVM tables themselves are not available offline:
#   table 0 "ops"   looks (key, value) tuples up TWICE per row (two filtered looking entries of the same table -> one CTL
#                   helper column + Z) and range-checks its keys against its own counter with a filtered logUp Lookup whose
#                   looked column is a linear combination;
#   table 1 "rom"   is the looked table: distinct (key, value) rows, filter = multiplicity column;
#   table 2 "extra" looks up once per row, through linear-combination Columns that also read the NEXT row (Z only).
def ctl_demo_tables(log_ops: int = 6, log_rom: int = 5, log_extra: int = 5, seed: int = 3):
    """-> (tables, ctls): tables = [(name, Program, trace)], ctls = [(looking table indices, looked table index)]."""
    from .synthetic import _rand

    n_ops, n_rom, n_ext = 1 << log_ops, 1 << log_rom, 1 << log_extra
    value = lambda k: (k.astype(object) * 1000003 + 17) % P  # the ROM contents
    mult = np.zeros(n_rom, dtype=np.int64)

    # ---- table 0 "ops": columns F1 K1 V1 F2 K2 V2 COUNTER FREQ HALF
    F1, K1, V1, F2, K2, V2, CNT, FREQ, HALF = range(9)
    t0 = np.zeros((9, n_ops), dtype=np.uint64)
    f1 = (_rand(seed, 1, n_ops) % np.uint64(4) != 0).astype(np.int64)
    f2 = (_rand(seed, 2, n_ops) % np.uint64(3) == 0).astype(np.int64)
    k1 = (_rand(seed, 3, n_ops) % np.uint64(n_rom)).astype(np.int64)
    k2 = (_rand(seed, 4, n_ops) % np.uint64(n_rom)).astype(np.int64)
    t0[F1], t0[K1], t0[F2], t0[K2] = f1, k1, f2, k2
    t0[V1] = np.array(value(k1), dtype=np.uint64)
    t0[V2] = np.array(value(k2), dtype=np.uint64)
    t0[CNT] = np.arange(n_ops, dtype=np.uint64)
    t0[HALF] = t0[K1] // np.uint64(2)  # K1 = 2*HALF + (K1 & 1): the lookup below range-checks 2*HALF (a linear combination)
    # filtered logUp: rows with F1 = 1 look 2*HALF up in COUNTER; K2 is looked up on every row
    freq = np.bincount((2 * (k1 // 2))[f1 == 1], minlength=n_ops) + np.bincount(k2, minlength=n_ops)
    t0[FREQ] = freq.astype(np.uint64)
    np.add.at(mult, k1[f1 == 1], 1)
    np.add.at(mult, k2[f2 == 1], 1)
    b0 = ProgramBuilder(9, 0, 3)
    for f in (F1, F2):
        b0.constraint(b0.lv(f) * (b0.lv(f) - 1))
    b0.first_row(b0.lv(CNT))
    b0.transition(b0.nv(CNT) - b0.lv(CNT) - 1)
    b0.add_lookup([Column([(HALF, 2)]), K2], CNT, FREQ, [Filter(constants=[Column.single(F1)]), Filter()])
    for k in range(NUM_CHALLENGES):
        b0.add_ctl_z(k, [([K1, V1], Filter(constants=[Column.single(F1)])), ([K2, V2], Filter(products=[(F2, F2)], constants=[]))])
    b0.emit_lookup_constraints()
    b0.emit_ctl_constraints()

    # ---- table 2 "extra": columns A B S with key = A + B (mod n_rom handled by construction), value column = V; the looked
    # tuple is (A + B, V) and the filter reads the NEXT row: active iff next row's S is 1
    A, B, V, S = range(4)
    t2 = np.zeros((4, n_ext), dtype=np.uint64)
    a = (_rand(seed, 5, n_ext) % np.uint64(n_rom // 2)).astype(np.int64)
    bb = (_rand(seed, 6, n_ext) % np.uint64(n_rom // 2)).astype(np.int64)
    s = (_rand(seed, 7, n_ext) % np.uint64(2)).astype(np.int64)
    t2[A], t2[B], t2[S] = a, bb, s
    t2[V] = np.array(value(a + bb), dtype=np.uint64)
    active = np.roll(s, -1)  # row i is active iff S[i + 1] = 1 (cyclic, as Column::eval_table does)
    np.add.at(mult, (a + bb)[active == 1], 1)
    b2 = ProgramBuilder(4, 0, 3)
    b2.constraint(b2.lv(S) * (b2.lv(S) - 1))
    for k in range(NUM_CHALLENGES):
        b2.add_ctl_z(k, [([Column([(A, 1), (B, 1)]), V], Filter(constants=[Column.single_next_row(S)]))])
    b2.emit_lookup_constraints()
    b2.emit_ctl_constraints()

    # ---- table 1 "rom": columns KEY VAL MULT
    KEY, VAL, MULT = range(3)
    t1 = np.zeros((3, n_rom), dtype=np.uint64)
    keys = np.arange(n_rom, dtype=np.int64)
    t1[KEY] = keys
    t1[VAL] = np.array(value(keys), dtype=np.uint64)
    t1[MULT] = mult.astype(np.uint64)
    b1 = ProgramBuilder(3, 0, 3)
    b1.first_row(b1.lv(KEY))
    b1.transition(b1.nv(KEY) - b1.lv(KEY) - 1)
    for k in range(NUM_CHALLENGES):
        b1.add_ctl_z(k, [([KEY, VAL], Filter(constants=[Column.single(MULT)]))])
    b1.emit_lookup_constraints()
    b1.emit_ctl_constraints()
    tables = [("ops", b0.build(), t0), ("rom", b1.build(), t1), ("extra", b2.build(), t2)]
    ctls = [([0, 0, 2], 1)]
    return tables, ctls


# ---- a synthetic TRANSACTION: seven tables of the evm_arithmetization shapes linked by cross-table lookups ------------------
# Table order and CTL topology follow evm_arithmetization 0.1.3 (src/all_stark.rs: Table::all(), all_cross_table_lookups;
# /root/reference/Cargo.lock:1675, reached from /root/reference/ops/src/lib.rs:52): cpu looks into arithmetic, byte packing,
# keccak sponge, logic and (three memory channels here) memory; keccak sponge looks into keccak (inputs and outputs), logic
# and memory; byte packing looks into memory.  The tables themselves are the shape stand-ins above (the real constraint
# sets are not available offline) with CTL "ports" appended: a looking port is (filter, key, value) columns with
# value = A_c * key + B_c, a looked port opens the tuple (row counter, A_c * counter + B_c) — the value as a
# linear-combination Column — filtered by a multiplicity column.  Proven through prover.prove_with_traces, i.e. with
# upstream's transcript threading and CTL challenges.
EVM_TABLE_ORDER = ("arithmetic", "byte_packing", "cpu", "keccak", "keccak_sponge", "logic", "memory")
EVM_CTLS = [  # (looking tables, looked table), indices into EVM_TABLE_ORDER
    ([2], 0), ([2], 1), ([2], 4), ([4], 3), ([4], 3), ([2, 4], 5), ([2, 2, 2, 4, 1], 6),
]


def evm_shaped_system(scale_bits: int = 0, seed: int = 1, degree_bits: dict = None):
    """-> (tables, ctls): tables = [(name, Program, trace)] in EVM_TABLE_ORDER, ctls = EVM_CTLS."""
    from .synthetic import _rand, memory_trace, MEMORY_COLUMNS
    from .synthetic import M_COUNTER as MEM_COUNTER

    bits = {k: max(5, v + scale_bits) for k, v in (degree_bits or TX_TABLE_DEGREE_BITS).items()}
    n_rows = {k: 1 << v for k, v in bits.items()}
    names = EVM_TABLE_ORDER
    # ports per table, in CtlData order: for each CTL (list order): looking entries grouped per table, then the looked one
    looking = {t: [] for t in range(7)}   # table -> [(ctl index, [port ordinals within that CTL for this table])]
    looked = {t: [] for t in range(7)}    # table -> [ctl index]
    for c, (lk, ld) in enumerate(EVM_CTLS):
        for t in dict.fromkeys(lk):
            looking[t].append((c, lk.count(t)))
        looked[ld].append(c)
    n_extra = {t: sum(3 * k for _, k in looking[t]) + len(looked[t]) for t in range(7)}
    A = lambda c: 1000003 + 7919 * c
    B = lambda c: 17 + c
    logic_limbs = 8
    base_cols, builders, traces, counter_col = {}, {}, {}, {}
    for t, name in enumerate(names):
        if name == "memory":
            base_cols[t] = MEMORY_COLUMNS
            builders[t] = memory_builder(n_extra[t])
            base = memory_trace(bits[name], seed=31 + seed)
            counter_col[t] = MEM_COUNTER
        elif name == "logic":
            L = logic_layout(logic_limbs)
            base_cols[t] = L["cols"] + 1  # + a row counter for the looked port's key
            builders[t] = logic_builder(logic_limbs, 1 + n_extra[t])
            base = np.concatenate([logic_trace(bits[name], logic_limbs, seed=37 + seed), np.arange(n_rows[name], dtype=np.uint64)[None, :]])
            counter_col[t] = L["cols"]
            bb = builders[t]
            bb.first_row(bb.lv(counter_col[t]))
            bb.transition(bb.nv(counter_col[t]) - bb.lv(counter_col[t]) - 1)
        else:
            cols, lk = EVM_TABLE_SHAPES[name]
            base_cols[t] = cols
            builders[t] = shape_builder(cols, lk, n_extra[t])
            base = shape_trace(bits[name], cols, lk, seed=41 + seed + t)
            counter_col[t] = 0
        traces[t] = np.concatenate([base, np.zeros((n_extra[t], n_rows[name]), dtype=np.uint64)])
    # fill the looking ports, count multiplicities
    mult = {c: np.zeros(n_rows[names[ld]], dtype=np.int64) for c, (_, ld) in enumerate(EVM_CTLS)}
    port_cols = {t: {} for t in range(7)}  # table -> {(ctl, j): (F, K, V)}; looked: {("m", ctl): MULT}
    for t, name in enumerate(names):
        col = base_cols[t]
        n = n_rows[name]
        for c, k in looking[t]:
            n_looked = n_rows[names[EVM_CTLS[c][1]]]
            for j in range(k):
                f = (_rand(seed, 100 * c + 10 * j + 1, n) % np.uint64(3) != 0).astype(np.int64)
                key = (_rand(seed, 100 * c + 10 * j + 2, n) % np.uint64(n_looked)).astype(np.int64)
                traces[t][col], traces[t][col + 1] = f, key
                traces[t][col + 2] = np.array((key.astype(object) * A(c) + B(c)) % P, dtype=np.uint64)
                np.add.at(mult[c], key[f == 1], 1)
                port_cols[t][(c, j)] = (col, col + 1, col + 2)
                col += 3
        for c in looked[t]:
            port_cols[t][("m", c)] = col
            col += 1
    for t in range(7):
        for c in looked[t]:
            traces[t][port_cols[t][("m", c)]] = mult[c].astype(np.uint64)
    # CtlZData per table in cross_table_lookup_data's order: per CTL, per challenge: looking tables (grouped), then looked
    for c, (lk, ld) in enumerate(EVM_CTLS):
        for k in range(NUM_CHALLENGES):
            for t in dict.fromkeys(lk):
                sets = []
                for j in range(lk.count(t)):
                    F, K, V = port_cols[t][(c, j)]
                    builders[t].constraint(builders[t].lv(F) * (builders[t].lv(F) - 1)) if k == 0 else None
                    sets.append(([K, V], Filter(constants=[Column.single(F)])))
                builders[t].add_ctl_z(k, sets)
            cc = counter_col[ld]
            builders[ld].add_ctl_z(k, [([cc, Column([(cc, A(c))], constant=B(c))], Filter(constants=[Column.single(port_cols[ld][("m", c)])]))])
    tables = []
    for t, name in enumerate(names):
        b = builders[t]
        b.emit_lookup_constraints()
        b.emit_ctl_constraints()
        tables.append((name, b.build(), traces[t]))
    return tables, [(list(lk), ld) for lk, ld in EVM_CTLS]
