"""EVM-style tables with REAL semantics, written as constraint programs for the generic table path (cprog.ProgramBuilder ->
etp_table_register_ex / the oracle's interpreter).  The sources of evm_arithmetization 0.1.3's tables are not available offline
(/root/reference/Cargo.lock:1675; SURVEY.md section 7, hard part 1), so these are NOT upstream's constraint lists: they are
this repository's own layouts, built from the published designs the upstream tables follow, and their traces are checked
against independent computations (Python integers, hashlib's SHA-3).  They show that the path proves real tables, not only
shape stand-ins; `cprog.EVM_TABLE_SHAPES` keeps the stand-ins the benchmark's synthetic transaction uses.

arithmetic   ADD, SUB, LT, GT, MUL on n_limbs x limb_bits-bit words (16 x 16 = 256 bits by default).
             addcy (evm_arithmetization/src/arithmetic/addcy.rs, recalled design): with s_i = x_i + y_i - z_i and
             t_i = s_i + cy_{i-1}, every t_i must be 0 or 2^w, and cy_i = t_i / 2^w is an EXPRESSION (no carry columns);
             the last carry equals the CY column.  ADD: A + B = C + CY 2^W.  SUB / LT / GT: the same relation on permuted
             operands (A - B = C with borrow CY; LT's result is CY).  MUL: schoolbook columns sum_{i+j=k} a_i b_j with
             carries split into two range-checked limbs.  Every limb is range-checked by a logUp lookup into a counter column
             (filters = the row's operation flags), and the table exposes (opcode, A, B, C / CY) as a CTL port.
keccak       Keccak-f[1600], one round per row, bit columns (below).
keccak256    Keccak-256 (Ethereum's hash) of one-block messages: the looking side of the keccak table's input / output CTLs.
byte_packing a 256-bit value <-> up to 32 bytes in memory: looked by a CPU-side port, looking into memory with 32 column sets.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from .cprog import NUM_CHALLENGES, P, Column, Filter, Program, ProgramBuilder

OP_ADD, OP_SUB, OP_LT, OP_GT, OP_MUL = range(5)


def arithmetic_layout(n_limbs: int = 16, limb_bits: int = 16) -> dict:
    L = {"FLAG": 0, "A": 5, "B": 5 + n_limbs, "C": 5 + 2 * n_limbs, "CY": 5 + 3 * n_limbs, "CARRY_LO": 6 + 3 * n_limbs,
         "CARRY_HI": 6 + 4 * n_limbs, "COUNTER": 6 + 5 * n_limbs, "FREQ": 7 + 5 * n_limbs, "cols": 8 + 5 * n_limbs,
         "n_limbs": n_limbs, "limb_bits": limb_bits}
    return L


def arithmetic_builder(n_limbs: int = 16, limb_bits: int = 16, with_ctl: bool = False, emit_lookups: bool = True) -> ProgramBuilder:
    """Constraint degree 3.  with_ctl: the table is the LOOKED side of a CTL on the tuple (opcode = 1 + operation, A limbs,
    B limbs, C limbs, CY) filtered by "the row holds an operation" (upstream opens the operands and the result to the CPU table
    the same way).  emit_lookups=False: the table's own constraints only (Program.check_trace on the trace domain)."""
    L = arithmetic_layout(n_limbs, limb_bits)
    b = ProgramBuilder(L["cols"], 0, 3)
    lv, nv = b.lv, b.nv
    W = 1 << limb_bits
    inv_w = pow(W, P - 2, P)
    flags = [lv(L["FLAG"] + k) for k in range(5)]
    for f in flags:
        b.constraint(f * (f - 1))
    fsum = flags[0] + flags[1] + flags[2] + flags[3] + flags[4]
    b.constraint(fsum * (fsum - 1))  # at most one operation per row; none = padding row
    A = [lv(L["A"] + i) for i in range(n_limbs)]
    B = [lv(L["B"] + i) for i in range(n_limbs)]
    C = [lv(L["C"] + i) for i in range(n_limbs)]
    cy = lv(L["CY"])
    b.constraint(cy * (cy - 1))

    def addcy(filt, x, y, z):
        """filt * [x + y == z + cy 2^(n w)], limb by limb without carry columns."""
        carry = None
        for i in range(n_limbs):
            t = x[i] + y[i] - z[i]
            if carry is not None:
                t = t + carry
            b.constraint(filt * (t * (t - W)))
            carry = t * inv_w
        b.constraint(filt * (carry - cy))

    addcy(flags[OP_ADD], A, B, C)                       # A + B = C + cy 2^W
    addcy(flags[OP_SUB] + flags[OP_LT], C, B, A)        # C + B = A + cy 2^W  <=>  A - B = C - cy 2^W: cy = [A < B]
    addcy(flags[OP_GT], C, A, B)                        # B - A = C - cy 2^W: cy = [B < A] = [A > B]
    # MUL: column k of the schoolbook product plus the incoming carry = c_k + 2^w * outgoing carry, carries = lo + 2^w hi
    lo = [lv(L["CARRY_LO"] + i) for i in range(n_limbs)]
    hi = [lv(L["CARRY_HI"] + i) for i in range(n_limbs)]
    for k in range(n_limbs):
        acc = None
        for i in range(k + 1):
            term = A[i] * B[k - i]
            acc = term if acc is None else acc + term
        if k:
            acc = acc + lo[k - 1] + hi[k - 1] * W
        b.constraint(flags[OP_MUL] * (acc - C[k] - (lo[k] + hi[k] * W) * W))
    # range counter: 0, then +0 / +1 steps, ending at 2^w - 1: every w-bit value occurs
    cnt, cnt_n = lv(L["COUNTER"]), nv(L["COUNTER"])
    b.first_row(cnt)
    d = cnt_n - cnt
    b.transition(d * (d - 1))
    b.last_row(cnt - (W - 1))
    looking = [L["A"] + i for i in range(n_limbs)] + [L["B"] + i for i in range(n_limbs)] + [L["C"] + i for i in range(n_limbs)] + \
              [L["CARRY_LO"] + i for i in range(n_limbs)] + [L["CARRY_HI"] + i for i in range(n_limbs)]
    b.add_lookup(looking, L["COUNTER"], L["FREQ"])
    if with_ctl:
        opcode = Column([(L["FLAG"] + k, k + 1) for k in range(5)])
        cols = [opcode] + [Column.single(L["A"] + i) for i in range(n_limbs)] + [Column.single(L["B"] + i) for i in range(n_limbs)] + \
               [Column.single(L["C"] + i) for i in range(n_limbs)] + [Column.single(L["CY"])]
        filt = Filter(constants=[Column([(L["FLAG"] + k, 1) for k in range(5)])])
        for k in range(NUM_CHALLENGES):
            b.add_ctl_z(k, [(cols, filt)])
    if emit_lookups:
        b.emit_lookup_constraints()
        if with_ctl:
            b.emit_ctl_constraints()
    return b


def arithmetic_program(n_limbs: int = 16, limb_bits: int = 16, with_ctl: bool = False, emit_lookups: bool = True) -> Program:
    return arithmetic_builder(n_limbs, limb_bits, with_ctl, emit_lookups).build()


def _limbs(v: int, n_limbs: int, limb_bits: int) -> List[int]:
    return [(v >> (limb_bits * i)) & ((1 << limb_bits) - 1) for i in range(n_limbs)]


def arithmetic_row(op: int, a: int, b_: int, n_limbs: int = 16, limb_bits: int = 16) -> Tuple[List[int], List[int], List[int], int, List[int], List[int]]:
    """One operation computed with Python integers -> (A limbs, B limbs, C limbs, CY, carry lo limbs, carry hi limbs)."""
    bits = n_limbs * limb_bits
    M = 1 << bits
    W = 1 << limb_bits
    lo, hi = [0] * n_limbs, [0] * n_limbs
    if op == OP_ADD:
        c, cy = (a + b_) % M, (a + b_) >> bits
    elif op in (OP_SUB, OP_LT):
        c, cy = (a - b_) % M, int(a < b_)
    elif op == OP_GT:
        c, cy = (b_ - a) % M, int(b_ < a)
    else:
        c, cy = (a * b_) % M, 0
        al, bl = _limbs(a, n_limbs, limb_bits), _limbs(b_, n_limbs, limb_bits)
        carry = 0
        for k in range(n_limbs):
            s = sum(al[i] * bl[k - i] for i in range(k + 1)) + carry
            carry = s >> limb_bits
            assert carry < W * W
            lo[k], hi[k] = carry % W, carry // W
    return _limbs(a, n_limbs, limb_bits), _limbs(b_, n_limbs, limb_bits), _limbs(c, n_limbs, limb_bits), cy, lo, hi


def arithmetic_trace(log_n: int, n_limbs: int = 16, limb_bits: int = 16, seed: int = 31, n_ops: int = None):
    """-> (trace, ops): `ops` = [(opcode, a, b, result)] of the active rows (result = C, or CY for LT / GT) — what a looking
    table would send through the CTL.  Needs 2^log_n >= 2^limb_bits rows (the range counter)."""
    from .synthetic import _rand

    L = arithmetic_layout(n_limbs, limb_bits)
    n, W = 1 << log_n, 1 << limb_bits
    if n < W:
        raise ValueError("the range counter needs 2^limb_bits rows")
    n_ops = n - n // 4 if n_ops is None else n_ops
    t = np.zeros((L["cols"], n), dtype=np.uint64)
    words = [[int(x) for x in _rand(seed, k, n)] for k in range(2 * ((n_limbs * limb_bits + 63) // 64) + 1)]
    per = (n_limbs * limb_bits + 63) // 64
    M = 1 << (n_limbs * limb_bits)
    freq = np.zeros(n, dtype=np.int64)
    ops = []
    for r in range(n):
        if r < n_ops:
            op = words[2 * per][r] % 5
            a = sum(words[k][r] << (64 * k) for k in range(per)) % M
            b_ = sum(words[per + k][r] << (64 * k) for k in range(per)) % M
            if r % 7 == 3:
                b_ = a  # equal operands: LT / GT give 0, SUB gives 0
            if r % 11 == 5:
                a, b_ = M - 1, M - 1 - (r % 3)  # carries ripple through every limb
            al, bl, cl, cy, lo, hi = arithmetic_row(op, a, b_, n_limbs, limb_bits)
            t[L["FLAG"] + op, r] = 1
            ops.append((op, a, b_, cy if op in (OP_LT, OP_GT) else sum(c << (limb_bits * i) for i, c in enumerate(cl))))
        else:
            al = bl = cl = lo = hi = [0] * n_limbs
            cy = 0
        for i in range(n_limbs):
            t[L["A"] + i, r], t[L["B"] + i, r], t[L["C"] + i, r] = al[i], bl[i], cl[i]
            t[L["CARRY_LO"] + i, r], t[L["CARRY_HI"] + i, r] = lo[i], hi[i]
        t[L["CY"], r] = cy
        for v in list(al) + list(bl) + list(cl) + list(lo) + list(hi):
            freq[v] += 1
    # counter: 0 .. W - 1, then stays at W - 1; the multiplicity of a value sits on its FIRST row
    cnt = np.minimum(np.arange(n), W - 1)
    t[L["COUNTER"]] = cnt.astype(np.uint64)
    fr = np.zeros(n, dtype=np.int64)
    fr[:W] = freq[:W]
    t[L["FREQ"]] = fr.astype(np.uint64)
    return t, ops


# ---- Keccak-f[1600]: one round per row, the state as 1600 bit columns ------------------------------------------------------------
# FIPS 202 section 3.2 (theta, rho, pi, chi, iota), lane (x, y) bit z at A[64 (5 y + x) + z].  Own layout, degree 3 — the same
# decomposition the upstream keccak table uses (evm_arithmetization/src/keccak/keccak_stark.rs, recalled: bit columns, XORs as
# low-degree polynomials, round flags, 24 rows per permutation), with more committed intermediates to keep every step explicit:
#   FLAG[24]  one-hot round flags, rotating
#   ID        permutation index (constant over a permutation's 24 rows): ties the input port to the output port
#   A[1600]   state before the round (boolean)
#   T1[320]   xor3(A[x][0], A[x][1], A[x][2])            C[320]  xor3(T1[x], A[x][3], A[x][4])   (theta's column parities)
#   AP[1600]  A xor D,  D[x][z] = C[x-1][z] xor C[x+1][z-1]                                     (after theta)
#   AQ[64]    chi's output bit for lane (0, 0), before iota
#   OUT[1600] the round's output: chi(pi(rho(AP))) and, for lane (0, 0), AQ xor RC[round]
#   REAL      1 on the rows of a permutation some other table asked for, 0 on padding permutations (constant over a permutation)
# next.A == OUT except across a permutation boundary (round 23 -> round 0 of the next permutation).
# CTL ports (the looked side of keccak_sponge -> keccak in upstream): (ID, 50 input limbs of 32 bits) on the round-0 rows and
# (ID, 50 output limbs) on the round-23 rows of REAL permutations (filter = REAL * round flag).
KECCAK_ROUNDS = 24


def _keccak_constants():
    rc = []
    r = 1
    for _ in range(KECCAK_ROUNDS):  # rc(t) LFSR, FIPS 202 algorithm 5
        v = 0
        for j in range(7):
            if r & 1:
                v |= 1 << ((1 << j) - 1)
            r <<= 1
            if r & 0x100:
                r ^= 0x171
        rc.append(v)
    rot = [[0] * 5 for _ in range(5)]  # rot[x][y], FIPS 202 algorithm 2
    x, y = 1, 0
    for t in range(24):
        rot[x][y] = ((t + 1) * (t + 2) // 2) % 64
        x, y = y, (2 * x + 3 * y) % 5
    return rc, rot


KECCAK_RC, KECCAK_ROT = _keccak_constants()


def keccak_layout() -> dict:
    L = {"FLAG": 0, "ID": 24, "A": 25, "T1": 1625, "C": 1945, "AP": 2265, "AQ": 3865, "OUT": 3929, "REAL": 5529, "cols": 5530}
    return L


def _bit(base: int, x: int, y: int, z: int) -> int:
    return base + 64 * (5 * y + x) + z


def keccak_builder(with_ctl: bool = False, emit_lookups: bool = True) -> ProgramBuilder:
    L = keccak_layout()
    b = ProgramBuilder(L["cols"], 0, 3)
    lv, nv = b.lv, b.nv
    xor2 = lambda p, q: p + q - p * q * 2
    xor3 = lambda p, q, r: p + q + r - (p * q + q * r + r * p) * 2 + p * q * r * 4
    f = [lv(L["FLAG"] + r) for r in range(KECCAK_ROUNDS)]
    fsum = None
    for r in range(KECCAK_ROUNDS):
        b.constraint(f[r] * (f[r] - 1))
        fsum = f[r] if fsum is None else fsum + f[r]
        b.transition(nv(L["FLAG"] + (r + 1) % KECCAK_ROUNDS) - f[r])
    b.constraint(fsum - 1)
    b.first_row(f[0] - 1)
    not_last = b.const(1) - f[KECCAK_ROUNDS - 1]
    b.transition(not_last * (nv(L["ID"]) - lv(L["ID"])))
    real = lv(L["REAL"])
    b.constraint(real * (real - 1))
    b.transition(not_last * (nv(L["REAL"]) - real))
    A = lambda x, y, z: lv(_bit(L["A"], x, y, z))
    AP = lambda x, y, z: lv(_bit(L["AP"], x, y, z))
    OUT = lambda x, y, z: lv(_bit(L["OUT"], x, y, z))
    for x in range(5):
        for y in range(5):
            for z in range(64):
                a = A(x, y, z)
                b.constraint(a * (a - 1))
    # theta
    C = [[None] * 64 for _ in range(5)]
    for x in range(5):
        for z in range(64):
            t1, c = lv(L["T1"] + 64 * x + z), lv(L["C"] + 64 * x + z)
            b.constraint(t1 - xor3(A(x, 0, z), A(x, 1, z), A(x, 2, z)))
            b.constraint(c - xor3(t1, A(x, 3, z), A(x, 4, z)))
            C[x][z] = c
    for x in range(5):
        for z in range(64):
            d = xor2(C[(x + 4) % 5][z], C[(x + 1) % 5][(z + 63) % 64])
            for y in range(5):
                b.constraint(AP(x, y, z) - xor2(A(x, y, z), d))
    # rho + pi: B[y][2x + 3y] = rot(AP[x][y], r[x][y]); bit z of a left rotation by r is bit z - r
    B = [[None] * 5 for _ in range(5)]
    for x in range(5):
        for y in range(5):
            r = KECCAK_ROT[x][y]
            B[y][(2 * x + 3 * y) % 5] = [AP(x, y, (z - r) % 64) for z in range(64)]
    # chi (+ iota on lane (0, 0))
    for x in range(5):
        for y in range(5):
            for z in range(64):
                e = xor2(B[x][y][z], (b.const(1) - B[(x + 1) % 5][y][z]) * B[(x + 2) % 5][y][z])
                if (x, y) == (0, 0):
                    aq = lv(L["AQ"] + z)
                    b.constraint(aq - e)
                    rc = None
                    for r in range(KECCAK_ROUNDS):
                        if (KECCAK_RC[r] >> z) & 1:
                            rc = f[r] if rc is None else rc + f[r]
                    b.constraint(OUT(0, 0, z) - (aq if rc is None else xor2(aq, rc)))
                else:
                    b.constraint(OUT(x, y, z) - e)
    # the next row continues the permutation
    for i in range(1600):
        b.transition(not_last * (nv(L["A"] + i) - lv(L["OUT"] + i)))
    if with_ctl:
        def limbs(base):
            return [Column.le_bits([base + 32 * k + j for j in range(32)]) for k in range(50)]

        for k in range(NUM_CHALLENGES):  # CTL "inputs": looked on round-0 rows
            b.add_ctl_z(k, [([Column.single(L["ID"])] + limbs(L["A"]), Filter(products=[(L["REAL"], L["FLAG"])]))])
        for k in range(NUM_CHALLENGES):  # CTL "outputs": looked on round-23 rows
            b.add_ctl_z(k, [([Column.single(L["ID"])] + limbs(L["OUT"]), Filter(products=[(L["REAL"], L["FLAG"] + KECCAK_ROUNDS - 1)]))])
        if emit_lookups:
            b.emit_lookup_constraints()
            b.emit_ctl_constraints()
    return b


def keccak_program(with_ctl: bool = False, emit_lookups: bool = True) -> Program:
    return keccak_builder(with_ctl, emit_lookups).build()


def keccak_f(lanes: List[int]) -> List[int]:
    """Keccak-f[1600] on 25 lanes (lane x + 5 y), plain Python — the definition the table is tested against."""
    a = list(lanes)
    M = (1 << 64) - 1
    rotl = lambda v, r: ((v << r) | (v >> (64 - r))) & M if r else v
    for rnd in range(KECCAK_ROUNDS):
        c = [a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20] for x in range(5)]
        d = [c[(x + 4) % 5] ^ rotl(c[(x + 1) % 5], 1) for x in range(5)]
        a = [a[i] ^ d[i % 5] for i in range(25)]
        bq = [0] * 25
        for x in range(5):
            for y in range(5):
                bq[y + 5 * ((2 * x + 3 * y) % 5)] = rotl(a[x + 5 * y], KECCAK_ROT[x][y])
        a = [bq[x + 5 * y] ^ ((~bq[(x + 1) % 5 + 5 * y] & M) & bq[(x + 2) % 5 + 5 * y]) for y in range(5) for x in range(5)]
        a[0] ^= KECCAK_RC[rnd]
    return a


def keccak_trace(log_n: int, inputs: List[List[int]] = None, seed: int = 17):
    """-> (trace, io): one permutation per 24 rows (the last rows of the table hold the first rounds of one more permutation:
    the round flags keep rotating); io = [(id, input lanes, output lanes)] of the COMPLETE permutations.  With `inputs` given,
    permutation p < len(inputs) is REAL (id p + 1) and the rest of the table is padding (zero state, REAL = 0)."""
    from .synthetic import _rand

    L = keccak_layout()
    n = 1 << log_n
    n_perm = -(-n // KECCAK_ROUNDS)
    n_real = n // KECCAK_ROUNDS if inputs is None else len(inputs)
    if n_real > n // KECCAK_ROUNDS:
        raise ValueError("every real permutation needs its 24 rows")
    if inputs is None:
        words = [_rand(seed, k, n_perm) for k in range(25)]
        inputs = [[int(words[k][p]) for k in range(25)] for p in range(n_perm)]
        inputs[0] = [0] * 25
    else:
        inputs = [list(x) for x in inputs] + [[0] * 25] * (n_perm - len(inputs))
    t = np.zeros((L["cols"], n), dtype=np.uint64)
    bits = lambda lanes: np.array([(lanes[i] >> z) & 1 for i in range(25) for z in range(64)], dtype=np.uint64)
    io = []
    M = (1 << 64) - 1
    rotl = lambda v, r: ((v << r) | (v >> (64 - r))) & M if r else v
    for p in range(n_perm):
        a = list(inputs[p])
        for rnd in range(KECCAK_ROUNDS):
            row = p * KECCAK_ROUNDS + rnd
            if row >= n:
                break
            t[L["FLAG"] + rnd, row] = 1
            t[L["ID"], row] = p + 1
            t[L["REAL"], row] = int(p < n_real)
            t[L["A"]:L["A"] + 1600, row] = bits(a)
            t1 = [a[x] ^ a[x + 5] ^ a[x + 10] for x in range(5)]
            c = [t1[x] ^ a[x + 15] ^ a[x + 20] for x in range(5)]
            t[L["T1"]:L["T1"] + 320, row] = np.array([(t1[x] >> z) & 1 for x in range(5) for z in range(64)], dtype=np.uint64)
            t[L["C"]:L["C"] + 320, row] = np.array([(c[x] >> z) & 1 for x in range(5) for z in range(64)], dtype=np.uint64)
            d = [c[(x + 4) % 5] ^ rotl(c[(x + 1) % 5], 1) for x in range(5)]
            ap = [a[i] ^ d[i % 5] for i in range(25)]
            t[L["AP"]:L["AP"] + 1600, row] = bits(ap)
            bq = [0] * 25
            for x in range(5):
                for y in range(5):
                    bq[y + 5 * ((2 * x + 3 * y) % 5)] = rotl(ap[x + 5 * y], KECCAK_ROT[x][y])
            out = [bq[x + 5 * y] ^ ((~bq[(x + 1) % 5 + 5 * y] & M) & bq[(x + 2) % 5 + 5 * y]) for y in range(5) for x in range(5)]
            t[L["AQ"]:L["AQ"] + 64, row] = np.array([(out[0] >> z) & 1 for z in range(64)], dtype=np.uint64)
            out[0] ^= KECCAK_RC[rnd]
            t[L["OUT"]:L["OUT"] + 1600, row] = bits(out)
            a = out
        else:
            if p < n_real:
                io.append((p + 1, list(inputs[p]), a))
    return t, io


# ---- Keccak-256 of short messages: the looking side of the two keccak CTLs ---------------------------------------------------------
# One row per message of at most 135 bytes (one rate block; upstream's keccak_sponge table absorbs any number of blocks and XORs them
# into the state through the logic table — with one block and a zero initial state the XOR is the block itself).  Columns:
#   F, ID, LEN, POS[136] (one-hot position of the first padding byte = LEN), BYTE[136] (the padded block: message bytes, 0x01 at LEN,
#   zeros, 0x80 added to the last byte — Keccak's pad10*1 with the legacy 0x01 domain byte Ethereum uses), OUT[50] (the permutation's
#   output as 32-bit limbs: OUT[0..8] is the digest), COUNTER / FREQ (every BYTE is range-checked to 8 bits by a logUp lookup).
# CTLs: (ID, 34 block limbs, 16 zero capacity limbs) into the keccak table's inputs, (ID, OUT[50]) into its outputs, filter F.
RATE_BYTES = 136


def keccak256_layout() -> dict:
    return {"F": 0, "ID": 1, "LEN": 2, "POS": 3, "BYTE": 3 + RATE_BYTES, "OUT": 3 + 2 * RATE_BYTES, "COUNTER": 53 + 2 * RATE_BYTES,
            "FREQ": 54 + 2 * RATE_BYTES, "cols": 55 + 2 * RATE_BYTES}


def keccak256_builder(with_ctl: bool = True, emit_lookups: bool = True) -> ProgramBuilder:
    L = keccak256_layout()
    b = ProgramBuilder(L["cols"], 0, 3)
    lv, nv = b.lv, b.nv
    f = lv(L["F"])
    b.constraint(f * (f - 1))
    pos = [lv(L["POS"] + i) for i in range(RATE_BYTES)]
    byte = [lv(L["BYTE"] + i) for i in range(RATE_BYTES)]
    total = length = None
    for i in range(RATE_BYTES):
        b.constraint(pos[i] * (pos[i] - 1))
        total = pos[i] if total is None else total + pos[i]
        if i:
            length = pos[i] * i if length is None else length + pos[i] * i
    b.constraint(total - f)                      # exactly one padding position on a message row, none on a padding row
    b.constraint(lv(L["LEN"]) - length)
    after = None                                 # after[i] = [LEN < i]
    for i in range(RATE_BYTES - 1):
        if after is not None:
            b.constraint(after * byte[i])        # past the first padding byte: zero
        b.constraint(pos[i] * (byte[i] - 1))     # the first padding byte: 0x01
        after = pos[i] if after is None else after + pos[i]
    b.constraint(byte[RATE_BYTES - 1] - (f * 0x80 + pos[RATE_BYTES - 1]))  # the last byte: 0x80, or 0x81 when LEN = 135
    cnt = lv(L["COUNTER"])
    b.first_row(cnt)
    d = nv(L["COUNTER"]) - cnt
    b.transition(d * (d - 1))
    b.last_row(cnt - 255)
    b.add_lookup([L["BYTE"] + i for i in range(RATE_BYTES)], L["COUNTER"], L["FREQ"])
    if with_ctl:
        block = [Column([(L["BYTE"] + 4 * k + j, 1 << (8 * j)) for j in range(4)]) for k in range(RATE_BYTES // 4)]
        zero = [Column.constant_(0)] * (50 - len(block))
        filt = Filter(constants=[Column.single(L["F"])])
        for k in range(NUM_CHALLENGES):
            b.add_ctl_z(k, [([Column.single(L["ID"])] + block + zero, filt)])
        for k in range(NUM_CHALLENGES):
            b.add_ctl_z(k, [([Column.single(L["ID"])] + [Column.single(L["OUT"] + i) for i in range(50)], filt)])
    if emit_lookups:
        b.emit_lookup_constraints()
        if with_ctl:
            b.emit_ctl_constraints()
    return b


def keccak256_program(with_ctl: bool = True, emit_lookups: bool = True) -> Program:
    return keccak256_builder(with_ctl, emit_lookups).build()


def keccak256(msg: bytes) -> bytes:
    """Keccak-256 (the legacy 0x01 padding: Ethereum's hash) of a message of any length, by the plain-Python permutation."""
    m = bytearray(msg) + b"\x01"
    m += b"\x00" * (-len(m) % RATE_BYTES)
    m[-1] |= 0x80
    st = [0] * 25
    for off in range(0, len(m), RATE_BYTES):
        for i in range(RATE_BYTES // 8):
            st[i] ^= int.from_bytes(m[off + 8 * i:off + 8 * i + 8], "little")
        st = keccak_f(st)
    return b"".join(x.to_bytes(8, "little") for x in st[:4])


def keccak256_system(messages: List[bytes], log_n_sponge: int = 8, log_n_keccak: int = None):
    """-> (tables, ctls, digests): the message table (looking) and the Keccak-f table (looked) linked by the input and output
    CTLs; digests read from the message table's OUT limbs."""
    L = keccak256_layout()
    n = 1 << log_n_sponge
    if n < 256 or len(messages) > n or any(len(m) >= RATE_BYTES for m in messages):
        raise ValueError("the message table needs >= 256 rows (byte range counter), one row per message of at most 135 bytes")
    t = np.zeros((L["cols"], n), dtype=np.uint64)
    lanes_in = []
    freq = np.zeros(256, dtype=np.int64)
    for r, msg in enumerate(messages):
        blk = bytearray(msg) + b"\x01" + b"\x00" * (RATE_BYTES - len(msg) - 1)
        blk[-1] |= 0x80
        t[L["F"], r], t[L["ID"], r], t[L["LEN"], r] = 1, r + 1, len(msg)
        t[L["POS"] + len(msg), r] = 1
        t[L["BYTE"]:L["BYTE"] + RATE_BYTES, r] = np.frombuffer(bytes(blk), dtype=np.uint8)
        lanes_in.append([int.from_bytes(blk[8 * i:8 * i + 8], "little") for i in range(17)] + [0] * 8)
    for r in range(n):
        for v in t[L["BYTE"]:L["BYTE"] + RATE_BYTES, r]:
            freq[int(v)] += 1
    t[L["COUNTER"]] = np.minimum(np.arange(n), 255).astype(np.uint64)
    t[L["FREQ"], :256] = freq.astype(np.uint64)
    if log_n_keccak is None:
        log_n_keccak = max(5, (KECCAK_ROUNDS * len(messages) - 1).bit_length())
    kt, io = keccak_trace(log_n_keccak, inputs=lanes_in)
    digests = []
    for r, (_, _, out) in enumerate(io):
        for k in range(50):
            t[L["OUT"] + k, r] = (out[k // 2] >> (32 * (k % 2))) & 0xFFFFFFFF
        digests.append(b"".join(x.to_bytes(8, "little") for x in out[:4]))
    tables = [("keccak256", keccak256_program(), t), ("keccak", keccak_program(with_ctl=True), kt)]
    ctls = [([0], 1), ([0], 1)]
    return tables, ctls, digests


# ---- byte packing: between a 256-bit value and a byte sequence in memory ----------------------------------------------------------
# evm_arithmetization/src/byte_packing/byte_packing_stark.rs, recalled design: one row per MLOAD_32BYTES / MSTORE_32BYTES-style
# operation of LEN <= 32 bytes.  Columns: IS_READ, LEN_FLAG[32] (one-hot: LEN = i + 1), CTX, SEG, VIRT (address of the first byte),
# TS, BYTE[32] (BYTE[0] = the LEAST significant byte of the value, i.e. the LAST byte in memory, as upstream stores them),
# COUNTER / FREQ (bytes range-checked to 8 bits).  Constraints: the flags are boolean and at most one is set (none = padding row),
# bytes at positions >= LEN are zero.  CTL ports:
#   to the CPU table (looked by it here; upstream: cpu looks into byte packing): (IS_READ, CTX, SEG, VIRT, LEN, TS, 8 value limbs of
#       32 bits = linear combinations of the bytes), filter = sum of the length flags;
#   into MEMORY (looking, 32 column sets of ONE Z per challenge — the chunked-helper path): for byte i the tuple
#       (IS_READ, CTX, SEG, VIRT + LEN - 1 - i, BYTE[i], TS), filter = [i < LEN] = sum_{j >= i} LEN_FLAG[j].
NUM_PACK_BYTES = 32


def byte_packing_layout() -> dict:
    return {"IS_READ": 0, "LEN_FLAG": 1, "CTX": 33, "SEG": 34, "VIRT": 35, "TS": 36, "BYTE": 37, "COUNTER": 69, "FREQ": 70, "cols": 71}


def byte_packing_builder(with_ctl: bool = True, emit_lookups: bool = True) -> ProgramBuilder:
    L = byte_packing_layout()
    b = ProgramBuilder(L["cols"], 0, 3)
    lv, nv = b.lv, b.nv
    is_read = lv(L["IS_READ"])
    b.constraint(is_read * (is_read - 1))
    flags = [lv(L["LEN_FLAG"] + i) for i in range(NUM_PACK_BYTES)]
    total = None
    for fl in flags:
        b.constraint(fl * (fl - 1))
        total = fl if total is None else total + fl
    b.constraint(total * (total - 1))
    b.constraint((b.const(1) - total) * is_read)          # a padding row carries no operation
    covered = None                                        # [i < LEN] = sum_{j >= i} flag_j, built from the top
    inside = [None] * NUM_PACK_BYTES
    for i in reversed(range(NUM_PACK_BYTES)):
        covered = flags[i] if covered is None else covered + flags[i]
        inside[i] = covered
    for i in range(NUM_PACK_BYTES):
        b.constraint((b.const(1) - inside[i]) * lv(L["BYTE"] + i))  # bytes past the length are zero
    cnt = lv(L["COUNTER"])
    b.first_row(cnt)
    d = nv(L["COUNTER"]) - cnt
    b.transition(d * (d - 1))
    b.last_row(cnt - 255)
    b.add_lookup([L["BYTE"] + i for i in range(NUM_PACK_BYTES)], L["COUNTER"], L["FREQ"])
    if with_ctl:
        length = Column([(L["LEN_FLAG"] + i, i + 1) for i in range(NUM_PACK_BYTES)])
        active = Column([(L["LEN_FLAG"] + i, 1) for i in range(NUM_PACK_BYTES)])
        limbs = [Column([(L["BYTE"] + 4 * k + j, 1 << (8 * j)) for j in range(4)]) for k in range(8)]
        cpu_cols = [Column.single(L[c]) for c in ("IS_READ", "CTX", "SEG", "VIRT")] + [length, Column.single(L["TS"])] + limbs
        mem_sets = []
        for i in range(NUM_PACK_BYTES):
            # address of byte i = VIRT + LEN - 1 - i
            addr = Column([(L["VIRT"], 1)] + [(L["LEN_FLAG"] + j, j + 1) for j in range(NUM_PACK_BYTES)], constant=P - 1 - i)
            cols = [Column.single(L["IS_READ"]), Column.single(L["CTX"]), Column.single(L["SEG"]), addr, Column.single(L["BYTE"] + i),
                    Column.single(L["TS"])]
            mem_sets.append((cols, Filter(constants=[Column([(L["LEN_FLAG"] + j, 1) for j in range(i, NUM_PACK_BYTES)])])))
        for k in range(NUM_CHALLENGES):  # CTL 0: the CPU side (this table is looked)
            b.add_ctl_z(k, [(cpu_cols, Filter(constants=[active]))])
        for k in range(NUM_CHALLENGES):  # CTL 1: 32 looking entries into memory
            b.add_ctl_z(k, mem_sets)
    if emit_lookups:
        b.emit_lookup_constraints()
        if with_ctl:
            b.emit_ctl_constraints()
    return b


def byte_packing_program(with_ctl: bool = True, emit_lookups: bool = True) -> Program:
    return byte_packing_builder(with_ctl, emit_lookups).build()


def byte_packing_system(log_n: int = 8, n_ops: int = 24, seed: int = 41):
    """-> (tables, ctls, ops): [cpu_side (looking CTL 0), byte_packing, memory_side (looked by CTL 1)] and the operations
    [(is_read, ctx, seg, virt, len, ts, value)].  The two side tables are minimal port tables: a row of cpu_side lists one
    operation with its 256-bit value as 8 limbs; a row of memory_side lists one (is_read, ctx, seg, addr, byte, ts) access."""
    from .synthetic import _rand

    L = byte_packing_layout()
    n = 1 << log_n
    if n < 256 or n_ops > n:
        raise ValueError("the byte range counter needs >= 256 rows")
    t = np.zeros((L["cols"], n), dtype=np.uint64)
    r64 = [[int(x) for x in _rand(seed, k, n)] for k in range(8)]
    ops, accesses = [], []
    freq = np.zeros(256, dtype=np.int64)
    for r in range(n_ops):
        ln = 1 + r64[0][r] % NUM_PACK_BYTES if r % 5 else (NUM_PACK_BYTES if r % 10 else 1)
        is_read, ctx, seg, virt, ts = r64[1][r] & 1, r64[2][r] % 7, r64[3][r] % 5, r64[4][r] % 1000, 1 + r
        value = (r64[5][r] | (r64[6][r] << 64) | (r64[7][r] << 128) | (r64[0][r] << 192)) % (1 << (8 * ln))
        t[L["IS_READ"], r], t[L["CTX"], r], t[L["SEG"], r], t[L["VIRT"], r], t[L["TS"], r] = is_read, ctx, seg, virt, ts
        t[L["LEN_FLAG"] + ln - 1, r] = 1
        for i in range(ln):
            byte = (value >> (8 * i)) & 0xFF
            t[L["BYTE"] + i, r] = byte
            accesses.append((is_read, ctx, seg, virt + ln - 1 - i, byte, ts))
        ops.append((is_read, ctx, seg, virt, ln, ts, value))
    for r in range(n):
        for v in t[L["BYTE"]:L["BYTE"] + NUM_PACK_BYTES, r]:
            freq[int(v)] += 1
    t[L["COUNTER"]] = np.minimum(np.arange(n), 255).astype(np.uint64)
    t[L["FREQ"], :256] = freq.astype(np.uint64)

    def port_table(rows, width):
        m = 1 << max(5, (len(rows) - 1).bit_length())
        pt = np.zeros((width + 1, m), dtype=np.uint64)
        for r, row in enumerate(rows):
            pt[0, r] = 1
            pt[1:, r] = np.array(row, dtype=np.uint64)
        pb = ProgramBuilder(width + 1, 0, 3)
        pb.constraint(pb.lv(0) * (pb.lv(0) - 1))
        for k in range(NUM_CHALLENGES):
            pb.add_ctl_z(k, [(list(range(1, width + 1)), Filter(constants=[Column.single(0)]))])
        pb.emit_lookup_constraints()
        pb.emit_ctl_constraints()
        return pb.build(), pt

    cpu_rows = [[o[0], o[1], o[2], o[3], o[4], o[5]] + [(o[6] >> (32 * k)) & 0xFFFFFFFF for k in range(8)] for o in ops]
    cpu_prog, cpu_t = port_table(cpu_rows, 14)
    mem_prog, mem_t = port_table([list(a) for a in accesses], 6)
    tables = [("cpu_side", cpu_prog, cpu_t), ("byte_packing", byte_packing_program(), t), ("memory_side", mem_prog, mem_t)]
    ctls = [([0], 1), ([1], 2)]
    return tables, ctls, ops


# ---- a transaction of seven tables with real semantics, wired like upstream's AllStark -----------------------------------------------
# Table order = evm_arithmetization's Table::all(): arithmetic, byte_packing, cpu, keccak, keccak_sponge (here: the Keccak-256 message
# table), logic, memory (here: a port table listing byte accesses; the recalled memory STARK of cprog.memory_program needs a
# read-consistent access sequence, which random byte-packing operations are not).  Cross-table lookups, in upstream's order:
#   cpu -> arithmetic | cpu -> byte_packing | cpu -> keccak_sponge | keccak_sponge -> keccak (inputs) | keccak_sponge -> keccak
#   (outputs) | cpu -> logic | byte_packing -> memory
# The cpu table is a DISPATCHER, not an interpreter: one row per operation with a one-hot family flag and a payload that is the
# tuple the family's table opens — it stands where upstream's CPU table stands in the CTL graph, without its instruction semantics.
REAL_TABLE_ORDER = ("arithmetic", "byte_packing", "cpu", "keccak", "keccak_sponge", "logic", "memory")
REAL_CTLS = [([2], 0), ([2], 1), ([2], 4), ([4], 3), ([4], 3), ([2], 5), ([1], 6)]


def _port_ctl(b: ProgramBuilder, cols, filt):
    for k in range(NUM_CHALLENGES):
        b.add_ctl_z(k, [(cols, filt)])


def real_transaction_system(n_limbs: int = 4, limb_bits: int = 5, logic_limbs: int = 2, messages: List[bytes] = None, seed: int = 3):
    """-> (tables, ctls): [(name, Program, trace)] in REAL_TABLE_ORDER and REAL_CTLS.  Small parameters keep the pure-Python
    trace generators and the CPU oracle fast; n_limbs = 16, limb_bits = 16 is the 256-bit arithmetic (needs 2^16 rows)."""
    from . import cprog as cp

    messages = [b"", b"abc", b"eth-tx-proof on B200"] if messages is None else messages
    # arithmetic (looked by cpu)
    AL = arithmetic_layout(n_limbs, limb_bits)
    log_a = max(6, limb_bits)
    at, _ = arithmetic_trace(log_a, n_limbs, limb_bits, seed=seed)
    arith = arithmetic_program(n_limbs, limb_bits, with_ctl=True)
    a_rows = [[sum((k + 1) * int(at[AL["FLAG"] + k, r]) for k in range(5))] + [int(at[AL[c] + i, r]) for c in "ABC" for i in range(n_limbs)] +
              [int(at[AL["CY"], r])] for r in range(at.shape[1]) if any(at[AL["FLAG"] + k, r] for k in range(5))]
    # byte packing (looked by cpu, looking into memory)
    ptables, _, pops = byte_packing_system(seed=seed + 1)
    pack_t, mem_prog, mem_t = ptables[1][2], ptables[2][1], ptables[2][2]
    p_rows = [[o[0], o[1], o[2], o[3], o[4], o[5]] + [(o[6] >> (32 * k)) & 0xFFFFFFFF for k in range(8)] for o in pops]
    # Keccak-256 messages (looked by cpu, looking into keccak twice) and Keccak-f
    ktables, _, digests = keccak256_system(messages)
    KL = keccak256_layout()
    kb = keccak256_builder(with_ctl=False, emit_lookups=False)  # own constraints + the byte lookup; ports added in CTL order below
    _port_ctl(kb, [Column.single(KL["ID"]), Column.single(KL["LEN"])] + [Column.single(KL["OUT"] + i) for i in range(8)],
              Filter(constants=[Column.single(KL["F"])]))
    block = [Column([(KL["BYTE"] + 4 * k + j, 1 << (8 * j)) for j in range(4)]) for k in range(RATE_BYTES // 4)]
    filt = Filter(constants=[Column.single(KL["F"])])
    _port_ctl(kb, [Column.single(KL["ID"])] + block + [Column.constant_(0)] * (50 - len(block)), filt)
    _port_ctl(kb, [Column.single(KL["ID"])] + [Column.single(KL["OUT"] + i) for i in range(50)], filt)
    kb.emit_lookup_constraints()
    kb.emit_ctl_constraints()
    k_rows = [[r + 1, len(m)] + [int.from_bytes(d[4 * i:4 * i + 4], "little") for i in range(8)] for r, (m, d) in enumerate(zip(messages, digests))]
    # logic (looked by cpu)
    LL = cp.logic_layout(logic_limbs)
    lt = cp.logic_trace(6, logic_limbs, seed=seed + 2)
    lb = cp.logic_builder(logic_limbs)
    limb = lambda base, k: Column.le_bits([base + 32 * k + j for j in range(32)])
    l_cols = [Column([(LL["IS_AND"], 1), (LL["IS_OR"], 2), (LL["IS_XOR"], 3)])] + [limb(LL["X"], k) for k in range(logic_limbs)] + \
             [limb(LL["Y"], k) for k in range(logic_limbs)] + [Column.single(LL["RES"] + k) for k in range(logic_limbs)]
    _port_ctl(lb, l_cols, Filter(constants=[Column([(LL["IS_AND"], 1), (LL["IS_OR"], 1), (LL["IS_XOR"], 1)])]))
    lb.emit_lookup_constraints()
    lb.emit_ctl_constraints()
    word = lambda base, k, r: sum(int(lt[base + 32 * k + j, r]) << j for j in range(32))
    l_rows = [[int(lt[LL["IS_AND"], r]) + 2 * int(lt[LL["IS_OR"], r]) + 3 * int(lt[LL["IS_XOR"], r])] + [word(LL["X"], k, r) for k in range(logic_limbs)] +
              [word(LL["Y"], k, r) for k in range(logic_limbs)] + [int(lt[LL["RES"] + k, r]) for k in range(logic_limbs)]
              for r in range(lt.shape[1]) if lt[LL["IS_AND"], r] or lt[LL["IS_OR"], r] or lt[LL["IS_XOR"], r]]
    # the cpu dispatcher: family flags + payload; CTL order: arithmetic, byte_packing, keccak_sponge, logic
    families = [a_rows, p_rows, k_rows, l_rows]
    width = max(len(rows[0]) for rows in families)
    n_rows = sum(len(rows) for rows in families)
    log_c = max(5, (n_rows - 1).bit_length())
    ct = np.zeros((4 + width, 1 << log_c), dtype=np.uint64)
    r = 0
    for fam, rows in enumerate(families):
        for row in rows:
            ct[fam, r] = 1
            ct[4:4 + len(row), r] = np.array(row, dtype=np.uint64)
            r += 1
    cb = ProgramBuilder(4 + width, 0, 3)
    fsum = None
    for fam in range(4):
        fl = cb.lv(fam)
        cb.constraint(fl * (fl - 1))
        fsum = fl if fsum is None else fsum + fl
    cb.constraint(fsum * (fsum - 1))
    for fam, rows in enumerate(families):
        _port_ctl(cb, list(range(4, 4 + len(rows[0]))), Filter(constants=[Column.single(fam)]))
    cb.emit_lookup_constraints()
    cb.emit_ctl_constraints()
    tables = [("arithmetic", arith, at), ("byte_packing", byte_packing_program(), pack_t), ("cpu", cb.build(), ct),
              ("keccak", keccak_program(with_ctl=True), ktables[1][2]), ("keccak_sponge", kb.build(), ktables[0][2]),
              ("logic", lb.build(), lt), ("memory", mem_prog, mem_t)]
    return tables, list(REAL_CTLS)
