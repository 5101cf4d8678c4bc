#!/usr/bin/env python
"""Dev-time generator for the FP64-pipe tables in eth_tx_proof_b200/csrc/poseidon_constants.h, and an
exact-integer model of the data flow that csrc/poseidon.cuh runs on the FP64 pipe.

The 360 Poseidon-12 round constants of plonky2's PoseidonGoldilocksConfig (plonky2/src/hash/
poseidon_goldilocks.rs ALL_ROUND_CONSTANTS; crate pinned at /root/reference/Cargo.lock:3441, not on disk)
are NOT derived here: they are read back from the committed header (table ETP_POSEIDON_RC_TABLE, whose
SHA-256 is pinned by tests/golden/poseidon_kat.json and checked below), so this tool never touches oracle/.

What is generated (all integers, exact in binary64):

  * the three-level split of the circulant MDS layer.  With x the 12 lane values of one 32-bit plane,
      S_k = x_k + x_{k+6}, D_k = x_k - x_{k+6}                  (k < 6)
      T_j = S_j + S_{j+3}, E_j = S_j - S_{j+3}                  (j < 3)
      P_j = sum_k CP_k T_{(k+j)%3}              (cyclic 3)       CP_k = (CA_k + CA_{k+3}) / 2
      Q_j = sum_k CQ_k (+-) E_{(k+j)%3}         (negacyclic 3)   CQ_k = (CA_k - CA_{k+3}) / 2
      B_r = sum_k CB_k (+-) D_{(k+r)%6}         (negacyclic 6)   CB_k = (C_k - C_{k+6}) / 2,  CA_k = (C_k + C_{k+6}) / 2
      A_j = P_j + Q_j, A_{j+3} = P_j - Q_j ;  out_r = A_r + B_r, out_{r+6} = A_r - B_r
    for the coefficient vector C of the layer (MDS_MATRIX_CIRC) and for C*C (two layers at once);
  * per full-round layer, 12 accumulator start values per plane (c_0..2 | q_0..2 | b_0..5) that inject the
    2^52 mantissa bias and the NEXT round's constants;
  * per PAIR of partial rounds (r, r+1): 13 values per plane (u-chain start | c | q | b), see poseidon.cuh.

`python tools/gen_poseidon_constants.py` rewrites the header; `model_permute` is imported by
tests/test_poseidon_f64_model.py, which checks it against the oracle and the upstream known-answer vectors
and records the largest magnitude any double ever holds (must stay below 2^53).
"""
import hashlib
import os
import re
import struct

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "eth_tx_proof_b200", "csrc", "poseidon_constants.h")
P = 0xFFFFFFFF00000001
CIRC = [17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20]
DIAG0 = 8
FULL_ROUNDS = [0, 1, 2, 3, 26, 27, 28, 29]
PAIR_ROUNDS = list(range(4, 26, 2))
RC_SHA = "d2fcbb5be293c50ab4b1ddcd9c81005b12d689816a54c91a054f97f6588a20a8"


def read_rc():
    txt = open(HEADER).read()
    body = txt.split("#define ETP_POSEIDON_RC_TABLE {", 1)[1].split("}", 1)[0]
    rc = [int(x, 16) for x in re.findall(r"0x([0-9a-f]{16})ULL", body)]
    assert len(rc) == 360
    assert hashlib.sha256(b"".join(struct.pack("<Q", x) for x in rc)).hexdigest() == RC_SHA
    return rc


# ------------------------------------------------------------------------------------------------
def circ_square(c):
    n = len(c)
    return [sum(c[i] * c[(k - i) % n] for i in range(n)) for k in range(n)]


def split_coeffs(c):
    """coefficient vector of a length-12 circulant (out_r = sum_i x_{(i+r)%12} c_i) -> CB[6], CQ[3], CP[3]"""
    assert all((c[k] + c[k + 6]) % 2 == 0 for k in range(6))
    ca = [(c[k] + c[k + 6]) // 2 for k in range(6)]
    cb = [(c[k] - c[k + 6]) // 2 for k in range(6)]
    assert all((ca[k] + ca[k + 3]) % 2 == 0 for k in range(3))
    cp = [(ca[k] + ca[k + 3]) // 2 for k in range(3)]
    cq = [(ca[k] - ca[k + 3]) // 2 for k in range(3)]
    return cb, cq, cp


def column0_split(c):
    """C[:,0] z in the split domain: the column is v_r = c_{(12-r)%12}; returns (PV[3], QV[3], BV[6])."""
    v = [c[(12 - r) % 12] for r in range(12)]
    av = [(v[r] + v[r + 6]) // 2 for r in range(6)]
    bv = [(v[r] - v[r + 6]) // 2 for r in range(6)]
    assert all((v[r] + v[r + 6]) % 2 == 0 for r in range(6)) and all((av[j] + av[j + 3]) % 2 == 0 for j in range(3))
    pv = [(av[j] + av[j + 3]) // 2 for j in range(3)]
    qv = [(av[j] - av[j + 3]) // 2 for j in range(3)]
    return pv, qv, bv


def mod4_planes(k):
    """field element k -> (K_lo, K_hi), K_lo + 2^32 K_hi == k (mod p), both == 0 (mod 4), both >= 0.
    Adding t*p moves (t, t*(2^32 - 1)) into the planes; moving m*2^32 from the high to the low plane
    changes only the high plane's residue."""
    k %= P
    lo, hi = k & 0xFFFFFFFF, k >> 32
    t = (-lo) % 4
    while True:
        l2, h2 = lo + t, hi + t * (2**32 - 1)
        m = h2 % 4
        if h2 - m >= 0:
            l2, h2 = l2 + m * 2**32, h2 - m
            assert l2 % 4 == 0 and h2 % 4 == 0 and (l2 + (h2 << 32)) % P == k
            return l2, h2
        t += 4


def start_values(K):
    """12 plane constants K_r == 0 (mod 4) -> accumulator starts (c[3], q[3], b[6]); the bias 2^52 rides on c."""
    c = [2**52 + (K[j] + K[j + 3] + K[j + 6] + K[j + 9]) // 4 for j in range(3)]
    q = [(K[j] + K[j + 6] - K[j + 3] - K[j + 9]) // 4 for j in range(3)]
    b = [(K[r] - K[r + 6]) // 2 for r in range(6)]
    return c + q + b


def build_tables(rc):
    """FULL[8][2][12], PAIR[11][2][13] as Python ints (may be negative for q/b)."""
    def planes_of_round(r):  # the NEXT round's constants, each as mod-4 planes
        if r + 1 >= 30:
            return [(0, 0)] * 12
        return [mod4_planes(rc[12 * (r + 1) + i]) for i in range(12)]

    full = []
    for r in FULL_ROUNDS:
        pl = planes_of_round(r)
        full.append([start_values([pl[i][h] for i in range(12)]) for h in range(2)])
    pair = []
    for r in PAIR_ROUNDS:
        # y = M s + K1 ; lane 0 replaced by the S-box output n0 ; out = M y' + K2
        k1 = [(rc[12 * (r + 1) + i] & 0xFFFFFFFF, rc[12 * (r + 1) + i] >> 32) for i in range(12)]
        rows = []
        # field-level constant of the pair, then the plane-level bookkeeping: out = C^2 s + C[:,0] z + 8 e0 n0 + Kp
        # with Kp = M K1 + K2 - 8 e0 K1_0 computed on plane integers, re-expressed mod p as mod-4 planes.
        kp_field = []
        for i in range(12):
            mk1 = sum(rc[12 * (r + 1) + (j + i) % 12] * CIRC[j] for j in range(12)) + (DIAG0 * rc[12 * (r + 1)] if i == 0 else 0)
            k2 = rc[12 * (r + 2) + i] if r + 2 < 30 else 0
            corr = 8 * rc[12 * (r + 1)] if i == 0 else 0
            kp_field.append((mk1 + k2 - corr) % P)
        pl = [mod4_planes(k) for k in kp_field]
        for h in range(2):
            rows.append([2**52 + k1[0][h]] + start_values([pl[i][h] for i in range(12)]))
        pair.append(rows)
    return full, pair


# ------------------------------------------------------------------------------------------------
class Track:
    """records the largest |value| that any modelled double holds (bias included)"""
    def __init__(self):
        self.max = 0

    def __call__(self, v):
        if abs(v) > self.max:
            self.max = abs(v)
        return v


def split_apply(x, coef, start, tr, z=None, zc=None):
    """x: 12 unbiased plane values; coef = (CB, CQ, CP); start = 12 ints (c, q, b).  Returns 12 outputs
    (bias 2^52 included).  z/zc: optional extra column term zc = (PV, QV, BV) times the scalar z."""
    cb, cq, cp = coef
    S = [tr(x[k] + x[k + 6]) for k in range(6)]
    D = [tr(x[k] - x[k + 6]) for k in range(6)]
    T = [tr(S[j] + S[j + 3]) for j in range(3)]
    E = [tr(S[j] - S[j + 3]) for j in range(3)]
    Pj, Qj, Br = [], [], []
    for j in range(3):
        acc = start[j]
        for k in range(3):
            acc = tr(acc + cp[k] * T[(k + j) % 3])
        if z is not None:
            acc = tr(acc + zc[0][j] * z)
        Pj.append(acc)
        acc = start[3 + j]
        for k in range(3):
            acc = tr(acc + (-cq[k] if k + j >= 3 else cq[k]) * E[(k + j) % 3])
        if z is not None:
            acc = tr(acc + zc[1][j] * z)
        Qj.append(acc)
    for r in range(6):
        acc = start[6 + r]
        for k in range(6):
            acc = tr(acc + (-cb[k] if k + r >= 6 else cb[k]) * D[(k + r) % 6])
        if z is not None:
            acc = tr(acc + zc[2][r] * z)
        Br.append(acc)
    A = [tr(Pj[j] + Qj[j]) for j in range(3)] + [tr(Pj[j] - Qj[j]) for j in range(3)]
    return [tr(A[r] + Br[r]) for r in range(6)] + [tr(A[r] - Br[r]) for r in range(6)]


def combine(L, H):
    return (L + (H << 32)) % P


def sbox(x):
    return pow(x, 7, P)


_TABLES = None


def tables():
    global _TABLES
    if _TABLES is None:
        rc = read_rc()
        c1 = split_coeffs(CIRC)
        c2 = split_coeffs(circ_square(CIRC))
        _TABLES = (rc, c1, c2, column0_split(CIRC), *build_tables(rc))
    return _TABLES


def model_permute(state, tr=None):
    """The permutation exactly as poseidon.cuh schedules it; every intermediate is a Python int."""
    tr = tr or Track()
    rc, c1, c2, col0, full, pair = tables()
    s = [(state[i] + rc[i]) % P for i in range(12)]
    BIAS = 2**52

    def full_layer(s, f):
        outs = []
        for h in range(2):
            x = [(v >> (32 * h)) & 0xFFFFFFFF for v in s]
            o = split_apply(x, c1, full[f][h], tr)
            o[0] = tr(o[0] + 8 * x[0])
            outs.append([v - BIAS for v in o])
        for v in outs[0] + outs[1]:
            assert 0 <= v < BIAS
        return [combine(outs[0][i], outs[1][i]) for i in range(12)]

    def pair_layer(s, i):
        xs = [[(v >> (32 * h)) & 0xFFFFFFFF for v in s] for h in range(2)]
        us = []
        for h in range(2):
            x = xs[h]
            acc = pair[i][h][0]
            for k in range(12):
                acc = tr(acc + CIRC[k] * x[k])
            us.append(tr(acc + 8 * x[0]) - BIAS)
        n0 = sbox(combine(us[0], us[1]))
        outs = []
        for h in range(2):
            x = xs[h]
            n0p = (n0 >> (32 * h)) & 0xFFFFFFFF
            z = tr(8 * x[0] + tr(n0p - us[h]))
            o = split_apply(x, c2, pair[i][h][1:], tr, z=z, zc=col0)
            o[0] = tr(o[0] + 8 * n0p)
            outs.append([v - BIAS for v in o])
        for v in outs[0] + outs[1]:
            assert 0 <= v < BIAS
        return [combine(outs[0][k], outs[1][k]) for k in range(12)]

    f = 0
    for _ in range(4):
        s = [sbox(v) for v in s]
        s = full_layer(s, f)
        f += 1
    for i in range(11):
        s[0] = sbox(s[0])
        s = pair_layer(s, i)
    for _ in range(4):
        s = [sbox(v) for v in s]
        s = full_layer(s, f)
        f += 1
    return s


def naive_permute(state):
    rc = read_rc()
    s = list(state)
    for r in range(30):
        s = [(s[i] + rc[12 * r + i]) % P for i in range(12)]
        if r < 4 or r >= 26:
            s = [sbox(v) for v in s]
        else:
            s[0] = sbox(s[0])
        s = [(sum(s[(i + k) % 12] * CIRC[i] for i in range(12)) + (DIAG0 * s[0] if k == 0 else 0)) % P for k in range(12)]
    return s


def worst_case_bound():
    """Upper bound of |any modelled double| over ALL inputs (lanes < 2^32 per plane, S-box output planes < 2^32):
    |start| + sum |coef| * max|input| for every accumulator, bias included."""
    rc, c1, c2, col0, full, pair = tables()
    X = 2**32 - 1
    worst = 0
    umax = max(abs(v[h][0]) for v in pair for h in range(2)) + (sum(CIRC) + 8) * X  # biased u accumulator
    zmax = 8 * X + X + (umax - 2**52)
    for coef, starts, z in ((c1, [pl for layer in full for pl in layer], 0), (c2, [pl[1:] for layer in pair for pl in layer], zmax)):
        cb, cq, cp = coef
        for st in starts:
            pj = max(abs(st[j]) for j in range(3)) + sum(abs(v) for v in cp) * 4 * X + max(abs(v) for v in col0[0]) * z
            qj = max(abs(st[3 + j]) for j in range(3)) + sum(abs(v) for v in cq) * 2 * X + max(abs(v) for v in col0[1]) * z
            br = max(abs(st[6 + r]) for r in range(6)) + sum(abs(v) for v in cb) * X + max(abs(v) for v in col0[2]) * z
            worst = max(worst, pj + qj + br + 8 * X)
    return max(worst, umax)


# ------------------------------------------------------------------------------------------------
def f64_bits(v):
    assert float(v) == v and abs(v) < 2**53, v
    return struct.unpack("<Q", struct.pack("<d", float(v)))[0]


def write_header():
    rc, c1, c2, col0, full, pair = tables()
    txt = open(HEADER).read()
    head = txt.split("// ---- FP64-pipe tables", 1)[0].split("// split-cyclic MDS initial values", 1)[0].rstrip() + "\n\n"
    out = [head]
    out.append("// ---- FP64-pipe tables (tools/gen_poseidon_constants.py; layout documented there and in poseidon.cuh)\n")

    def arr(name, vals):
        return f"#define {name} {{{', '.join(str(v) for v in vals)}}}\n"

    for tag, co in (("C1", c1), ("C2", c2)):
        cb, cq, cp = co
        out.append(arr(f"ETP_MDS_{tag}_CB", [f"{v}.0" for v in cb]))
        out.append(arr(f"ETP_MDS_{tag}_CQ", [f"{v}.0" for v in cq]))
        # cyclic-3 part: two equal coefficients a and one odd one at index K  ->  P_j = a*(T0+T1+T2) + (CP_K - a)*T_{(K+j)%3}
        kx = [k for k in range(3) if cp.count(cp[k]) == 1]
        assert len(kx) == 1, cp
        a = cp[(kx[0] + 1) % 3]
        out.append(f"#define ETP_MDS_{tag}_CPA {a}.0\n#define ETP_MDS_{tag}_CPD {cp[kx[0]] - a}.0\n#define ETP_MDS_{tag}_CPK {kx[0]}\n")
    out.append(arr("ETP_MDS_COL0_PV", [f"{v}.0" for v in col0[0]]))
    out.append(arr("ETP_MDS_COL0_QV", [f"{v}.0" for v in col0[1]]))
    out.append(arr("ETP_MDS_COL0_BV", [f"{v}.0" for v in col0[2]]))

    def table(name, rows, comment):
        vals = [f64_bits(v) for layer in rows for plane in layer for v in plane]
        s = f"// {comment}\n#define {name} {{ \\\n"
        for i in range(0, len(vals), 4):
            s += "  " + ", ".join(f"0x{x:016x}ULL" for x in vals[i:i + 4]) + (", \\\n" if i + 4 < len(vals) else " \\\n")
        return s + "}\n"

    out.append(table("ETP_POSEIDON_FULL_F64_TABLE", full, "full-round layers 0..3, 26..29: [8][2 planes][c0 c1 c2 q0 q1 q2 b0..b5] as f64 bit patterns"))
    out.append(table("ETP_POSEIDON_PAIR_F64_TABLE", pair, "partial-round pairs (4,5) .. (24,25): [11][2 planes][u | c0 c1 c2 q0 q1 q2 b0..b5] as f64 bit patterns"))
    open(HEADER, "w").write("".join(out))
    print("wrote", HEADER)


if __name__ == "__main__":
    import random

    rng = random.Random(1)
    tr = Track()
    for trial in range(50):
        st = [rng.randrange(2**64) for _ in range(12)] if trial > 2 else [[0] * 12, list(range(12)), [P - 1] * 12][trial]
        assert model_permute(st, tr) == naive_permute(st), trial
    print("model == naive on 50 states; largest double magnitude 2^%.2f" % __import__("math").log2(tr.max))
    _, c1, c2, col0, _, _ = tables()
    print("C  split:", c1, "\nC^2 split:", c2, "\ncolumn 0:", col0)
    print("worst-case magnitude bound 2^%.2f" % __import__("math").log2(worst_case_bound()))
    if "--write" in __import__("sys").argv:
        write_header()
