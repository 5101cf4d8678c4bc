"""Second slice of the recursion layers: plonky2's circuit prover — `plonky2::plonk::prover::prove`,
plonky2 0.2.2 (/root/reference/Cargo.lock:3441), what every shrink / root / aggregation / block proof of the reference runs
(/root/reference/ops/src/lib.rs:52,72,95; /root/reference/leader/src/prover.rs:26-36) — with the QUOTIENT on the device:

    circuit (gates, selectors, constants, copy constraints -> sigmas)                       host, once per circuit
    wires_commitment = from_values(witness wires)                                           device
    Z / partial products of the permutation argument, their commitment                      device
    compute_quotient_polys: eval_vanishing_poly_base_batch on the 8x LDE, / Z_H, coset iFFT,
        split in quotient_degree_factor chunks; commitment                                  device  (new in this slice)
    openings at zeta / g*zeta, prove_openings over the four-oracle FriInstanceInfo          device

How the vanishing polynomial gets to the device.  plonk/vanishing_poly.rs reduces, with every alpha, ONE list of terms:
`L_0(x)(Z_i(x) - 1)` per challenge, the partial-product checks per challenge, then the gate-constraint slots
`sum_gates filter_gate(x) * constraint_k(x)` (evaluate_gate_constraints; filters from the selector polynomials,
gates/selectors.rs).  That list is recorded once per circuit as a constraint program (cprog.py — the same SSA form the
starky tables use) over a VIRTUAL column space [constants | sigmas | wires | Zs | partial products | X]; the library compiles
it with NVRTC and `etp_compute_quotient_polys_cols_dev` evaluates it over the LDE columns of the three committed oracles plus
the LDE of the polynomial X (the point itself), all read in place through a pointer table.  `reduce_with_powers` weights
term i with alpha^i, the consumer of the quotient kernels weights emission i of N with alpha^(N-1-i): the terms are emitted
in reverse.

Gates (constraints restated from plonky2/src/gates/*.rs): NoopGate, ConstantGate, PublicInputGate, ArithmeticGate,
PoseidonGate (123 constraints of degree 7; the partial rounds in upstream's fast form — `partial_first_constant_layer`,
`mds_partial_layer_init`, `mds_partial_layer_fast` — with the sparse matrices and moved constants derived here from the MDS matrix
and the round constants, checked against the plain round function), ArithmeticExtensionGate,
MulExtensionGate, BaseSumGate<2>, ReducingGate, ReducingExtensionGate, RandomAccessGate, ExponentiationGate, PoseidonMdsGate,
CosetInterpolationGate.  Not built: lookup tables / lookup gates, witness generation by generators (witnesses here are computed directly), zero-knowledge blinding
(off in standard_recursion_config).  The circuit digest is a stand-in (hash of the constants/sigmas cap and degree_bits).
Nothing here can be checked against real plonky2 offline: parity is against the pure-Python evaluation in the tests.
"""
from __future__ import annotations

import time
from typing import Dict, List, Sequence, Tuple

import numpy as np

from . import cprog
from .api import Challenger, Context, FriParams, PolynomialBatch, load_library, poseidon_constants

P = 0xFFFFFFFF00000001
NUM_WIRES, NUM_ROUTED, NUM_CHALLENGES, QUOTIENT_DEGREE_FACTOR = 135, 80, 2, 8
RATE_BITS, CAP_HEIGHT, POW_BITS, NUM_QUERIES = 3, 4, 16, 28
NUM_PARTIAL_PRODUCTS = -(-NUM_ROUTED // QUOTIENT_DEGREE_FACTOR) - 1  # 9
UNUSED_SELECTOR = (1 << 32) - 1


def root_of_unity(n_log: int) -> int:
    return pow(1753635133440165772, 1 << (32 - n_log), P)


def coset_shifts(num_shifts: int) -> List[int]:
    """plonky2::field::cosets::get_unique_coset_shifts: k_j = g^j for the multiplicative generator g = 7."""
    return [pow(7, j, P) for j in range(num_shifts)]


# ---- Poseidon over Python ints (witness generation of PoseidonGate rows; the product hashes on the GPU) ---------------------
_PC = None


def _pc():
    global _PC
    if _PC is None:
        _PC = poseidon_constants()
    return _PC


def mds_layer(state, add=None, mulc=None):
    """res[r] = sum_i state[(i + r) % 12] * MDS_MATRIX_CIRC[i] + state[r] * MDS_MATRIX_DIAG[r] (poseidon.rs mds_row_shf)."""
    _, circ, diag = _pc()
    add = add or (lambda x, y: (x + y) % P)
    mulc = mulc or (lambda x, c: x * c % P)
    out = []
    for r in range(12):
        acc = None
        for i in range(12):
            t = mulc(state[(i + r) % 12], circ[i])
            acc = t if acc is None else add(acc, t)
        if diag[r]:
            acc = add(acc, mulc(state[r], diag[r]))
        out.append(acc)
    return out


def mds_matrix() -> List[List[int]]:
    """M[r][c]: out[r] = sum_c M[r][c] in[c] (column-vector form of mds_row_shf)."""
    _, circ, diag = _pc()
    m = [[0] * 12 for _ in range(12)]
    for r in range(12):
        for i in range(12):
            m[r][(i + r) % 12] = (m[r][(i + r) % 12] + circ[i]) % P
        m[r][r] = (m[r][r] + diag[r]) % P
    return m


def _mat_inv(a: List[List[int]]) -> List[List[int]]:
    """Inverse of a square matrix over the field (Gauss-Jordan, Python ints)."""
    n = len(a)
    m = [list(row) + [1 if i == j else 0 for j in range(n)] for i, row in enumerate(a)]
    for col in range(n):
        piv = next(r for r in range(col, n) if m[r][col] % P)
        m[col], m[piv] = m[piv], m[col]
        inv = pow(m[col][col], P - 2, P)
        m[col] = [x * inv % P for x in m[col]]
        for r in range(n):
            if r != col and m[r][col]:
                f = m[r][col]
                m[r] = [(x - f * y) % P for x, y in zip(m[r], m[col])]
    return [row[n:] for row in m]


def _mat_mul(a, b):
    return [[sum(a[i][k] * b[k][j] for k in range(len(b))) % P for j in range(len(b[0]))] for i in range(len(a))]


_FAST = None


def fast_partial_round_constants():
    """The partial rounds in the form of plonky2's `partial_first_constant_layer` / `mds_partial_layer_init` /
    `mds_partial_layer_fast` (poseidon.rs; Poseidon paper, appendix B), derived here from the MDS matrix and the round
    constants rather than copied:
      constants: c_r = M d  =>  M S(x) + c_r = M (S(x) + d): d[1:] moves in front of the previous S-box, d[0] stays behind it;
      matrices:  D M = S diag(1, H) with D = diag(1, A) the dense part pushed back from the next round, H = A M^, and the
                 sparse S = [[m00, a^T H^-1], [A b, I]]; the last dense part, diag(1, H_0), is applied once before the first S-box.
    -> (first_constants[12], scalar_constants[21], initial_matrix[11][11], w_hats[22][11], vs[22][11], m00)."""
    global _FAST
    if _FAST is not None:
        return _FAST
    rc, _, _ = _pc()
    M = mds_matrix()
    M_inv = _mat_inv(M)
    c = [list(rc[4 + r]) for r in range(22)]
    k = [0] * 22  # k[r]: added to element 0 after the S-box of partial round r - 1
    for r in range(21, 0, -1):
        d = [sum(M_inv[i][j] * c[r][j] for j in range(12)) % P for i in range(12)]
        c[r - 1] = [c[r - 1][0]] + [(c[r - 1][i] + d[i]) % P for i in range(1, 12)]
        k[r] = d[0]
    m00 = M[0][0]
    a = [M[0][j] for j in range(1, 12)]
    b = [M[i][0] for i in range(1, 12)]
    M_hat = [[M[i][j] for j in range(1, 12)] for i in range(1, 12)]
    A = [[1 if i == j else 0 for j in range(11)] for i in range(11)]
    w_hats, vs = [None] * 22, [None] * 22
    for r in range(21, -1, -1):
        H = _mat_mul(A, M_hat)
        H_inv = _mat_inv(H)
        w_hats[r] = [sum(a[i] * H_inv[i][j] for i in range(11)) % P for j in range(11)]
        vs[r] = [sum(A[i][j] * b[j] for j in range(11)) % P for i in range(11)]
        A = H
    _FAST = (c[0], k[1:], A, w_hats, vs, m00)
    return _FAST


class _IntRing:
    add = staticmethod(lambda x, y: (x + y) % P)
    addc = staticmethod(lambda x, c: (x + c) % P)
    mulc = staticmethod(lambda x, c: x * c % P)
    mul = staticmethod(lambda x, y: x * y % P)


class _ExprRing:
    def __init__(self, b):
        self.b = b

    add = staticmethod(lambda x, y: x + y)
    mul = staticmethod(lambda x, y: x * y)

    def addc(self, x, c):
        return x + self.b.const(c)

    def mulc(self, x, c):
        return x * self.b.const(c)


def partial_rounds_fast(state, ring, sbox_in_hook=None):
    """The 22 partial rounds in the fast form over a ring (_IntRing / _ExprRing).  sbox_in_hook(r, state0) -> the value whose 7th
    power continues (the gate records `state0 - wire` and returns the wire); None: plain evaluation."""
    first, ks, init, w_hats, vs, m00 = fast_partial_round_constants()

    def sbox(x):
        x2 = ring.mul(x, x)
        x4 = ring.mul(x2, x2)
        return ring.mul(x4, ring.mul(x2, x))

    state = [ring.addc(x, first[i]) for i, x in enumerate(state)]  # partial_first_constant_layer
    rest = []  # mds_partial_layer_init: element 0 untouched, diag(1, H_0) on the others
    for i in range(11):
        acc = None
        for j in range(11):
            t = ring.mulc(state[1 + j], init[i][j])
            acc = t if acc is None else ring.add(acc, t)
        rest.append(acc)
    state = [state[0]] + rest
    for r in range(22):
        x0 = state[0] if sbox_in_hook is None else sbox_in_hook(r, state[0])
        x0 = sbox(x0)
        if r < 21:
            x0 = ring.addc(x0, ks[r])
        d = ring.mulc(x0, m00)  # mds_partial_layer_fast: d = [m00 | w_hat] . state ; state[i] += x0 * v[i]
        for j in range(11):
            d = ring.add(d, ring.mulc(state[1 + j], w_hats[r][j]))
        state = [d] + [ring.add(state[1 + j], ring.mulc(x0, vs[r][j])) for j in range(11)]
    return state


def poseidon_gate_wires(inputs: Sequence[int], swap: int) -> List[int]:
    """All 135 wires of one PoseidonGate row (gates/poseidon.rs PoseidonGenerator::run_once): inputs, outputs, swap, deltas
    and the S-box inputs of every round after the first.  Computed by the library's host code (etp_host_poseidon_gate_wires:
    a recursion layer's witness is tens of thousands of these rows); poseidon_gate_wires_py is the same in Python integers."""
    import ctypes as C

    a = (C.c_uint64 * 12)(*[int(x) % P for x in inputs])
    out = (C.c_uint64 * NUM_WIRES)()
    load_library().etp_host_poseidon_gate_wires(a, int(swap), out)
    return list(out)


def poseidon_gate_wires_py(inputs: Sequence[int], swap: int) -> List[int]:
    """poseidon_gate_wires restated with Python integers (the definition the C helper is tested against)."""
    rc, _, _ = _pc()
    w = [0] * NUM_WIRES
    inputs = [int(x) % P for x in inputs]
    w[0:12] = inputs
    w[PoseidonGate.WIRE_SWAP] = swap
    state = list(inputs)
    for i in range(4):
        delta = swap * (inputs[i + 4] - inputs[i]) % P
        w[PoseidonGate.wire_delta(i)] = delta
        state[i] = (inputs[i] + delta) % P
        state[i + 4] = (inputs[i + 4] - delta) % P
    rnd = 0
    for r in range(4):
        state = [(s + c) % P for s, c in zip(state, rc[rnd])]
        if r != 0:
            for i in range(12):
                w[PoseidonGate.wire_full_sbox_0(r, i)] = state[i]
        state = mds_layer([pow(s, 7, P) for s in state])
        rnd += 1
    for r in range(22):
        state = [(s + c) % P for s, c in zip(state, rc[rnd])]
        w[PoseidonGate.wire_partial_sbox(r)] = state[0]
        state[0] = pow(state[0], 7, P)
        state = mds_layer(state)
        rnd += 1
    for r in range(4):
        state = [(s + c) % P for s, c in zip(state, rc[rnd])]
        for i in range(12):
            w[PoseidonGate.wire_full_sbox_1(r, i)] = state[i]
        state = mds_layer([pow(s, 7, P) for s in state])
        rnd += 1
    w[12:24] = state
    return w


def hash_no_pad(elements: Sequence[int]) -> List[int]:
    """PoseidonHash::hash_no_pad (overwrite-mode sponge, rate 8) with the library's host permutation."""
    import ctypes as C

    L = load_library()
    st = (C.c_uint64 * 12)()
    elements = [int(x) % P for x in elements]
    for off in range(0, len(elements), 8):
        for k, v in enumerate(elements[off:off + 8]):
            st[k] = v
        L.etp_host_poseidon_permute(st)
    return [int(st[k]) for k in range(4)]


# ---- gates ---------------------------------------------------------------------------------------------------------------------
class Gate:
    """eval(b, wire, const, pi) -> the gate's constraints as cprog expressions (Gate::eval_unfiltered); wire(j) / const(j):
    local wire j / the gate's j-th constant (selectors already stripped), pi(i): public_inputs_hash[i]."""
    name, degree, num_constants, num_constraints = "gate", 0, 0, 0

    def eval(self, b, wire, const, pi):
        return []

    def id(self):
        return self.name


class NoopGate(Gate):
    name = "NoopGate"


class ConstantGate(Gate):
    """gates/constant.rs: wire_i - const_i."""

    def __init__(self, num_consts: int = 2):
        self.name, self.degree, self.num_constants, self.num_constraints = f"ConstantGate {{ num_consts: {num_consts} }}", 1, num_consts, num_consts

    def eval(self, b, wire, const, pi):
        return [const(i) - wire(i) for i in range(self.num_constants)]


class PublicInputGate(Gate):
    """gates/public_input.rs: wires 0..4 - public_inputs_hash."""
    name, degree, num_constants, num_constraints = "PublicInputGate", 1, 0, 4

    def eval(self, b, wire, const, pi):
        return [wire(i) - pi(i) for i in range(4)]


class ArithmeticGate(Gate):
    """gates/arithmetic_base.rs: per operation  output - (multiplicand_0 * multiplicand_1 * const_0 + addend * const_1)."""

    def __init__(self, num_ops: int = 20):
        self.num_ops = num_ops
        self.name, self.degree, self.num_constants, self.num_constraints = f"ArithmeticGate {{ num_ops: {num_ops} }}", 3, 2, num_ops

    def eval(self, b, wire, const, pi):
        out = []
        for i in range(self.num_ops):
            m0, m1, addend, output = wire(4 * i), wire(4 * i + 1), wire(4 * i + 2), wire(4 * i + 3)
            out.append(output - (m0 * m1 * const(0) + addend * const(1)))
        return out


class PoseidonGate(Gate):
    """gates/poseidon.rs: one permutation per row, with an optional swap of the two 4-element input halves (Merkle paths)."""
    name, degree, num_constants, num_constraints = "PoseidonGate", 7, 0, 123
    fast_partial = True
    WIRE_SWAP = 24
    START_DELTA, START_FULL_0, START_PARTIAL, START_FULL_1 = 25, 29, 65, 87

    @staticmethod
    def wire_delta(i):
        return PoseidonGate.START_DELTA + i

    @staticmethod
    def wire_full_sbox_0(rnd, i):  # rnd 1..3
        return PoseidonGate.START_FULL_0 + 12 * (rnd - 1) + i

    @staticmethod
    def wire_partial_sbox(rnd):
        return PoseidonGate.START_PARTIAL + rnd

    @staticmethod
    def wire_full_sbox_1(rnd, i):
        return PoseidonGate.START_FULL_1 + 12 * rnd + i

    def eval(self, b, wire, const, pi):
        rc, _, _ = _pc()
        cons = []
        swap = wire(self.WIRE_SWAP)
        cons.append(swap * (swap - 1))
        state = [None] * 12
        for i in range(4):
            lhs, rhs, delta = wire(i), wire(i + 4), wire(self.wire_delta(i))
            cons.append(swap * (rhs - lhs) - delta)
            state[i], state[i + 4] = lhs + delta, rhs - delta
        for i in range(8, 12):
            state[i] = wire(i)

        def sbox(x):
            x2 = x * x
            x4 = x2 * x2
            return x4 * (x2 * x)

        def mds(st):
            return mds_layer(st, add=lambda x, y: x + y, mulc=lambda x, c: x * b.const(c))

        rnd = 0
        for r in range(4):
            state = [s + b.const(c) for s, c in zip(state, rc[rnd])]
            if r != 0:
                for i in range(12):
                    sbox_in = wire(self.wire_full_sbox_0(r, i))
                    cons.append(state[i] - sbox_in)
                    state[i] = sbox_in
            state = mds([sbox(s) for s in state])
            rnd += 1
        def partial_hook(r, state0):  # constraint on the S-box input of partial round r, then continue from the wire
            sbox_in = wire(self.wire_partial_sbox(r))
            cons.append(state0 - sbox_in)
            return sbox_in

        if self.fast_partial:  # partial_first_constant_layer, mds_partial_layer_init, mds_partial_layer_fast (poseidon.rs)
            state = partial_rounds_fast(state, _ExprRing(b), partial_hook)
        else:  # the plain round function: the same polynomials, 6x the multiplications
            for r in range(22):
                state = [s + b.const(c) for s, c in zip(state, rc[rnd + r])]
                state[0] = sbox(partial_hook(r, state[0]))
                state = mds(state)
        rnd += 22
        for r in range(4):
            state = [s + b.const(c) for s, c in zip(state, rc[rnd])]
            for i in range(12):
                sbox_in = wire(self.wire_full_sbox_1(r, i))
                cons.append(state[i] - sbox_in)
                state[i] = sbox_in
            state = mds([sbox(s) for s in state])
            rnd += 1
        for i in range(12):
            cons.append(state[i] - wire(12 + i))
        assert len(cons) == self.num_constraints
        return cons


# quadratic extension (X^2 = 7) over pairs — Python ints or cprog expressions alike
def _e_mul(a, b, seven=7):
    return (a[0] * b[0] + a[1] * b[1] * seven, a[0] * b[1] + a[1] * b[0])


def _e_add(a, b):
    return (a[0] + b[0], a[1] + b[1])


def _e_sub(a, b):
    return (a[0] - b[0], a[1] - b[1])


def _e_mod(a):
    return (a[0] % P, a[1] % P)


class ArithmeticExtensionGate(Gate):
    """gates/arithmetic_extension.rs: per operation, over the quadratic extension: output - (m0 * m1 * c0 + addend * c1)."""

    def __init__(self, num_ops: int = 10):
        self.num_ops = num_ops
        self.name, self.degree, self.num_constants, self.num_constraints = f"ArithmeticExtensionGate {{ num_ops: {num_ops} }}", 3, 2, 2 * num_ops

    def eval(self, b, wire, const, pi):
        out = []
        pair = lambda k: (wire(k), wire(k + 1))
        for i in range(self.num_ops):
            m0, m1, addend, output = (pair(8 * i + 2 * j) for j in range(4))
            prod = _e_mul(m0, m1, b.const(7))
            out += list(_e_sub(output, _e_add((prod[0] * const(0), prod[1] * const(0)), (addend[0] * const(1), addend[1] * const(1)))))
        return out

    def witness(self, rnd, c0, c1):
        w = []
        for _ in range(self.num_ops):
            m0, m1, ad = (rnd(), rnd()), (rnd(), rnd()), (rnd(), rnd())
            pr = _e_mod(_e_mul(m0, m1))
            w += [*m0, *m1, *ad, (pr[0] * c0 + ad[0] * c1) % P, (pr[1] * c0 + ad[1] * c1) % P]
        return w


class MulExtensionGate(Gate):
    """gates/multiplication_extension.rs: output - m0 * m1 * c0 over the extension."""

    def __init__(self, num_ops: int = 13):
        self.num_ops = num_ops
        self.name, self.degree, self.num_constants, self.num_constraints = f"MulExtensionGate {{ num_ops: {num_ops} }}", 3, 1, 2 * num_ops

    def eval(self, b, wire, const, pi):
        out = []
        pair = lambda k: (wire(k), wire(k + 1))
        for i in range(self.num_ops):
            m0, m1, output = (pair(6 * i + 2 * j) for j in range(3))
            prod = _e_mul(m0, m1, b.const(7))
            out += list(_e_sub(output, (prod[0] * const(0), prod[1] * const(0))))
        return out

    def witness(self, rnd, c0):
        w = []
        for _ in range(self.num_ops):
            m0, m1 = (rnd(), rnd()), (rnd(), rnd())
            pr = _e_mod(_e_mul(m0, m1))
            w += [*m0, *m1, pr[0] * c0 % P, pr[1] * c0 % P]
        return w


class BaseSumGate(Gate):
    """gates/base_sum.rs with B = 2: wire 0 = sum_i limb_i 2^i, every limb boolean."""

    def __init__(self, num_limbs: int = 63):
        self.num_limbs = num_limbs
        self.name, self.degree, self.num_constants, self.num_constraints = f"BaseSumGate {{ num_limbs: {num_limbs} }} + Base: 2", 2, 0, 1 + num_limbs

    def eval(self, b, wire, const, pi):
        limbs = [wire(1 + i) for i in range(self.num_limbs)]
        acc = None
        for l in reversed(limbs):  # reduce_with_powers(limbs, 2)
            acc = l if acc is None else acc * b.const(2) + l
        return [acc - wire(0)] + [l * (l - 1) for l in limbs]

    def witness(self, value):
        assert 0 <= value < 1 << self.num_limbs
        return [value] + [(value >> i) & 1 for i in range(self.num_limbs)]


class ReducingGate(Gate):
    """gates/reducing.rs: Horner over base-field coefficients with an extension alpha: acc_i = acc_{i-1} * alpha + coeff_i."""

    def __init__(self, num_coeffs: int = 43):
        self.num_coeffs = num_coeffs
        self.name, self.degree, self.num_constants, self.num_constraints = f"ReducingGate {{ num_coeffs: {num_coeffs} }}", 2, 0, 2 * num_coeffs
        self.start_coeffs, self.start_accs = 6, 6 + num_coeffs

    def _acc(self, i):  # wires of accumulator i; the last one is the output
        return (0, 1) if i == self.num_coeffs - 1 else (self.start_accs + 2 * i, self.start_accs + 2 * i + 1)

    def eval(self, b, wire, const, pi):
        alpha, acc = (wire(2), wire(3)), (wire(4), wire(5))
        out = []
        for i in range(self.num_coeffs):
            nxt = tuple(wire(k) for k in self._acc(i))
            t = _e_mul(acc, alpha, b.const(7))
            out += [t[0] + wire(self.start_coeffs + i) - nxt[0], t[1] - nxt[1]]
            acc = nxt
        return out

    def witness(self, rnd):
        w = [0] * NUM_WIRES
        alpha, acc = (rnd(), rnd()), (rnd(), rnd())
        w[2:6] = [*alpha, *acc]
        for i in range(self.num_coeffs):
            c = rnd()
            w[self.start_coeffs + i] = c
            t = _e_mod(_e_mul(acc, alpha))
            acc = ((t[0] + c) % P, t[1])
            a0, a1 = self._acc(i)
            w[a0], w[a1] = acc
        return w


class ReducingExtensionGate(Gate):
    """gates/reducing_extension.rs: the same with extension coefficients."""

    def __init__(self, num_coeffs: int = 32):
        self.num_coeffs = num_coeffs
        self.name, self.degree, self.num_constants, self.num_constraints = f"ReducingExtensionGate {{ num_coeffs: {num_coeffs} }}", 2, 0, 2 * num_coeffs
        self.start_coeffs, self.start_accs = 6, 6 + 2 * num_coeffs

    def _acc(self, i):
        return (0, 1) if i == self.num_coeffs - 1 else (self.start_accs + 2 * i, self.start_accs + 2 * i + 1)

    def eval(self, b, wire, const, pi):
        alpha, acc = (wire(2), wire(3)), (wire(4), wire(5))
        out = []
        for i in range(self.num_coeffs):
            nxt = tuple(wire(k) for k in self._acc(i))
            coeff = (wire(self.start_coeffs + 2 * i), wire(self.start_coeffs + 2 * i + 1))
            t = _e_add(_e_mul(acc, alpha, b.const(7)), coeff)
            out += [t[0] - nxt[0], t[1] - nxt[1]]
            acc = nxt
        return out

    def witness(self, rnd):
        w = [0] * NUM_WIRES
        alpha, acc = (rnd(), rnd()), (rnd(), rnd())
        w[2:6] = [*alpha, *acc]
        for i in range(self.num_coeffs):
            c = (rnd(), rnd())
            w[self.start_coeffs + 2 * i], w[self.start_coeffs + 2 * i + 1] = c
            acc = _e_mod(_e_add(_e_mul(acc, alpha), c))
            a0, a1 = self._acc(i)
            w[a0], w[a1] = acc
        return w


class RandomAccessGate(Gate):
    """gates/random_access.rs: num_copies look-ups claimed_element = list[access_index] in lists of 2^bits items (index bits
    as advice wires, the list folded bit by bit), plus num_extra_constants constant wires."""

    def __init__(self, bits: int = 4, num_copies: int = 4, num_extra_constants: int = 2):
        self.bits, self.num_copies, self.num_extra = bits, num_copies, num_extra_constants
        self.vec = 1 << bits
        self.name = f"RandomAccessGate {{ bits: {bits}, num_copies: {num_copies}, num_extra_constants: {num_extra_constants} }}"
        self.degree, self.num_constants = bits + 1, num_extra_constants
        self.num_constraints = num_copies * (bits + 2) + num_extra_constants
        self.start_extra = (2 + self.vec) * num_copies
        self.num_routed = self.start_extra + num_extra_constants

    def eval(self, b, wire, const, pi):
        out = []
        for c in range(self.num_copies):
            base = (2 + self.vec) * c
            index, claimed = wire(base), wire(base + 1)
            items = [wire(base + 2 + j) for j in range(self.vec)]
            bits = [wire(self.num_routed + c * self.bits + j) for j in range(self.bits)]
            out += [x * (x - 1) for x in bits]
            acc = None
            for x in reversed(bits):
                acc = x if acc is None else acc * b.const(2) + x
            out.append(acc - index)
            for x in bits:  # fold pairs: x + b (y - x)
                items = [items[k] + x * (items[k + 1] - items[k]) for k in range(0, len(items), 2)]
            out.append(items[0] - claimed)
        out += [const(i) - wire(self.start_extra + i) for i in range(self.num_extra)]
        return out

    def witness(self, rnd, rng, extra):
        w = [0] * NUM_WIRES
        for c in range(self.num_copies):
            base = (2 + self.vec) * c
            items = [rnd() for _ in range(self.vec)]
            idx = int(rng.integers(0, self.vec))
            w[base], w[base + 1] = idx, items[idx]
            w[base + 2:base + 2 + self.vec] = items
            for j in range(self.bits):
                w[self.num_routed + c * self.bits + j] = (idx >> j) & 1
        for i, v in enumerate(extra):
            w[self.start_extra + i] = v
        return w


class ExponentiationGate(Gate):
    """gates/exponentiation.rs: output = base^power by square-and-multiply over the power's bits (most significant first)."""

    def __init__(self, num_power_bits: int = 66):
        self.n_bits = num_power_bits
        self.name, self.degree, self.num_constants, self.num_constraints = f"ExponentiationGate {{ num_power_bits: {num_power_bits} }}", 4, 0, num_power_bits + 1

    def eval(self, b, wire, const, pi):
        base, output = wire(0), wire(1 + self.n_bits)
        inter = [wire(2 + self.n_bits + i) for i in range(self.n_bits)]
        one = b.const(1)
        out = []
        for i in range(self.n_bits):
            prev = one if i == 0 else inter[i - 1] * inter[i - 1]
            bit = wire(1 + self.n_bits - 1 - i)
            out.append(prev * (bit * base + one - bit) - inter[i])
        out.append(output - inter[-1])
        return out

    def witness(self, base, power):
        w = [0] * NUM_WIRES
        w[0] = base
        bits = [(power >> i) & 1 for i in range(self.n_bits)]
        w[1:1 + self.n_bits] = bits
        cur = 1
        for i in range(self.n_bits):
            prev = 1 if i == 0 else cur * cur % P
            cur = prev * (base if bits[self.n_bits - 1 - i] else 1) % P
            w[2 + self.n_bits + i] = cur
        w[1 + self.n_bits] = cur
        assert cur == pow(base, power, P)
        return w


class PoseidonMdsGate(Gate):
    """gates/poseidon_mds.rs: the MDS layer on 12 extension elements."""
    name, degree, num_constants, num_constraints = "PoseidonMdsGate", 1, 0, 24

    def eval(self, b, wire, const, pi):
        out = []
        for comp in range(2):
            ins = [wire(2 * i + comp) for i in range(12)]
            res = mds_layer(ins, add=lambda x, y: x + y, mulc=lambda x, c: x * b.const(c))
            out += [(i, comp, res[i] - wire(24 + 2 * i + comp)) for i in range(12)]
        return [e for _, _, e in sorted(out, key=lambda t: (t[0], t[1]))]

    def witness(self, rnd):
        ins = [(rnd(), rnd()) for _ in range(12)]
        outs = list(zip(mds_layer([x[0] for x in ins]), mds_layer([x[1] for x in ins])))
        return [v for x in ins for v in x] + [v for x in outs for v in x]


class CosetInterpolationGate(Gate):
    """gates/coset_interpolation.rs: the value at `evaluation_point` of the polynomial that interpolates 2^subgroup_bits
    extension values on the coset shift * H — barycentric form without divisions, evaluated in chunks so that the constraint
    degree stays `degree`: with x' = evaluation_point / shift (a wire, constrained by x' * shift = evaluation_point),
        (eval, prod) <- (eval * (x' - x_i) + w_i v_i prod, prod * (x' - x_i))      for the points of a chunk,
    the (eval, prod) pair after every chunk but the last stored in `intermediate` wires, the last eval = evaluation_value.
    The FRI verifier inside the recursive circuits computes its `compute_evaluation` with this gate."""

    @classmethod
    def with_max_degree(cls, subgroup_bits: int, max_degree: int):
        """The smallest degree that needs no more intermediates than max_degree does (a larger selector group fits then)."""
        n_points = 1 << subgroup_bits
        n_intermediates = (n_points - 2) // (max_degree - 1)
        return cls(subgroup_bits, (n_points - 2) // (n_intermediates + 1) + 2)

    def __init__(self, subgroup_bits: int = 4, degree: int = 6):
        self.subgroup_bits, self.num_points, self.deg = subgroup_bits, 1 << subgroup_bits, degree
        self.num_intermediates = (self.num_points - 2) // (degree - 1)
        self.name = f"CosetInterpolationGate {{ subgroup_bits: {subgroup_bits}, degree: {degree} }}"
        self.degree, self.num_constants, self.num_constraints = degree, 0, 2 + 2 * 2 * self.num_intermediates + 2
        self.start_values = 1
        self.start_point = 1 + 2 * self.num_points
        self.start_value = self.start_point + 2
        self.start_intermediates = self.start_value + 2
        self.start_shifted = self.start_intermediates + 4 * self.num_intermediates
        g = root_of_unity(subgroup_bits)
        self.domain = [pow(g, i, P) for i in range(self.num_points)]
        self.weights = []
        for i, xi in enumerate(self.domain):  # barycentric weights 1 / prod_{j != i} (x_i - x_j)
            d = 1
            for j, xj in enumerate(self.domain):
                if i != j:
                    d = d * (xi - xj) % P
            self.weights.append(pow(d, P - 2, P))

    def _chunks(self):
        out, start = [range(0, self.deg)], 0
        for i in range(self.num_intermediates):
            start = 1 + (self.deg - 1) * (i + 1)
            out.append(range(start, min(start + self.deg - 1, self.num_points)))
        return out

    def _wire_eval(self, i):
        return self.start_intermediates + 2 * i

    def _wire_prod(self, i):
        return self.start_intermediates + 2 * (self.num_intermediates + i)

    def eval(self, b, wire, const, pi):
        pair = lambda k: (wire(k), wire(k + 1))
        seven = b.const(7)
        shift, point, shifted = wire(0), pair(self.start_point), pair(self.start_shifted)
        cons = [point[0] - shifted[0] * shift, point[1] - shifted[1] * shift]
        values = [pair(self.start_values + 2 * i) for i in range(self.num_points)]
        zero, one = b.const(0), b.const(1)

        def partial(idx, ev, pr):
            for i in idx:
                term = (shifted[0] - b.const(self.domain[i]), shifted[1])
                wv = (values[i][0] * b.const(self.weights[i]), values[i][1] * b.const(self.weights[i]))
                ev = _e_add(_e_mul(ev, term, seven), _e_mul(wv, pr, seven))
                pr = _e_mul(pr, term, seven)
            return ev, pr

        chunks = self._chunks()
        ev, pr = partial(chunks[0], (zero, zero), (one, zero))
        for i in range(self.num_intermediates):
            iev, ipr = pair(self._wire_eval(i)), pair(self._wire_prod(i))
            cons += [iev[0] - ev[0], iev[1] - ev[1], ipr[0] - pr[0], ipr[1] - pr[1]]
            ev, pr = partial(chunks[i + 1], iev, ipr)
        out_v = pair(self.start_value)
        cons += [out_v[0] - ev[0], out_v[1] - ev[1]]
        assert len(cons) == self.num_constraints
        return cons

    def witness(self, rnd):
        w = [0] * NUM_WIRES
        shift = rnd() or 1
        values = [(rnd(), rnd()) for _ in range(self.num_points)]
        point = (rnd(), rnd())
        s_inv = pow(shift, P - 2, P)
        shifted = (point[0] * s_inv % P, point[1] * s_inv % P)
        w[0] = shift
        for i, v in enumerate(values):
            w[self.start_values + 2 * i], w[self.start_values + 2 * i + 1] = v
        w[self.start_point], w[self.start_point + 1] = point
        w[self.start_shifted], w[self.start_shifted + 1] = shifted
        ev, pr = (0, 0), (1, 0)
        chunks = self._chunks()
        for ci, idx in enumerate(chunks):
            for i in idx:
                term = ((shifted[0] - self.domain[i]) % P, shifted[1])
                wv = (values[i][0] * self.weights[i] % P, values[i][1] * self.weights[i] % P)
                ev = _e_mod(_e_add(_e_mul(ev, term), _e_mul(wv, pr)))
                pr = _e_mod(_e_mul(pr, term))
            if ci < self.num_intermediates:
                w[self._wire_eval(ci)], w[self._wire_eval(ci) + 1] = ev
                w[self._wire_prod(ci)], w[self._wire_prod(ci) + 1] = pr
        w[self.start_value], w[self.start_value + 1] = ev
        self._last = (shift, values, point, ev)
        return w


# ---- selectors (gates/selectors.rs) -----------------------------------------------------------------------------------------------
def selector_groups(gates: Sequence[Gate], max_degree: int) -> List[range]:
    """Greedy grouping of the (degree-sorted) gates: a group of `size` gates costs a filter of degree size - 1 (+ 1 for the
    UNUSED factor when there are several groups), so size + degree of its last gate must stay below max_degree."""
    n = len(gates)
    if max(g.degree for g in gates) + n - 1 <= max_degree:
        return [range(0, n)]
    groups, start = [], 0
    while start < n:
        size = 0
        while start + size < n and size + gates[start + size].degree < max_degree:
            size += 1
        assert size > 0, "a gate of this degree does not fit the quotient degree factor"
        groups.append(range(start, start + size))
        start += size
    return groups


class Circuit:
    """CommonCircuitData + ProverOnlyCircuitData, as far as the prover steps here need them."""

    # virtual column space of the vanishing program
    def col_const(self, j):
        return j

    def col_sigma(self, j):
        return self.num_constants + j

    def col_wire(self, j):
        return self.num_constants + NUM_ROUTED + j

    def col_z(self, i):
        return self.num_constants + NUM_ROUTED + NUM_WIRES + i

    def col_pp(self, i, k):
        return self.num_constants + NUM_ROUTED + NUM_WIRES + NUM_CHALLENGES + i * NUM_PARTIAL_PRODUCTS + k

    @property
    def col_x(self):
        return self.num_constants + NUM_ROUTED + NUM_WIRES + NUM_CHALLENGES * (1 + NUM_PARTIAL_PRODUCTS)

    @property
    def num_virtual_columns(self):
        return self.col_x + 1

    def __init__(self, degree_bits: int, gates: List[Gate], gate_of_row: np.ndarray, gate_constants: np.ndarray, sigmas: np.ndarray,
                 num_public_inputs: int):
        self.degree_bits, self.n = degree_bits, 1 << degree_bits
        self.gates = gates
        self.groups = selector_groups(gates, QUOTIENT_DEGREE_FACTOR)
        self.num_selectors = len(self.groups)
        self.num_gate_constants = max(g.num_constants for g in gates)
        self.num_constants = self.num_selectors + self.num_gate_constants
        self.num_gate_constraints = max(g.num_constraints for g in gates)
        self.num_public_inputs = num_public_inputs
        self.k_is = coset_shifts(NUM_ROUTED)
        self.gate_of_row = gate_of_row
        # constants = selector polynomials ++ gate constants
        consts = np.zeros((self.num_constants, self.n), dtype=np.uint64)
        for s, grp in enumerate(self.groups):
            in_group = (gate_of_row >= grp.start) & (gate_of_row < grp.stop)
            consts[s] = np.where(in_group, gate_of_row, UNUSED_SELECTOR).astype(np.uint64)
        consts[self.num_selectors:] = gate_constants
        self.constants, self.sigmas = consts, sigmas
        self.program = self._vanishing_program()

    def selector_index(self, gate_index: int) -> int:
        return next(s for s, grp in enumerate(self.groups) if gate_index in grp)

    def _vanishing_program(self) -> cprog.Program:
        """eval_vanishing_poly as ONE constraint program (module docstring).  CH 0..1 = betas, CH 2..3 = gammas; PI 0..3 =
        public_inputs_hash; X = the point (a virtual column)."""
        b = cprog.ProgramBuilder(self.num_virtual_columns, 4, QUOTIENT_DEGREE_FACTOR + 1)
        one = b.const(1)
        x = b.lv(self.col_x)
        wire = lambda j: b.lv(self.col_wire(j))
        # ---- evaluate_gate_constraints: slot k = sum over gates of filter * constraint k
        slots = [None] * self.num_gate_constraints
        many = self.num_selectors > 1
        for gi, gate in enumerate(self.gates):
            if gate.num_constraints == 0:
                continue
            s_idx = self.selector_index(gi)
            s = b.lv(self.col_const(s_idx))
            filt = None
            for i in self.groups[s_idx]:  # compute_filter: prod_{i in group, i != row} (i - s) [* (UNUSED - s)]
                if i != gi:
                    f = b.const(i) - s
                    filt = f if filt is None else filt * f
            if many:
                f = b.const(UNUSED_SELECTOR) - s
                filt = f if filt is None else filt * f
            const = lambda j: b.lv(self.col_const(self.num_selectors + j))
            for k, c in enumerate(gate.eval(b, wire, const, b.pi)):
                t = c if filt is None else filt * c
                slots[k] = t if slots[k] is None else slots[k] + t
        zero = b.const(0)
        slots = [zero if s is None else s for s in slots]
        # ---- permutation argument
        z_1_terms, pp_terms = [], []
        for i in range(NUM_CHALLENGES):
            beta, gamma = b.challenge(i), b.challenge(NUM_CHALLENGES + i)
            z_x, z_gx = b.lv(self.col_z(i)), b.nv(self.col_z(i))
            z_1_terms.append(("first", z_x - one))  # L_0(x) (Z(x) - 1): the consumer's first-row selector is L_0
            num = [wire(j) + beta * (x * b.const(self.k_is[j])) + gamma for j in range(NUM_ROUTED)]
            den = [wire(j) + beta * b.lv(self.col_sigma(j)) + gamma for j in range(NUM_ROUTED)]
            accs = [z_x] + [b.lv(self.col_pp(i, k)) for k in range(NUM_PARTIAL_PRODUCTS)] + [z_gx]
            for c in range(0, NUM_ROUTED, QUOTIENT_DEGREE_FACTOR):  # check_partial_products
                prod = lambda xs: xs[0] if len(xs) == 1 else prod(xs[:len(xs) // 2]) * prod(xs[len(xs) // 2:])
                k = c // QUOTIENT_DEGREE_FACTOR
                pp_terms.append(("all", accs[k] * prod(num[c:c + 8]) - accs[k + 1] * prod(den[c:c + 8])))
        terms = z_1_terms + pp_terms + [("all", s) for s in slots]
        for kind, e in reversed(terms):  # reduce_with_powers: term i * alpha^i
            (b.first_row if kind == "first" else b.constraint)(e)
        self.num_vanishing_terms = len(terms)
        return b.build()

    def virtual_trace(self, wires: np.ndarray, zs_pp: np.ndarray) -> np.ndarray:
        """[constants | sigmas | wires | Zs | partial products | X] on the trace domain (tests: Program.check_trace)."""
        g = root_of_unity(self.degree_bits)
        xs = np.zeros(self.n, dtype=np.uint64)
        cur = 1
        for i in range(self.n):
            xs[i] = cur
            cur = cur * g % P
        return np.concatenate([self.constants, self.sigmas, wires, zs_pp, xs[None, :]], axis=0)


def gl_mul_vec(a: np.ndarray, b) -> np.ndarray:
    """Element-wise product mod p of canonical uint64 arrays (b: array or int), numpy only: 32-bit limbs, 2^64 = 2^32 - 1,
    2^96 = -1 (mod p)."""
    M, EPS = np.uint64(0xFFFFFFFF), np.uint64(0xFFFFFFFF)
    a = np.asarray(a, dtype=np.uint64)
    b = np.asarray(b, dtype=np.uint64) if isinstance(b, np.ndarray) else np.uint64(int(b) % P)
    s32 = np.uint64(32)
    with np.errstate(over="ignore"):
        a0, a1, b0, b1 = a & M, a >> s32, b & M, b >> s32
        ll, lh, hl, hh = a0 * b0, a0 * b1, a1 * b0, a1 * b1
        mid = lh + hl
        carry = (mid < lh).astype(np.uint64)
        hi = hh + (mid >> s32) + (carry << s32)
        lo = ((mid & M) << s32) + ll
        hi = hi + (lo < ll).astype(np.uint64)
        h0, h1 = hi & M, hi >> s32
        t0 = lo - h1
        t0 = t0 - np.where(lo < h1, EPS, np.uint64(0))
        t1 = h0 * EPS
        t2 = t0 + t1
        t2 = t2 + np.where(t2 < t1, EPS, np.uint64(0))
        return np.where(t2 >= np.uint64(P), t2 - np.uint64(P), t2)


def _components(n_nodes: int, edges: np.ndarray) -> np.ndarray:
    """Connected-component label of every node (scipy's union-find when available, a plain one otherwise)."""
    try:
        from scipy.sparse import coo_matrix
        from scipy.sparse.csgraph import connected_components

        m = coo_matrix((np.ones(len(edges), dtype=np.int8), (edges[:, 0], edges[:, 1])), shape=(n_nodes, n_nodes))
        return connected_components(m, directed=False)[1].astype(np.int64)
    except ImportError:
        parent = list(range(n_nodes))

        def find(u):
            while parent[u] != u:
                parent[u] = parent[parent[u]]
                u = parent[u]
            return u

        for a, b_ in edges.tolist():
            ra, rb = find(a), find(b_)
            if ra != rb:
                parent[ra] = rb
        return np.array([find(u) for u in range(n_nodes)], dtype=np.int64)


class NoWitness(AssertionError):
    """The circuit being built has no satisfying witness for the given data (e.g. the inner proof of a verifier circuit is invalid)."""


class CircuitBuilder:
    """Rows of gates with their constants, copy constraints between routed wires, direct wire assignment (no generators)."""

    def __init__(self):
        self.rows: List[Tuple[Gate, List[int]]] = []
        self.wires: List[List[int]] = []
        self.copies: List[Tuple[Tuple[int, int], Tuple[int, int]]] = []
        self.public_inputs: List[int] = []

    def add_gate(self, gate: Gate, constants: Sequence[int] = (), wires: Sequence[int] = None) -> int:
        self.rows.append((gate, [int(c) % P for c in constants]))
        w = [0] * NUM_WIRES
        if wires is not None:
            w[:len(wires)] = [int(v) % P for v in wires]
        self.wires.append(w)
        return len(self.rows) - 1

    def connect(self, a: Tuple[int, int], b_: Tuple[int, int]):
        """Copy constraint a == b.  The values must already agree: this is how an invalid inner proof is rejected while a verifier
        circuit is built (NoWitness is an AssertionError, raised explicitly so that `python -O` cannot strip the check)."""
        if a[1] >= NUM_ROUTED or b_[1] >= NUM_ROUTED:
            raise ValueError("only routed wires can be copy-constrained")
        if self.wires[a[0]][a[1]] != self.wires[b_[0]][b_[1]]:
            raise NoWitness(f"copy constraint between different values {a} {b_}")
        self.copies.append((a, b_))

    def build(self, min_degree_bits: int = 0):
        """-> (Circuit, wires (135, n)).  Gates are sorted by (degree, id) as CircuitBuilder::build does; rows are padded with
        NoopGates to a power of two."""
        n_rows = max(len(self.rows), 2)
        degree_bits = max((n_rows - 1).bit_length(), min_degree_bits)
        n = 1 << degree_bits
        noop = next((g for g, _ in self.rows if isinstance(g, NoopGate)), NoopGate())
        rows = self.rows + [(noop, [])] * (n - len(self.rows))
        wires = np.zeros((NUM_WIRES, n), dtype=np.uint64)
        wires[:, :len(self.wires)] = np.array(self.wires, dtype=np.uint64).T
        uniq: Dict[str, Gate] = {}
        for g, _ in rows:
            uniq.setdefault(g.id(), g)
        gates = sorted(uniq.values(), key=lambda g: (g.degree, g.id()))
        index = {g.id(): i for i, g in enumerate(gates)}
        gate_of_row = np.array([index[g.id()] for g, _ in rows], dtype=np.int64)
        n_gc = max(g.num_constants for g in gates)
        gate_constants = np.zeros((n_gc, n), dtype=np.uint64)
        for r, (g, cs) in enumerate(rows):
            for j, c in enumerate(cs):
                gate_constants[j, r] = c
        # copy constraints -> partition of the routed wire positions -> sigma: every position maps to the next of its set
        # (members in ascending position order, the last one back to the first).  position = column * n + row
        g = root_of_unity(degree_bits)
        xs = np.ones(n, dtype=np.uint64)  # xs[i] = g^i, by doubling
        for k in range(degree_bits):
            xs[1 << k:2 << k] = gl_mul_vec(xs[:1 << k], pow(g, 1 << k, P))
        k_is = coset_shifts(NUM_ROUTED)
        ident = np.concatenate([gl_mul_vec(xs, k_is[c]) for c in range(NUM_ROUTED)])
        sig = ident.copy()
        if self.copies:
            cp = np.array([(c0 * n + r0, c1 * n + r1) for (r0, c0), (r1, c1) in self.copies], dtype=np.int64)
            pos, inv = np.unique(cp.ravel(), return_inverse=True)
            label = _components(len(pos), inv.reshape(-1, 2))
            order = np.lexsort((pos, label))
            p_sorted, l_sorted = pos[order], label[order]
            nxt = np.roll(p_sorted, -1)
            last = np.nonzero(l_sorted != np.roll(l_sorted, -1))[0]          # last member of every set
            first = np.concatenate([[0], last[:-1] + 1])                      # first member of every set
            nxt[last] = p_sorted[first]
            sig[p_sorted] = ident[nxt]
        circuit = Circuit(degree_bits, gates, gate_of_row, gate_constants, sig.reshape(NUM_ROUTED, n), len(self.public_inputs))
        return circuit, wires


def hash_chain_circuit(degree_bits: int, seed: int = 1, poseidon_fraction: float = 0.6, arithmetic_fraction: float = 0.3, all_gates: bool = False,
                       extra_rows: int = 1, witness_seed: int = None):
    """A recursion-verifier-shaped synthetic circuit with its witness: the public inputs are hashed in-circuit (PoseidonGate
    row wired to the PublicInputGate), a Merkle-path-like chain of swapped PoseidonGates, chains of ArithmeticGate operations,
    constants through a ConstantGate — all linked by copy constraints — padded with NoopGates to 2^degree_bits rows.
    all_gates: also `extra_rows` rows of each of ArithmeticExtension, MulExtension, BaseSum, Exponentiation (its exponent bits wired
    to the base-sum limbs), Reducing, ReducingExtension, RandomAccess, PoseidonMds and CosetInterpolation gates (selector groups by
upstream's greedy rule).
    witness_seed: the gate constants (the circuit) come from `seed`, every witness value from `witness_seed` — different
    witnesses of ONE circuit; None: one stream for both.
    -> (Circuit, wires (135, n), public_inputs)"""
    rng = np.random.default_rng(seed if witness_seed is None else (seed, witness_seed))
    rng_c = rng if witness_seed is None else np.random.default_rng(seed)
    n = 1 << degree_bits
    cb = CircuitBuilder()
    rnd = lambda: int(rng.integers(0, 2**63)) % P
    rnd_c = lambda: int(rng_c.integers(0, 2**63)) % P
    public_inputs = [rnd() for _ in range(8)]
    cb.public_inputs = list(public_inputs)
    pi_hash = hash_no_pad(public_inputs)
    r_pi = cb.add_gate(PublicInputGate(), wires=pi_hash)
    r_const = cb.add_gate(ConstantGate(2), constants=[0, 1], wires=[0, 1])
    pos = PoseidonGate()
    w = poseidon_gate_wires(public_inputs + [0, 0, 0, 0], 0)
    r_h = cb.add_gate(pos, wires=w)
    assert w[12:16] == pi_hash
    for i in range(4):
        cb.connect((r_h, 12 + i), (r_pi, i))
        cb.connect((r_h, 8 + i), (r_const, 0))
    cb.connect((r_h, PoseidonGate.WIRE_SWAP), (r_const, 0))
    budget = n - len(cb.rows) - (9 * extra_rows if all_gates else 0)
    n_pos = max(int(budget * poseidon_fraction), 1)
    n_ar = max(int(budget * arithmetic_fraction), 1)
    prev_row, digest = r_h, w[12:16]
    for _ in range(n_pos):  # Merkle-path shape: state = hash(swap ? (sibling, state) : (state, sibling))
        swap = int(rng.integers(0, 2))
        sibling = [rnd() for _ in range(4)]
        w = poseidon_gate_wires(list(digest) + sibling + [0, 0, 0, 0], swap)
        r = cb.add_gate(pos, wires=w)
        for i in range(4):
            cb.connect((r, i), (prev_row, 12 + i))
            cb.connect((r, 8 + i), (r_const, 0))
        prev_row, digest = r, w[12:16]
    ar = ArithmeticGate(20)
    prev = None
    for _ in range(n_ar):
        c0, c1 = rnd_c(), rnd_c()
        w = [0] * (4 * ar.num_ops)
        r = len(cb.rows)
        links = []
        for i in range(ar.num_ops):
            m0 = prev[2] if prev is not None else rnd()
            m1, addend = rnd(), rnd()
            out = (m0 * m1 % P * c0 + addend * c1) % P
            w[4 * i:4 * i + 4] = [m0, m1, addend, out]
            if prev is not None:
                links.append(((r, 4 * i), (prev[0], prev[1])))
            prev = (r, 4 * i + 3, out)
        cb.add_gate(ar, constants=[c0, c1], wires=w)
        for a, b_ in links:
            cb.connect(a, b_)
    if all_gates:  # a few rows of every other gate of the recursive verifier's gate set (CosetInterpolationGate excepted)
        for _ in range(extra_rows):
            c0, c1 = rnd_c(), rnd_c()
            g = ArithmeticExtensionGate(10)
            r_ae = cb.add_gate(g, constants=[c0, c1], wires=g.witness(rnd, c0, c1))
            g = MulExtensionGate(13)
            w = g.witness(rnd, c0)
            w[0:2] = cb.wires[r_ae][6:8]  # first product's m0 = the extension output of the row above
            pr = _e_mod(_e_mul((w[0], w[1]), (w[2], w[3])))
            w[4:6] = [pr[0] * c0 % P, pr[1] * c0 % P]
            r_me = cb.add_gate(g, constants=[c0], wires=w)
            cb.connect((r_me, 0), (r_ae, 6))
            cb.connect((r_me, 1), (r_ae, 7))
            g = BaseSumGate(63)
            value = int(rng.integers(0, 2**62))
            r_bs = cb.add_gate(g, wires=g.witness(value))
            g = ExponentiationGate(66)
            w = g.witness(rnd(), value)
            r_ex = cb.add_gate(g, wires=w)
            for i in range(62):  # the exponent's bits are the base-sum limbs
                cb.connect((r_ex, 1 + i), (r_bs, 1 + i))
            cb.add_gate(ReducingGate(43), wires=ReducingGate(43).witness(rnd))
            cb.add_gate(ReducingExtensionGate(32), wires=ReducingExtensionGate(32).witness(rnd))
            g = RandomAccessGate(4, 4, 2)
            e0, e1 = rnd_c(), rnd_c()
            cb.add_gate(g, constants=[e0, e1], wires=g.witness(rnd, rng, [e0, e1]))
            cb.add_gate(PoseidonMdsGate(), wires=PoseidonMdsGate().witness(rnd))
            g = CosetInterpolationGate.with_max_degree(4, QUOTIENT_DEGREE_FACTOR)  # degree 6, two intermediate (eval, prod) pairs
            cb.add_gate(g, wires=g.witness(rnd))
    return (*cb.build(degree_bits), public_inputs)


class _MerkleGadget:
    """verify_merkle_proof_to_cap (plonky2 hash/merkle_proofs.rs) on a CircuitBuilder: shared gates, the constant-zero wire,
    Poseidon sponges with overwrite-mode carry, and one opening = index bits (BaseSumGate) + leaf digest (hash_or_noop) + one
    swapped PoseidonGate per level + RandomAccessGate selection of the cap entry, all tied by copy constraints."""

    def __init__(self, cb: "CircuitBuilder"):
        self.cb = cb
        self.pos, self.bs, self.ra = PoseidonGate(), BaseSumGate(63), RandomAccessGate(4, 4, 2)
        self.r_const = cb.add_gate(ConstantGate(2), constants=[0, 1], wires=[0, 1])
        self.zero = (self.r_const, 0)

    def sponge(self, elements):
        """hash_n_to_hash_no_pad over `elements` with PoseidonGate rows (overwrite mode: lanes a chunk does not overwrite keep
        the previous output).  -> (row of the last permutation, the wire each element entered through)."""
        cb, prev, entered = self.cb, None, []
        for off in range(0, len(elements), 8):
            chunk = [int(x) % P for x in elements[off:off + 8]]
            state = list(chunk) + ([0] * (8 - len(chunk)) if prev is None else cb.wires[prev][12 + len(chunk):20]) + \
                ([0, 0, 0, 0] if prev is None else cb.wires[prev][20:24])
            r = cb.add_gate(self.pos, wires=poseidon_gate_wires(state, 0))
            cb.connect((r, PoseidonGate.WIRE_SWAP), self.zero)
            for k in range(len(chunk), 12):
                cb.connect((r, k), self.zero if prev is None else (prev, 12 + k))
            entered += [(r, k) for k in range(len(chunk))]
            prev = r
        return prev, entered

    def public_inputs(self, values):
        """Registers `values` as the public inputs: hashed in-circuit, the digest wired to a PublicInputGate.  -> their wires."""
        cb = self.cb
        cb.public_inputs = [int(x) % P for x in values]
        r_pi = cb.add_gate(PublicInputGate(), wires=hash_no_pad(cb.public_inputs))
        r_h, wires = self.sponge(cb.public_inputs)
        for i in range(4):
            cb.connect((r_h, 12 + i), (r_pi, i))
        return wires

    def opening(self, leaf, leaf_index, siblings, cap, cap_wires):
        """One verify_merkle_proof_to_cap.  cap_wires[4 e + c]: the wire of element c of cap entry e (public-input wires).
        -> (the wire holding leaf_index, the wires the leaf's elements entered through)."""
        cb, zero, bs, ra, pos = self.cb, self.zero, self.bs, self.ra, self.pos
        h = (len(cap) - 1).bit_length()
        assert len(cap) == 1 << h == 16, "RandomAccessGate(bits = 4): cap_height 4"
        levels = len(siblings)
        r_bits = cb.add_gate(bs, wires=bs.witness(leaf_index))
        cap_index = leaf_index >> levels
        assert cap_index < (1 << h)
        r_hi = cb.add_gate(bs, wires=bs.witness(cap_index))
        for j in range(63):
            cb.connect((r_hi, 1 + j), (r_bits, 1 + levels + j) if j < h else zero)
        for j in range(levels + h, 63):
            cb.connect((r_bits, 1 + j), zero)
        leaf = [int(x) % P for x in leaf]
        if len(leaf) <= 4:  # hash_or_noop: the digest is the zero-padded leaf (advice wires of a NoopGate row)
            r_leaf = cb.add_gate(NoopGate(), wires=leaf + [0] * (4 - len(leaf)))
            for k in range(len(leaf), 4):
                cb.connect((r_leaf, k), zero)
            cur, leaf_wires = (r_leaf, 0), [(r_leaf, k) for k in range(len(leaf))]
        else:
            r_leaf, leaf_wires = self.sponge(leaf)
            cur = (r_leaf, 12)
        digest = cb.wires[cur[0]][cur[1]:cur[1] + 4]
        for lvl, sib in enumerate(siblings):
            bit = (leaf_index >> lvl) & 1
            r = cb.add_gate(pos, wires=poseidon_gate_wires(list(digest) + [int(x) % P for x in sib] + [0, 0, 0, 0], bit))
            for k in range(4):
                cb.connect((r, k), (cur[0], cur[1] + k))
                cb.connect((r, 8 + k), zero)
            cb.connect((r, PoseidonGate.WIRE_SWAP), (r_bits, 1 + lvl))
            cur, digest = (r, 12), cb.wires[r][12:16]
        w = [0] * NUM_WIRES
        for c in range(4):
            base = (2 + ra.vec) * c
            w[base], w[base + 1] = cap_index, int(cap[cap_index][c]) % P
            w[base + 2:base + 2 + ra.vec] = [int(cap[e][c]) % P for e in range(ra.vec)]
            for j in range(ra.bits):
                w[ra.num_routed + c * ra.bits + j] = (cap_index >> j) & 1
        r_ra = cb.add_gate(ra, constants=[0, 0], wires=w)
        for c in range(4):
            base = (2 + ra.vec) * c
            cb.connect((r_ra, base), (r_hi, 0))
            cb.connect((r_ra, base + 1), (cur[0], cur[1] + c))  # fails unless the path leads to the cap
            for e in range(ra.vec):
                cb.connect((r_ra, base + 2 + e), cap_wires[4 * e + c])
        return (r_bits, 0), leaf_wires


def merkle_openings_circuit(openings: Sequence[tuple], min_degree_bits: int = 0):
    """A fragment of the recursive verifier as a circuit: `verify_merkle_proof_to_cap` (plonky2 hash/merkle_proofs.rs, the gadget
    every FRI query of a recursive proof runs) for a list of openings (leaf, leaf_index, siblings, cap) — real data in, e.g.
    rows, paths and caps of batches committed on the device, or EVERY Merkle check of the FRI verification of an inner proof
    (28 queries x (4 initial trees + the layer trees): the Poseidon-dominated bulk of a recursive verifier circuit).
      * the distinct caps are the public inputs: hashed in-circuit by a PoseidonGate sponge wired to the PublicInputGate;
      * per opening: see _MerkleGadget.opening.
    -> (Circuit, wires, public_inputs).  Building fails (AssertionError in connect) when a path does not lead to its cap:
    no witness exists."""
    cb = CircuitBuilder()
    gadget = _MerkleGadget(cb)
    caps, cap_id = [], []
    for _, _, _, cap in openings:
        key = tuple(int(x) for d in cap for x in d)
        if key not in caps:
            caps.append(key)
        cap_id.append(caps.index(key))
    wires = gadget.public_inputs([x for key in caps for x in key])
    for (leaf, index, siblings, cap), cid in zip(openings, cap_id):
        gadget.opening(leaf, int(index), siblings, cap, wires[64 * cid:64 * (cid + 1)])
    return (*cb.build(min_degree_bits), list(cb.public_inputs))


def merkle_proof_circuit(leaf: Sequence[int], leaf_index: int, siblings: Sequence[Sequence[int]], cap: Sequence[Sequence[int]],
                         min_degree_bits: int = 0):
    """merkle_openings_circuit for ONE opening of a committed batch."""
    return merkle_openings_circuit([(leaf, leaf_index, siblings, cap)], min_degree_bits)


def fri_fold_check_circuit(values: Sequence[Sequence[int]], coset_start: int, beta: Sequence[int], expected: Sequence[int],
                           min_degree_bits: int = 0):
    """Another fragment of the recursive verifier: one `compute_evaluation` of the FRI verifier (plonky2 fri/verifier.rs — the
    arity-16 fold consistency check of a query step) as a circuit: a CosetInterpolationGate interpolates the 16 extension
    values of a coset coset_start * <w16> (natural order) and evaluates at beta; the result must equal `expected` (the
    next layer's opened value).  Public inputs: beta and expected, hashed in-circuit and wired to the PublicInputGate.
    -> (Circuit, wires, public_inputs); AssertionError when the fold is inconsistent (no witness)."""
    g = CosetInterpolationGate.with_max_degree(4, QUOTIENT_DEGREE_FACTOR)
    assert len(values) == g.num_points
    cb = CircuitBuilder()
    r_const = cb.add_gate(ConstantGate(2), constants=[0, 1], wires=[0, 1])
    zero = (r_const, 0)
    public_inputs = [int(beta[0]) % P, int(beta[1]) % P, int(expected[0]) % P, int(expected[1]) % P]
    cb.public_inputs = list(public_inputs)
    r_pi = cb.add_gate(PublicInputGate(), wires=hash_no_pad(public_inputs))
    r_h = cb.add_gate(PoseidonGate(), wires=poseidon_gate_wires(public_inputs + [0] * 8, 0))
    cb.connect((r_h, PoseidonGate.WIRE_SWAP), zero)
    for k in range(4, 12):
        cb.connect((r_h, k), zero)
    for i in range(4):
        cb.connect((r_h, 12 + i), (r_pi, i))
    vals = [(int(v[0]) % P, int(v[1]) % P) for v in values]
    seq = iter([int(coset_start) % P] + [x for v in vals for x in v] + public_inputs[:2])
    w = g.witness(lambda: next(seq))  # witness() draws shift, the 16 values, then the point, in this order
    r_g = cb.add_gate(g, wires=w)
    cb.connect((r_g, g.start_point), (r_h, 0))
    cb.connect((r_g, g.start_point + 1), (r_h, 1))
    cb.connect((r_g, g.start_value), (r_h, 2))       # fails unless interpolate(values)(beta) == expected
    cb.connect((r_g, g.start_value + 1), (r_h, 3))
    return (*cb.build(min_degree_bits), public_inputs)


# ---- the prover -----------------------------------------------------------------------------------------------------------------
class CircuitProver:
    """etp_circuit (include/etp_b200.h): per-circuit state built once (ProverOnlyCircuitData / CommonCircuitData — the
    constants/sigmas commitment, the digest, the compiled vanishing program, the LDE of X); `prove(wires, public_inputs)` is ONE
    C-ABI call that runs plonk::prover::prove's steps on the device and returns the proof."""

    def __init__(self, ctx: Context, circuit: Circuit, circuit_digest: Sequence[int] = None):
        import ctypes as C

        self.ctx, self.c = ctx, circuit
        self.fri_params = FriParams.make(circuit.degree_bits, RATE_BITS, CAP_HEIGHT, POW_BITS, NUM_QUERIES)
        u64 = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.uint64))
        ptr = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint64))
        prog, consts, sig, k_is = u64(circuit.program.words), u64(circuit.constants), u64(circuit.sigmas), u64(circuit.k_is)
        dig = u64(circuit_digest) if circuit_digest is not None else None
        h = C.c_void_p()
        ctx.check(ctx.L.etp_circuit_create(ctx.h, ptr(prog), prog.size, ptr(consts), circuit.num_constants, ptr(sig), ptr(k_is), NUM_ROUTED, NUM_WIRES,
                                           circuit.degree_bits, QUOTIENT_DEGREE_FACTOR, NUM_CHALLENGES, C.byref(self.fri_params),
                                           ptr(dig) if dig is not None else None, C.byref(h)))
        self.h = h
        d = np.zeros(4, dtype=np.uint64)
        ctx.check(ctx.L.etp_circuit_digest(h, ptr(d)))
        self.digest = [int(x) for x in d]
        cap = np.zeros((1 << CAP_HEIGHT, 4), dtype=np.uint64)
        ctx.check(ctx.L.etp_circuit_constants_sigmas_cap(h, ptr(cap)))
        self.constants_sigmas_cap = cap
        self.proof_words = int(ctx.L.etp_circuit_proof_words(h))

    def prove_words(self, wires: np.ndarray, public_inputs: Sequence[int]) -> np.ndarray:
        """-> the flat "B200PLK1" proof (wire.parse_circuit_proof)."""
        import ctypes as C

        w = np.ascontiguousarray(np.asarray(wires, dtype=np.uint64))
        assert w.shape == (NUM_WIRES, self.c.n)
        pi_hash = np.array(hash_no_pad(public_inputs), dtype=np.uint64)
        out = np.zeros(self.proof_words, dtype=np.uint64)
        ptr = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint64))
        self.ctx.check(self.ctx.L.etp_circuit_prove_host(self.h, ptr(w), ptr(pi_hash), ptr(out)))
        return out

    def prove(self, wires: np.ndarray, public_inputs: Sequence[int]) -> dict:
        """-> the proof as a dict (the layout oracle.circuit_prove and tests/plonk_verifier.py use) + per-phase device times."""
        from . import wire

        t0 = time.perf_counter()
        words = self.prove_words(wires, public_inputs)
        total = (time.perf_counter() - t0) * 1e3
        p = wire.parse_circuit_proof(words)
        op = p["openings"]
        ms = dict(self.ctx.last_prove_timings())
        ms["total"] = total
        return {"degree_bits": self.c.degree_bits, "public_inputs": [int(x) % P for x in public_inputs], "words": words,
                "wires_cap": p["wires_cap"], "plonk_zs_partial_products_cap": p["plonk_zs_partial_products_cap"],
                "quotient_polys_cap": p["quotient_polys_cap"],
                "openings": {"constants_sigmas": np.concatenate([op["constants"], op["plonk_sigmas"]]), "wires": op["wires"],
                             "zs_partial_products": np.concatenate([op["plonk_zs"], op["partial_products"]]),
                             "quotient_polys": op["quotient_polys"], "plonk_zs_next": op["plonk_zs_next"]},
                "opening_proof": p["opening_proof"], "ms": ms}

    def __del__(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.L.etp_circuit_free(self.h)
            self.h = None


# ---- recursion: the Merkle checks of the verification of an inner proof, as the witness of an outer circuit ----------------------
def fri_query_openings(prover: CircuitProver, words: np.ndarray, public_inputs: Sequence[int]) -> List[tuple]:
    """Replays the verifier's transcript of a circuit proof (plonk/get_challenges.rs) far enough to know the FRI query indices
    and returns EVERY Merkle opening the FRI verifier checks: per query the rows of the four initial oracles and of every
    commit-phase layer, with their paths and caps — [(leaf, leaf_index, siblings, cap)], the input of
    `merkle_openings_circuit`.  This is witness generation for a recursive verifier circuit, not a verification: nothing is
    checked here."""
    from . import wire

    c = prover.c
    p = wire.parse_circuit_proof(words)
    h = p["header"]
    op = p["openings"]
    ch = Challenger()
    ch.observe(prover.digest)
    ch.observe(hash_no_pad(public_inputs))
    ch.observe_cap(p["wires_cap"])
    ch.get_n_challenges(2 * NUM_CHALLENGES)
    ch.observe_cap(p["plonk_zs_partial_products_cap"])
    ch.get_n_challenges(NUM_CHALLENGES)
    ch.observe_cap(p["quotient_polys_cap"])
    ch.get_extension_challenge()
    for k in ("constants", "plonk_sigmas", "wires", "plonk_zs", "partial_products", "quotient_polys", "plonk_zs_next"):
        ch.observe(op[k])
    ch.get_extension_challenge()  # FRI alpha
    fri = [int(x) for x in p["opening_proof"]]
    capw = 4 << h["cap_height"]
    n_layers, log_lde = h["n_fri_layers"], h["degree_bits"] + h["rate_bits"]
    pos = 0

    def take(n):
        nonlocal pos
        out = fri[pos:pos + n]
        pos += n
        return out

    layer_caps = []
    for _ in range(n_layers):
        cap = take(capw)
        layer_caps.append([cap[4 * i:4 * i + 4] for i in range(capw // 4)])
        ch.observe(cap)
        ch.get_extension_challenge()
    shapes = [h["num_constants"] + h["num_routed_wires"], h["num_wires"], h["num_challenges"] * (1 + h["num_partial_products"]),
              h["num_challenges"] * h["quotient_degree_factor"]]
    per_query = sum(n + 4 * (log_lde - h["cap_height"]) for n in shapes)
    bits = log_lde
    for _ in range(n_layers):
        bits -= h["arity_bits"]
        per_query += 2 * (1 << h["arity_bits"]) + 4 * (bits - h["cap_height"])
    queries_at = pos
    pos += per_query * h["num_queries"]
    ch.observe(take(2 * h["final_poly_len"]))
    ch.observe(take(1))  # the proof-of-work witness
    ch.get_challenge()   # the proof-of-work response
    indices = [ch.get_challenge() % (1 << log_lde) for _ in range(h["num_queries"])]
    caps = [[[int(x) for x in row] for row in np.asarray(cp).reshape(-1, 4)]
            for cp in (prover.constants_sigmas_cap, p["wires_cap"], p["plonk_zs_partial_products_cap"], p["quotient_polys_cap"])]
    pos = queries_at
    out = []
    quads = lambda ws: [ws[4 * i:4 * i + 4] for i in range(len(ws) // 4)]
    for x in indices:
        for n, cap in zip(shapes, caps):
            leaf = take(n)
            out.append((leaf, x, quads(take(4 * (log_lde - h["cap_height"]))), cap))
        bits, idx = log_lde, x
        for layer in range(n_layers):
            bits -= h["arity_bits"]
            idx >>= h["arity_bits"]
            leaf = take(2 * (1 << h["arity_bits"]))
            out.append((leaf, idx, quads(take(4 * (bits - h["cap_height"]))), layer_caps[layer]))
    return out


def recursive_merkle_verifier_circuit(inner: Sequence[tuple], min_degree_bits: int = 0):
    """The Merkle part of a recursive verifier: one outer circuit that checks every Merkle opening of the FRI verification of the
    inner proofs `[(CircuitProver, proof words, public inputs)]` — 28 queries x (4 initial trees + the layer trees) each.  A 2^12-row
    inner proof gives ~3.2 k outer rows (2^12), two of them 2^13: the sizes, and the PoseidonGate share, of the reference's
    shrinking / aggregation circuits.  Not the whole verifier: the FRI fold checks (`fri_fold_check_circuit`), fri_combine_initial,
    the in-circuit challenger and the vanishing-polynomial check are not assembled into it."""
    openings = []
    for prover, words, public_inputs in inner:
        openings += fri_query_openings(prover, words, public_inputs)
    return merkle_openings_circuit(openings, min_degree_bits)
