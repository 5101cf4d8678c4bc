"""The FRI verifier as a circuit (plonky2 0.2.2 fri/recursive_verifier.rs verify_fri_proof: the bulk of every shrinking /
aggregation / block circuit of the reference, /root/reference/ops/src/lib.rs:52,72,95), assembled from the gates of circuit.py and
fed with a REAL inner proof made on the device.  Per query round, for an inner circuit proof (four initial oracles):

    fri_verify_initial_proof   Merkle openings of the four oracle rows            PoseidonGate sponges + swapped path levels,
                                                                                  BaseSum index bits, RandomAccess cap entry
    subgroup_x                 g * w^rev(x_index)                                 ExponentiationGate over the index bits
    fri_combine_initial        sum_b alpha^.. (reduce(alpha, evals_b) - y_b) / (x - z_b)   ReducingGate rows, extension
                                                                                  arithmetic (division = witnessed quotient)
    per commit-phase layer     evals[x_index mod 16] == previous value            RandomAccessGate on the 16 opened values
                               compute_evaluation at beta                         CosetInterpolationGate (coset start by an
                                                                                  ExponentiationGate over the low index bits)
                               Merkle opening of the layer row, x^16              as above; ArithmeticGate squarings
    final polynomial           Horner at subgroup_x == last value                 ReducingExtensionGate

The challenges (alpha, betas, zeta, the query indices), the reduced openings, the caps and the final polynomial are PUBLIC INPUTS
of the outer circuit (hashed in-circuit): the in-circuit challenger that derives them from the transcript, the proof-of-work
check and the vanishing-polynomial check of the inner circuit are the parts of plonky2's recursive verifier NOT assembled here.
The builder assigns wire values directly while it places gates (no generators): witness generation and circuit construction
are one pass, and the resulting circuit's structure (gates, constants, copy constraints) does not depend on the proof's data.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

from . import circuit as cc
from .circuit import (NUM_WIRES, P, ArithmeticExtensionGate, ArithmeticGate, BaseSumGate, CircuitBuilder, ConstantGate,
                      CosetInterpolationGate, ExponentiationGate, RandomAccessGate, ReducingExtensionGate, ReducingGate, _MerkleGadget,
                      _e_mod, _e_mul)

Target = Tuple[int, int]          # (row, wire)
ExtTarget = Tuple[Target, Target]


def _bitrev(x: int, bits: int) -> int:
    return int(format(x, f"0{bits}b")[::-1], 2) if bits else 0


def _e_inv(a):
    n = (a[0] * a[0] - 7 * a[1] * a[1]) % P
    ni = pow(n, P - 2, P)
    return (a[0] * ni % P, (P - a[1]) * ni % P)


class GadgetBuilder(CircuitBuilder):
    """Targets over a CircuitBuilder: every helper places (or reuses a free slot of) a gate row, assigns the wire values,
    connects the operands by copy constraints and returns the target(s) of the result — plonky2's CircuitBuilder helpers
    (arithmetic, arithmetic_extension, split_le, exp_from_bits, random_access, reduce_with_powers, interpolate_coset)."""

    def __init__(self):
        super().__init__()
        self.merkle = _MerkleGadget(self)
        self.zero, self.one = self.merkle.zero, (self.merkle.r_const, 1)
        self._consts = {0: self.zero, 1: self.one}
        self._const_row = None      # (row, next free slot) of a ConstantGate with free slots
        self._arith = {}            # (c0, c1) -> (row, next op) of an ArithmeticGate row with free ops
        self._arith_ext = {}
        self.ar, self.are = ArithmeticGate(20), ArithmeticExtensionGate(10)
        self.bs, self.exp, self.ra = BaseSumGate(63), ExponentiationGate(66), RandomAccessGate(4, 4, 2)
        self.red, self.rede = ReducingGate(43), ReducingExtensionGate(32)
        self.coset = CosetInterpolationGate.with_max_degree(4, cc.QUOTIENT_DEGREE_FACTOR)

    # ---- values
    def val(self, t: Target) -> int:
        return self.wires[t[0]][t[1]]

    def vale(self, t: ExtTarget):
        return (self.val(t[0]), self.val(t[1]))

    def _set(self, t: Target, v: int, src: Target = None):
        self.wires[t[0]][t[1]] = int(v) % P
        if src is not None:
            self.connect(t, src)

    def constant(self, v: int) -> Target:
        v = int(v) % P
        if v in self._consts:
            return self._consts[v]
        # ConstantGate rows hold two constants each; the gate's constants are per row, so a half-used row is completed in place
        if self._const_row is None:
            r = self.add_gate(ConstantGate(2), constants=[v, 0], wires=[v, 0])
            self._const_row = r
            t = (r, 0)
        else:
            r = self._const_row
            self.rows[r] = (self.rows[r][0], [self.rows[r][1][0], v])
            self.wires[r][1] = v
            self._const_row = None
            t = (r, 1)
        self._consts[v] = t
        return t

    def constant_ext(self, v) -> ExtTarget:
        return (self.constant(v[0]), self.constant(v[1]))

    # ---- base-field arithmetic: output = c0 * m0 * m1 + c1 * addend (ArithmeticGate, 20 operations per row, constants per row)
    def arith(self, m0: Target, m1: Target, addend: Target, c0: int = 1, c1: int = 1) -> Target:
        key = (c0 % P, c1 % P)
        slot = self._arith.get(key)
        if slot is None or slot[1] == self.ar.num_ops:
            slot = (self.add_gate(self.ar, constants=list(key)), 0)
        r, i = slot
        self._arith[key] = (r, i + 1)
        for k, src in enumerate((m0, m1, addend)):
            self._set((r, 4 * i + k), self.val(src), src)
        out = (r, 4 * i + 3)
        self._set(out, self.val(m0) * self.val(m1) % P * key[0] + self.val(addend) * key[1])
        return out

    def mul(self, a, b):
        return self.arith(a, b, self.zero, 1, 0)

    def sub(self, a, b):
        return self.arith(a, self.one, b, 1, P - 1)

    def mul_const(self, a, c):
        return self.arith(a, self.one, self.zero, c, 0)

    def neg(self, a):
        return self.arith(self.zero, self.zero, a, 0, P - 1)

    # ---- extension arithmetic (ArithmeticExtensionGate, 10 operations per row)
    def arith_ext(self, m0: ExtTarget, m1: ExtTarget, addend: ExtTarget, c0: int = 1, c1: int = 1) -> ExtTarget:
        key = (c0 % P, c1 % P)
        slot = self._arith_ext.get(key)
        if slot is None or slot[1] == self.are.num_ops:
            slot = (self.add_gate(self.are, constants=list(key)), 0)
        r, i = slot
        self._arith_ext[key] = (r, i + 1)
        for k, src in enumerate((m0, m1, addend)):
            for comp in range(2):
                self._set((r, 8 * i + 2 * k + comp), self.val(src[comp]), src[comp])
        pr = _e_mod(_e_mul(self.vale(m0), self.vale(m1)))
        ad = self.vale(addend)
        out = ((r, 8 * i + 6), (r, 8 * i + 7))
        self._set(out[0], pr[0] * key[0] + ad[0] * key[1])
        self._set(out[1], pr[1] * key[0] + ad[1] * key[1])
        return out

    @property
    def zero_ext(self) -> ExtTarget:
        return (self.zero, self.zero)

    @property
    def one_ext(self) -> ExtTarget:
        return (self.one, self.zero)

    def mul_ext(self, a, b):
        return self.arith_ext(a, b, self.zero_ext, 1, 0)

    def sub_ext(self, a, b):
        return self.arith_ext(a, self.one_ext, b, 1, P - 1)

    def div_ext(self, num: ExtTarget, den: ExtTarget) -> ExtTarget:
        """q with q * den == num: the quotient is witnessed (advice), one ArithmeticExtension operation constrains it
        (div_add_extension's trick).  The quotient enters as m0 of a fresh operation whose output is connected to num."""
        q = _e_mod(_e_mul(self.vale(num), _e_inv(self.vale(den))))
        key = (1, 0)
        slot = self._arith_ext.get(key)
        if slot is None or slot[1] == self.are.num_ops:
            slot = (self.add_gate(self.are, constants=list(key)), 0)
        r, i = slot
        self._arith_ext[key] = (r, i + 1)
        qt = ((r, 8 * i), (r, 8 * i + 1))
        self._set(qt[0], q[0])
        self._set(qt[1], q[1])
        for comp in range(2):
            self._set((r, 8 * i + 2 + comp), self.val(den[comp]), den[comp])
            self._set((r, 8 * i + 4 + comp), 0, self.zero)
            self._set((r, 8 * i + 6 + comp), self.val(num[comp]), num[comp])
        if _e_mod(_e_mul(q, self.vale(den))) != self.vale(num):  # den == 0 and num != 0
            raise cc.NoWitness("division by zero in a witnessed quotient")
        return qt

    def connect_ext(self, a: ExtTarget, b: ExtTarget):
        self.connect(a[0], b[0])
        self.connect(a[1], b[1])

    # ---- bits (BaseSumGate: wire 0 = sum of limb_i 2^i, limbs boolean)
    def split_bits(self, t: Target, n_bits: int) -> List[Target]:
        r = self.add_gate(self.bs, wires=self.bs.witness(self.val(t)))
        if self.val(t) >= (1 << n_bits):
            raise cc.NoWitness(f"value does not fit {n_bits} bits")
        self.connect((r, 0), t)
        for j in range(n_bits, 63):
            self.connect((r, 1 + j), self.zero)
        return [(r, 1 + j) for j in range(n_bits)]

    def bits_to_target(self, bits: Sequence[Target]) -> Target:
        v = sum(self.val(b) << j for j, b in enumerate(bits))
        r = self.add_gate(self.bs, wires=self.bs.witness(v))
        for j in range(63):
            self.connect((r, 1 + j), bits[j] if j < len(bits) else self.zero)
        return (r, 0)

    # ---- exp_from_bits (ExponentiationGate): base^(sum bits_j 2^j)
    def exp_from_bits(self, base: Target, bits: Sequence[Target]) -> Target:
        power = sum(self.val(b) << j for j, b in enumerate(bits))
        r = self.add_gate(self.exp, wires=self.exp.witness(self.val(base), power))
        self.connect((r, 0), base)
        for j in range(self.exp.n_bits):
            self.connect((r, 1 + j), bits[j] if j < len(bits) else self.zero)
        return (r, 1 + self.exp.n_bits)

    # ---- random access into 16 extension values (RandomAccessGate: copy 0 / 1 = the two components)
    def random_access_ext(self, index: Target, items: Sequence[ExtTarget]) -> ExtTarget:
        ra = self.ra
        assert len(items) == ra.vec
        idx = self.val(index)
        w = [0] * NUM_WIRES
        for c in range(ra.num_copies):
            base = (2 + ra.vec) * c
            comp = c if c < 2 else 0  # copies 2, 3 repeat component 0 (the gate has four copies)
            w[base], w[base + 1] = idx, self.val(items[idx][comp])
            w[base + 2:base + 2 + ra.vec] = [self.val(it[comp]) for it in items]
            for j in range(ra.bits):
                w[ra.num_routed + c * ra.bits + j] = (idx >> j) & 1
        r = self.add_gate(ra, constants=[0, 0], wires=w)
        for c in range(ra.num_copies):
            base = (2 + ra.vec) * c
            comp = c if c < 2 else 0
            self.connect((r, base), index)
            for e in range(ra.vec):
                self.connect((r, base + 2 + e), items[e][comp])
        return ((r, 1), (r, (2 + ra.vec) + 1))

    # ---- reduce_with_powers: sum_k alpha^k coeff_k
    def reduce_base(self, alpha: ExtTarget, coeffs: Sequence[Target]) -> ExtTarget:
        """Base-field coefficients, extension alpha (ReducingGate rows chained through old_acc; Horner from the last coefficient,
        leading slots of the first row padded with zeros)."""
        g = self.red
        seq = list(reversed(coeffs))
        pad = (-len(seq)) % g.num_coeffs
        seq = [self.zero] * pad + seq
        acc_t, acc = self.zero_ext, (0, 0)
        a = self.vale(alpha)
        for off in range(0, len(seq), g.num_coeffs):
            w = [0] * NUM_WIRES
            w[2:6] = [*a, *acc]
            chunk = seq[off:off + g.num_coeffs]
            for i, ct in enumerate(chunk):
                c = self.val(ct)
                w[g.start_coeffs + i] = c
                t = _e_mod(_e_mul(acc, a))
                acc = ((t[0] + c) % P, t[1])
                a0, a1 = g._acc(i)
                w[a0], w[a1] = acc
            r = self.add_gate(g, wires=w)
            self.connect_ext(((r, 2), (r, 3)), alpha)
            self.connect_ext(((r, 4), (r, 5)), acc_t)
            for i, ct in enumerate(chunk):
                self.connect((r, g.start_coeffs + i), ct)
            acc_t = ((r, 0), (r, 1))
        return acc_t

    def reduce_ext(self, alpha: ExtTarget, coeffs: Sequence[ExtTarget]) -> ExtTarget:
        g = self.rede
        seq = list(reversed(coeffs))
        pad = (-len(seq)) % g.num_coeffs
        seq = [self.zero_ext] * pad + seq
        acc_t, acc = self.zero_ext, (0, 0)
        a = self.vale(alpha)
        for off in range(0, len(seq), g.num_coeffs):
            w = [0] * NUM_WIRES
            w[2:6] = [*a, *acc]
            chunk = seq[off:off + g.num_coeffs]
            for i, ct in enumerate(chunk):
                c = self.vale(ct)
                w[g.start_coeffs + 2 * i], w[g.start_coeffs + 2 * i + 1] = c
                t = _e_mod(_e_mul(acc, a))
                acc = ((t[0] + c[0]) % P, (t[1] + c[1]) % P)
                a0, a1 = g._acc(i)
                w[a0], w[a1] = acc
            r = self.add_gate(g, wires=w)
            self.connect_ext(((r, 2), (r, 3)), alpha)
            self.connect_ext(((r, 4), (r, 5)), acc_t)
            for i, ct in enumerate(chunk):
                self.connect_ext(((r, g.start_coeffs + 2 * i), (r, g.start_coeffs + 2 * i + 1)), ct)
            acc_t = ((r, 0), (r, 1))
        return acc_t

    # ---- interpolate_coset (CosetInterpolationGate): the interpolant of 16 values on shift * <w16> at `point`
    def interpolate_coset(self, shift: Target, values: Sequence[ExtTarget], point: ExtTarget) -> ExtTarget:
        g = self.coset
        seq = iter([self.val(shift)] + [x for v in values for x in self.vale(v)] + list(self.vale(point)))
        r = self.add_gate(g, wires=g.witness(lambda: next(seq)))
        self.connect((r, 0), shift)
        for i, v in enumerate(values):
            self.connect_ext(((r, g.start_values + 2 * i), (r, g.start_values + 2 * i + 1)), v)
        self.connect_ext(((r, g.start_point), (r, g.start_point + 1)), point)
        return ((r, g.start_value), (r, g.start_value + 1))


def fri_challenges_and_openings(prover, words: np.ndarray, public_inputs: Sequence[int]) -> dict:
    """Transcript replay of a circuit proof (witness generation for the outer circuit; nothing is checked): FRI alpha, betas,
    zeta, query indices, the reduced openings of the two batches, the caps, the final polynomial and, per query, the opened rows
    and paths.  See circuit.fri_query_openings for the Merkle part."""
    from . import wire
    from .api import Challenger

    p = wire.parse_circuit_proof(words)
    h = p["header"]
    op = p["openings"]
    ch = Challenger()
    ch.observe(prover.digest)
    ch.observe(cc.hash_no_pad(public_inputs))
    ch.observe_cap(p["wires_cap"])
    plonk_betas = [int(x) for x in ch.get_n_challenges(cc.NUM_CHALLENGES)]
    plonk_gammas = [int(x) for x in ch.get_n_challenges(cc.NUM_CHALLENGES)]
    ch.observe_cap(p["plonk_zs_partial_products_cap"])
    plonk_alphas = [int(x) for x in ch.get_n_challenges(cc.NUM_CHALLENGES)]
    ch.observe_cap(p["quotient_polys_cap"])
    zeta = [int(x) for x in ch.get_extension_challenge()]
    order = ("constants", "plonk_sigmas", "wires", "plonk_zs", "partial_products", "quotient_polys")
    for k in order + ("plonk_zs_next",):
        ch.observe(op[k])
    alpha = [int(x) for x in ch.get_extension_challenge()]
    fri = [int(x) for x in p["opening_proof"]]
    capw = 4 << h["cap_height"]
    betas, pos = [], 0
    for _ in range(h["n_fri_layers"]):
        ch.observe(fri[pos:pos + capw])
        pos += capw
        betas.append([int(x) for x in ch.get_extension_challenge()])
    final = fri[len(fri) - 1 - 2 * h["final_poly_len"]:len(fri) - 1]

    def reduce(vals):
        acc = (0, 0)
        for v in reversed(vals):
            acc = _e_mod(_e_mul(acc, alpha))
            acc = ((acc[0] + int(v[0])) % P, (acc[1] + int(v[1])) % P)
        return acc

    batch0 = [v for k in order for v in op[k]]
    return {"header": h, "alpha": alpha, "betas": betas, "zeta": zeta, "reduced": [reduce(batch0), reduce(list(op["plonk_zs_next"]))],
            "plonk_betas": plonk_betas, "plonk_gammas": plonk_gammas, "plonk_alphas": plonk_alphas,
            "batch0": [(int(v[0]), int(v[1])) for v in batch0], "zs_next": [(int(v[0]), int(v[1])) for v in op["plonk_zs_next"]],
            "pi_hash": [int(x) for x in p["public_inputs_hash"]], "pow_witness": fri[-1],
            "final_poly": [(final[2 * i], final[2 * i + 1]) for i in range(h["final_poly_len"])],
            "openings": cc.fri_query_openings(prover, words, public_inputs)}


class _T:
    """The targets one inner proof's verification reads (public-input wires or in-circuit challenger outputs)."""


def _pi_layout(b: GadgetBuilder, d: dict, nq: int, with_challenges: bool, with_openings: bool):
    """The public-input values of one inner proof, and a function that maps the wires back into a _T.
    caps | [alpha | zeta | betas | reduced | indices] | final polynomial | [plonk betas, gammas, alphas] | pi_hash | [openings] | pow."""
    h = d["header"]
    per_query = 4 + h["n_fri_layers"]
    openings = d["openings"][:nq * per_query]
    caps = [tuple(int(x) for dg in openings[o][3] for x in dg) for o in range(per_query)]
    indices = [openings[q * per_query][1] for q in range(nq)]
    vals = [x for cp in caps for x in cp] + [x for cf in d["final_poly"] for x in cf] + d["pi_hash"] + [d["pow_witness"]]
    if with_challenges:
        vals += d["alpha"] + d["zeta"] + [x for be in d["betas"] for x in be] + indices
        if with_openings:
            vals += d["plonk_betas"] + d["plonk_gammas"] + d["plonk_alphas"]
        else:
            vals += [x for r in d["reduced"] for x in r]
    if with_openings:
        vals += [x for v in d["batch0"] for x in v] + [x for v in d["zs_next"] for x in v]

    def assign(pw: Sequence[Target]) -> _T:
        T = _T()
        ext_at = lambda k: (pw[k], pw[k + 1])
        T.caps = [pw[64 * o:64 * (o + 1)] for o in range(per_query)]
        at = 64 * per_query
        T.final = [ext_at(at + 2 * i) for i in range(len(d["final_poly"]))]
        at += 2 * len(d["final_poly"])
        T.pi_hash = pw[at:at + 4]
        T.pow_witness = pw[at + 4]
        at += 5
        if with_challenges:
            T.alpha, T.zeta = ext_at(at), ext_at(at + 2)
            T.betas = [ext_at(at + 4 + 2 * i) for i in range(h["n_fri_layers"])]
            at += 4 + 2 * h["n_fri_layers"]
            T.index_t = pw[at:at + nq]
            at += nq
            if with_openings:
                K = cc.NUM_CHALLENGES
                T.plonk_betas, T.plonk_gammas, T.plonk_alphas = pw[at:at + K], pw[at + K:at + 2 * K], pw[at + 2 * K:at + 3 * K]
                at += 3 * K
            else:
                T.reduced = [ext_at(at), ext_at(at + 2)]
                at += 4
        if with_openings:
            n0 = len(d["batch0"])
            T.batch0 = [ext_at(at + 2 * i) for i in range(n0)]
            T.zs_next = [ext_at(at + 2 * n0 + 2 * i) for i in range(cc.NUM_CHALLENGES)]
        return T

    return vals, assign, per_query, openings


def fri_verifier_circuit(inner: Sequence[tuple], max_queries: int = None, min_degree_bits: int = 0, vanishing: bool = False):
    """verify_fri_proof of the inner circuit proofs `[(CircuitProver, proof words, public inputs)]` as ONE outer circuit (module
    docstring), the challenges given as public inputs: a 2^12-row inner proof costs ~3.8 k outer rows (2^12), two of them 2^13.
    -> (Circuit, wires, public inputs of the outer circuit).  Building fails (AssertionError in connect / div_ext) when an inner
    proof's FRI part is not valid: no witness exists.
    vanishing=True adds the PLONK part of the verifier (plonk/recursive_verifier.rs verify_with_challenges_circuit): the openings
    become public inputs, the reduced openings are computed from them in-circuit, and the inner circuit's vanishing polynomial
    is evaluated at zeta over the extension field — the inner circuit's recorded constraint program re-interpreted with
    ArithmeticExtension operations — and checked against Z_H(zeta) * the reduced quotient chunks.
    recursive_verifier_circuit derives the challenges in-circuit as well."""
    if isinstance(inner, tuple) and not isinstance(inner[0], tuple):
        inner = [inner]
    datas = [fri_challenges_and_openings(*x) for x in inner]
    b = GadgetBuilder()
    layouts, pi_values = [], []
    for d in datas:
        h = d["header"]
        nq = h["num_queries"] if max_queries is None else min(max_queries, h["num_queries"])
        vals, assign, per_query, openings = _pi_layout(b, d, nq, True, vanishing)
        layouts.append((len(pi_values), assign, nq, per_query, openings))
        pi_values += vals
    all_pw = b.merkle.public_inputs(pi_values)
    for (prover, _, _), d, (base, assign, nq, per_query, openings) in zip(inner, datas, layouts):
        T = assign(all_pw[base:])
        log_lde = d["header"]["degree_bits"] + d["header"]["rate_bits"]
        T.index = lambda q, T=T, log_lde=log_lde: (T.index_t[q], b.split_bits(T.index_t[q], log_lde))
        if vanishing:
            T.reduced = [b.reduce_ext(T.alpha, T.batch0), b.reduce_ext(T.alpha, T.zs_next)]
        _fri_part(b, d, T, nq, per_query, openings)
        if vanishing:
            _plonk_part(b, prover.c, d, T)
    circuit, wires = b.build(min_degree_bits)
    return circuit, wires, list(b.public_inputs)


class CircuitChallenger:
    """plonky2 iop/challenger.rs RecursiveChallenger: the duplex sponge of the transcript as PoseidonGate rows (overwrite mode,
    rate 8, challenges popped from the end of the 8-word output buffer) — the in-circuit twin of api.Challenger."""

    def __init__(self, b: GadgetBuilder):
        self.b = b
        self.state = [b.zero] * 12
        self.inp: List[Target] = []
        self.out: List[Target] = []

    def observe(self, targets: Sequence[Target]):
        for t in targets:
            self.out = []
            self.inp.append(t)
            if len(self.inp) == 8:
                self._duplex()

    def observe_ext(self, ts: Sequence[ExtTarget]):
        self.observe([c for t in ts for c in t])

    def _duplex(self):
        b = self.b
        ins = self.inp + self.state[len(self.inp):]
        r = b.add_gate(b.merkle.pos, wires=cc.poseidon_gate_wires([b.val(t) for t in ins], 0))
        b.connect((r, cc.PoseidonGate.WIRE_SWAP), b.zero)
        for k, t in enumerate(ins):
            b.connect((r, k), t)
        self.state = [(r, 12 + k) for k in range(12)]
        self.inp = []
        self.out = self.state[:8]

    def get_challenge(self) -> Target:
        if self.inp or not self.out:
            self._duplex()
        return self.out.pop()

    def get_n(self, n: int) -> List[Target]:
        return [self.get_challenge() for _ in range(n)]

    def compact(self) -> List[Target]:
        """RecursiveChallenger::compact: absorb what is buffered, drop the unused outputs -> the 12 sponge-state targets."""
        if self.inp:
            self._duplex()
        self.out = []
        return list(self.state)

    def get_ext(self) -> ExtTarget:
        a = self.get_challenge()
        return (a, self.get_challenge())


def recursive_verifier_circuit(inner: Sequence[tuple], max_queries: int = None, min_degree_bits: int = 0):
    """The recursive verifier of circuit proofs (plonky2 plonk/recursive_verifier.rs verify_proof): the transcript replayed by an
    in-circuit challenger (get_challenges: circuit digest, public-input hash, caps, openings, FRI caps, final polynomial,
    proof-of-work witness -> betas, gammas, alphas, zeta, FRI alpha, FRI betas, the proof-of-work response, the query indices),
    the proof-of-work check, the vanishing-polynomial check at zeta and the whole FRI verification.  Public inputs of the outer
    circuit: the inner proof's public-input hash, caps, openings, final polynomial and proof-of-work witness; the queried rows
    and paths are advice.  A 2^12-row inner proof gives a 2^13-row outer circuit.  max_queries < 28 builds a partial verifier
    (the transcript still draws all the indices).  Building fails when the inner proof is not valid: no witness exists."""
    if isinstance(inner, tuple) and not isinstance(inner[0], tuple):
        inner = [inner]
    datas = [fri_challenges_and_openings(*x) for x in inner]
    b = GadgetBuilder()
    layouts, pi_values = [], []
    for d in datas:
        h = d["header"]
        nq = h["num_queries"] if max_queries is None else min(max_queries, h["num_queries"])
        vals, assign, per_query, openings = _pi_layout(b, d, nq, False, True)
        layouts.append((len(pi_values), assign, nq, per_query, openings))
        pi_values += vals
    all_pw = b.merkle.public_inputs(pi_values)
    for (prover, _, _), d, (base, assign, nq, per_query, openings) in zip(inner, datas, layouts):
        T = assign(all_pw[base:])
        verify_circuit_proof_in_circuit(b, prover, d, T, nq, per_query, openings)
    circuit, wires = b.build(min_degree_bits)
    return circuit, wires, list(b.public_inputs)


def verify_circuit_proof_in_circuit(b: "GadgetBuilder", prover, d: dict, T: "_T", nq: int, per_query: int, openings: Sequence[tuple]):
    """verify_proof for ONE inner circuit proof on the builder `b`: T holds the targets of the proof's caps, openings, final
    polynomial, public-input hash and proof-of-work witness (public inputs of the outer circuit, or advice when the caller
    publishes something else — stark_circuit.root_circuit); everything else is derived in-circuit."""
    h = d["header"]
    K, log_lde = cc.NUM_CHALLENGES, h["degree_bits"] + h["rate_bits"]
    # ---- get_challenges in-circuit
    ch = CircuitChallenger(b)
    ch.observe([b.constant(x) for x in prover.digest])  # the inner circuit's digest: a constant of the outer circuit
    ch.observe(T.pi_hash)
    ch.observe(T.caps[1])
    T.plonk_betas, T.plonk_gammas = ch.get_n(K), ch.get_n(K)
    ch.observe(T.caps[2])
    T.plonk_alphas = ch.get_n(K)
    ch.observe(T.caps[3])
    T.zeta = ch.get_ext()
    ch.observe_ext(T.batch0)
    ch.observe_ext(T.zs_next)
    T.alpha = ch.get_ext()
    T.betas = []
    for layer in range(h["n_fri_layers"]):
        ch.observe(T.caps[4 + layer])
        T.betas.append(ch.get_ext())
    ch.observe_ext(T.final)
    ch.observe([T.pow_witness])
    pow_response = ch.get_challenge()
    index_challenges = ch.get_n(h["num_queries"])
    for name, want in (("alpha", d["alpha"]), ("zeta", d["zeta"])):  # the in-circuit transcript == the host's
        if list(b.vale(getattr(T, name))) != want:
            raise cc.NoWitness(f"in-circuit transcript diverges from the host's at {name}")
    # ---- proof of work: the response's top pow_bits bits are zero
    lo_bits, hi_bit = _split_64(b, pow_response)
    for t in lo_bits[64 - h["pow_bits"]:] + [hi_bit]:
        b.connect(t, b.zero)
    # ---- query indices: the low log_lde bits of the index challenges
    def index(q, index_challenges=index_challenges, log_lde=log_lde):
        lo, _ = _split_64(b, index_challenges[q])
        bits = lo[:log_lde]
        return b.bits_to_target(bits), bits

    T.index = index
    T.reduced = [b.reduce_ext(T.alpha, T.batch0), b.reduce_ext(T.alpha, T.zs_next)]
    _fri_part(b, d, T, nq, per_query, openings)
    _plonk_part(b, prover.c, d, T)


def _split_64(b: GadgetBuilder, t: Target):
    """split_le(x, 64): two BaseSumGates (63 limbs + 1) recombined as low + 2^63 * high == x.  -> (63 low bits, the top bit).
    (As upstream, the decomposition is not forced to be the canonical one of the two that exist for x < 2^64 - p.)"""
    v = b.val(t)
    lo = b.split_bits(b_advice := _advice(b, v & ((1 << 63) - 1)), 63)
    hi = b.split_bits(_advice(b, v >> 63), 1)
    b.connect(b.arith(b.bits_to_target(hi), b.constant(1 << 63), b_advice, 1, 1), t)
    return lo, hi[0]


def _advice(b: GadgetBuilder, v: int) -> Target:
    """A fresh routed wire holding v (an operand slot of an ArithmeticGate row: v = v * 1 + 0 is constrained, v itself is free)."""
    key = (1, 0)
    slot = b._arith.get(key)
    if slot is None or slot[1] == b.ar.num_ops:
        slot = (b.add_gate(b.ar, constants=list(key)), 0)
    r, i = slot
    b._arith[key] = (r, i + 1)
    b._set((r, 4 * i), v)
    b._set((r, 4 * i + 1), 1, b.one)
    b._set((r, 4 * i + 2), 0, b.zero)
    b._set((r, 4 * i + 3), v)
    return (r, 4 * i)


def _fri_part(b: GadgetBuilder, d: dict, T: _T, nq: int, per_query: int, openings: Sequence[tuple]):
    h = d["header"]
    n_layers, arity_bits = h["n_fri_layers"], h["arity_bits"]
    log_lde = h["degree_bits"] + h["rate_bits"]
    alpha_t, zeta_t, beta_t, reduced_t, final_t = T.alpha, T.zeta, T.betas, T.reduced, T.final
    # ---- per-proof values: zeta_next = g * zeta, alpha^2 (the shift of the first batch past the two openings of the second)
    g = cc.root_of_unity(h["degree_bits"])
    zeta_next_t = (b.mul_const(zeta_t[0], g), b.mul_const(zeta_t[1], g))
    alpha2_t = b.mul_ext(alpha_t, alpha_t)
    w_lde_t = b.constant(cc.root_of_unity(log_lde))
    g16_inv_t = b.constant(pow(cc.root_of_unity(arity_bits), P - 2, P))
    neg_z1 = [b.neg(zeta_t[1]), b.neg(zeta_next_t[1])]
    points = [zeta_t, zeta_next_t]
    for q in range(nq):
        ops = openings[q * per_query:(q + 1) * per_query]
        x_t, bits = T.index(q)
        # fri_verify_initial_proof
        leaf_t = []
        for o in range(4):
            leaf, idx, sib, cap = ops[o]
            it, lt = b.merkle.opening(leaf, idx, sib, cap, T.caps[o])
            b.connect(it, x_t)
            leaf_t.append(lt)
        # subgroup_x = g * w^rev(x_index)
        sx_t = b.mul_const(b.exp_from_bits(w_lde_t, list(reversed(bits))), 7)
        # fri_combine_initial: batch 0 = every polynomial of the four oracles at zeta, batch 1 = the Zs at g * zeta
        evals = [[t for lt in leaf_t for t in lt], leaf_t[2][:cc.NUM_CHALLENGES]]
        quot = []
        for bi in range(2):
            num = b.sub_ext(b.reduce_base(alpha_t, evals[bi]), reduced_t[bi])
            den = (b.sub(sx_t, points[bi][0]), neg_z1[bi])
            quot.append(b.div_ext(num, den))
        old = b.arith_ext(quot[0], alpha2_t, quot[1], 1, 1)
        for layer in range(n_layers):
            leaf, idx, sib, cap = ops[4 + layer]
            lo = arity_bits * layer
            within_bits = bits[lo:lo + arity_bits]
            it, lt = b.merkle.opening(leaf, idx, sib, cap, T.caps[4 + layer])
            b.connect(it, b.bits_to_target(bits[lo + arity_bits:]))
            ev = [(lt[2 * k], lt[2 * k + 1]) for k in range(1 << arity_bits)]
            b.connect_ext(b.random_access_ext(b.bits_to_target(within_bits), ev), old)  # evals[x_index mod 16] == the previous value
            coset_start = b.mul(sx_t, b.exp_from_bits(g16_inv_t, list(reversed(within_bits))))
            nat = [ev[_bitrev(k, arity_bits)] for k in range(1 << arity_bits)]
            old = b.interpolate_coset(coset_start, nat, beta_t[layer])
            for _ in range(arity_bits):
                sx_t = b.mul(sx_t, sx_t)
        # final polynomial at subgroup_x
        b.connect_ext(b.reduce_ext((sx_t, b.zero), final_t), old)


def _plonk_part(b: GadgetBuilder, inner_circuit, d: dict, T: _T):
    """verify_with_challenges_circuit's PLONK check for one inner proof: eval_vanishing_poly at zeta (the inner circuit's
    recorded program over extension targets), the consumer's fold with the alphas, L_0(zeta), and
    vanishing(zeta) == Z_H(zeta) * sum_k zeta^(n k) q_k(zeta) per challenge."""
    h = d["header"]
    K, n_bits = cc.NUM_CHALLENGES, h["degree_bits"]
    lift = lambda t: (t, b.zero)
    zeta_t, batch0_t, zs_next_t = T.zeta, T.batch0, T.zs_next
    n0 = len(batch0_t)
    n_q = K * cc.QUOTIENT_DEGREE_FACTOR
    lv = batch0_t[:n0 - n_q] + [zeta_t]
    nv = [None] * len(lv)
    for i in range(K):
        nv[inner_circuit.col_z(i)] = zs_next_t[i]
    consts = {}

    def lift_const(c):
        c = int(c) % P
        if c not in consts:
            consts[c] = lift(b.constant(c))
        return consts[c]

    # PI / CH operands arrive as targets already; Program.evaluate calls lift() on them as well as on immediates
    ops_lift = lambda x: x if isinstance(x, tuple) else lift_const(x)
    out = inner_circuit.program.evaluate(lv, nv, pi=[lift(t) for t in T.pi_hash], ch=[lift(t) for t in list(T.plonk_betas) + list(T.plonk_gammas)],
                                         add=lambda x, y: b.arith_ext(x, b.one_ext, y, 1, 1), sub=lambda x, y: b.sub_ext(x, y),
                                         mul=lambda x, y: b.mul_ext(x, y), lift=ops_lift)
    # L_0(zeta) = Z_H(zeta) / (n (zeta - 1)), Z_H(zeta) = zeta^n - 1
    zeta_n = zeta_t
    for _ in range(n_bits):
        zeta_n = b.mul_ext(zeta_n, zeta_n)
    z_h = b.sub_ext(zeta_n, b.one_ext)
    l_0 = b.div_ext(z_h, b.mul_ext(lift_const(1 << n_bits), b.sub_ext(zeta_t, b.one_ext)))
    quot_t = batch0_t[n0 - n_q:]
    for j in range(K):
        a = lift(T.plonk_alphas[j])
        acc = b.zero_ext
        for kind, c in out:  # the consumer's fold: acc <- acc * alpha + c * multiplier
            if kind == cc.cprog.EMIT_FIRST_ROW:
                c = b.mul_ext(c, l_0)
            acc = b.arith_ext(acc, a, c, 1, 1)
        f = cc.QUOTIENT_DEGREE_FACTOR
        reduced_q = b.reduce_ext(zeta_n, quot_t[j * f:(j + 1) * f])
        b.connect_ext(b.mul_ext(z_h, reduced_q), acc)  # vanishing(zeta) == Z_H(zeta) * sum_k zeta^(n k) q_k(zeta)
