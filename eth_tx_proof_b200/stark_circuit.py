"""The STARK verifier as a circuit: the FIRST recursion layer of the reference (one wrapper circuit per table, then the root
circuit that ties the tables together).  Upstream: starky 0.4.0 src/recursive_verifier.rs `verify_stark_proof_circuit` /
`verify_stark_proof_with_challenges_circuit`, src/get_challenges.rs `get_challenges_circuit`, src/cross_table_lookup.rs
`verify_cross_table_lookups_circuit`; evm_arithmetization 0.1.3 src/fixed_recursive_verifier.rs `recursive_stark_circuit` and
`create_root_circuit` (crates pinned at /root/reference/Cargo.lock:4529,1675, not on disk; reached from
/root/reference/ops/src/lib.rs:52 through proof_gen::generate_txn_proof, and the shape every later layer —
/root/reference/ops/src/lib.rs:72,95 — aggregates).

`stark_wrapper_circuit` verifies ONE table proof of this library ("B200STK2" words, from etp_stark_prove_* or
etp_prove_with_commitment) completely in-circuit:

    transcript        in-circuit duplex-sponge challenger (fri_circuit.CircuitChallenger), started either fresh (stand-alone
                      proofs: public inputs, trace cap) or from `init_challenger_state` (multi-table proofs: the state
                      prove_single_table compacted before this table); lookup challenges, alphas, zeta, FRI alpha and betas,
                      proof-of-work response, query indices; the state compacted after the proof is exposed
    vanishing check   the table's recorded constraint program (own constraints, lookup checks with filters, CTL checks)
                      re-interpreted over extension targets at zeta; ConstraintConsumer multipliers z_last, L_first, L_last;
                      vanishing(zeta) == Z_H(zeta) * sum_k zeta^(n k) q_k(zeta) per challenge
    FRI               starky's instance — oracles trace / auxiliary / quotient, batches at zeta, g zeta and (CTL tables) 1 for
                      the Z columns — through Merkle openings, fri_combine_initial, arity-16 folds and the final polynomial

Public inputs of a wrapper circuit, in upstream's spirit (trace cap, CTL data, challenger states; the proof body is advice):

    trace_cap (64) | the table's public inputs | ctl_zs_first | ctl challenges (beta, gamma) x 2 | init_challenger_state (12)
    | challenger state after the proof (12)                  (the last three only for multi-table proofs)

`root_circuit` verifies the wrapper proofs of all tables of a transaction (circuit proofs: fri_circuit's recursive verifier)
and adds what create_root_circuit adds: the wrapper public inputs are re-hashed in-circuit and tied to each proof's
public-input hash; ONE challenger observes every trace cap, draws the CTL challenges (equal to every table's) and hands its
compacted state to the first table; every table's final state is the next table's initial state; and
verify_cross_table_lookups: per CTL and challenge, the looking tables' Z(1) sum to the looked table's.

Everything here is host logic (witness generation and circuit construction in one pass, like fri_circuit.py); the proofs are
made by the circuit prover on the device (`circuit.CircuitProver`) or, in CPU tests, by the oracle.
"""
from __future__ import annotations

from typing import List, Sequence

from . import circuit as cc
from . import cprog, wire
from .circuit import NoopGate, P
from .fri_circuit import (CircuitChallenger, ExtTarget, GadgetBuilder, Target, _bitrev, _pi_layout, _split_64, fri_challenges_and_openings,
                          verify_circuit_proof_in_circuit)

NUM_CHALLENGES = 2  # StarkConfig::standard_fast_config()


def advice(b: GadgetBuilder, values: Sequence[int]) -> List[Target]:
    """Fresh routed wires holding `values` (NoopGate rows: 80 unconstrained routed wires each) — the proof body of a recursive
    verifier is advice, only what the next layer needs is a public input."""
    out = []
    vals = [int(v) % P for v in values]
    for off in range(0, len(vals), cc.NUM_ROUTED):
        chunk = vals[off:off + cc.NUM_ROUTED]
        r = b.add_gate(NoopGate(), wires=chunk)
        out += [(r, k) for k in range(len(chunk))]
    return out


def _advice_ext(b, values) -> List[ExtTarget]:
    flat = advice(b, [c for v in values for c in v])
    return [(flat[2 * i], flat[2 * i + 1]) for i in range(len(values))]


def pow_ext(b: GadgetBuilder, x: ExtTarget, n: int) -> ExtTarget:
    """x^n by square and multiply (n is a constant of the circuit: a batch length)."""
    assert n >= 1
    acc, sq = None, x
    while n:
        if n & 1:
            acc = sq if acc is None else b.mul_ext(acc, sq)
        n >>= 1
        if n:
            sq = b.mul_ext(sq, sq)
    return acc


class WrapperTargets:
    """What a wrapper circuit exposes (targets inside the builder; `values()` are the public inputs in order)."""

    def __init__(self):
        self.trace_cap: List[Target] = []
        self.public_inputs: List[Target] = []
        self.ctl_zs_first: List[Target] = []
        self.ctl_challenges: List[Target] = []
        self.state_in: List[Target] = []
        self.state_out: List[Target] = []

    def flat(self) -> List[Target]:
        return self.trace_cap + self.public_inputs + self.ctl_zs_first + self.ctl_challenges + self.state_in + self.state_out


def wrapper_public_input_layout(program: cprog.Program, multi_table: bool, n_public_inputs: int = None) -> dict:
    """Offsets of the fields inside a wrapper circuit's public inputs (module docstring)."""
    n_pi = program.n_pi if n_public_inputs is None else n_public_inputs
    z = len(program.ctl_zs)
    at = {"trace_cap": (0, 64), "public_inputs": (64, n_pi), "ctl_zs_first": (64 + n_pi, z)}
    pos = 64 + n_pi + z
    if multi_table:
        at["ctl_challenges"] = (pos, 2 * NUM_CHALLENGES)
        at["state_in"] = (pos + 2 * NUM_CHALLENGES, 12)
        at["state_out"] = (pos + 2 * NUM_CHALLENGES + 12, 12)
        pos += 2 * NUM_CHALLENGES + 24
    at["total"] = pos
    return at


def verify_stark_proof_in_circuit(b: GadgetBuilder, program: cprog.Program, words, init_challenger_state=None, ctl_challenges=None,
                                  max_queries: int = None) -> WrapperTargets:
    """verify_stark_proof_circuit on the builder `b` for one table proof.  Multi-table proofs (prove_with_commitment on a
    shared transcript) pass `init_challenger_state` (the 12 words of challenger.compact() before the table) and
    `ctl_challenges` (beta0, gamma0, beta1, gamma1).  Fails while building (AssertionError of a copy constraint / a witnessed
    division) when the proof is not valid: such a circuit has no witness.  -> the targets a caller publishes or links."""
    multi = init_challenger_state is not None
    pp = wire.parse(words)
    h, pr = pp["header"], pp["proof"]
    op = pr["openings"]
    K, db, rate_bits = h["num_challenges"], h["degree_bits"], h["rate_bits"]
    n_trace, n_aux, n_quot, n_z = h["n_trace"], h["n_aux"], h["n_quot"], h["n_ctl_zs"]
    factor = max(1, program.degree - 1)
    if K != NUM_CHALLENGES or h["cap_height"] != 4:
        raise ValueError("the wrapper circuit is built for StarkConfig::standard_fast_config()")
    if (n_trace, n_aux, n_quot) != (program.n_trace, program.n_aux, factor * K) or n_z != len(program.ctl_zs):
        raise ValueError("the proof is not of this table")
    if n_z and not multi:
        raise ValueError("a proof with CTL openings needs init_challenger_state and the CTL challenges")
    lde_bits = db + rate_bits
    flat_cap = lambda cap: [x for d in cap for x in d["elements"]]
    cap_rows = lambda cap: [[int(x) for x in d["elements"]] for d in cap]
    exts = lambda vs: [(int(v[0]), int(v[1])) for v in (vs or [])]

    # ---- the proof as advice
    W = WrapperTargets()
    W.trace_cap = advice(b, flat_cap(pr["trace_cap"]))
    W.public_inputs = advice(b, pp["public_inputs"])
    aux_cap_t = advice(b, flat_cap(pr["auxiliary_polys_cap"])) if n_aux else None
    quot_cap_t = advice(b, flat_cap(pr["quotient_polys_cap"]))
    local_t, next_t = _advice_ext(b, exts(op["local_values"])), _advice_ext(b, exts(op["next_values"]))
    aux_t, aux_next_t = _advice_ext(b, exts(op["auxiliary_polys"])), _advice_ext(b, exts(op["auxiliary_polys_next"]))
    W.ctl_zs_first = advice(b, op["ctl_zs_first"] or [])
    quot_t = _advice_ext(b, exts(op["quotient_polys"]))
    fp = pr["opening_proof"]
    layer_caps_t = [advice(b, flat_cap(c)) for c in fp["commit_phase_merkle_caps"]]
    final_t = _advice_ext(b, exts(fp["final_poly"]["coeffs"]))
    pow_t = advice(b, [fp["pow_witness"]])[0]
    if multi:
        W.state_in = advice(b, init_challenger_state)
        W.ctl_challenges = advice(b, ctl_challenges)
        if len(W.state_in) != 12 or len(W.ctl_challenges) != 2 * K:
            raise ValueError("init_challenger_state has 12 words, ctl_challenges 2 per challenge")

    # ---- get_challenges_circuit
    ch = CircuitChallenger(b)
    if multi:
        ch.state = list(W.state_in)  # the trace cap was observed, with every other table's, before the CTL challenges were drawn
    else:
        ch.observe(W.public_inputs)
        ch.observe(W.trace_cap)
    lookup_ch = None
    if n_aux:
        if multi:
            lookup_ch = [W.ctl_challenges[2 * i] for i in range(K)]  # the CTL betas double as the lookup challenges
        else:
            lookup_ch = ch.get_n(2 * K)[0::2]  # get_grand_product_challenge_set: beta used, gamma drawn and dropped
        ch.observe(aux_cap_t)
    scalars = list(lookup_ch or [b.zero] * K) + list(W.ctl_challenges)
    alphas = ch.get_n(K)
    ch.observe(quot_cap_t)
    zeta = ch.get_ext()
    zeta_batch = local_t + aux_t + quot_t
    next_batch = next_t + aux_next_t
    ctl_batch = [(t, b.zero) for t in W.ctl_zs_first]
    ch.observe_ext(zeta_batch)
    ch.observe_ext(next_batch)
    ch.observe_ext(ctl_batch)
    fri_alpha = ch.get_ext()
    betas = []
    for cap_t in layer_caps_t:
        ch.observe(cap_t)
        betas.append(ch.get_ext())
    ch.observe_ext(final_t)
    ch.observe([pow_t])
    pow_response = ch.get_challenge()
    index_challenges = ch.get_n(h["num_queries"])
    W.state_out = ch.compact() if multi else []

    # ---- proof of work, query indices
    lo_bits, hi_bit = _split_64(b, pow_response)
    if h["pow_bits"]:
        for t in lo_bits[64 - h["pow_bits"]:] + [hi_bit]:
            b.connect(t, b.zero)

    def index(q):
        lo, _ = _split_64(b, index_challenges[q])
        bits = lo[:lde_bits]
        return b.bits_to_target(bits), bits

    # ---- verify_stark_proof_with_challenges_circuit: the constraints at zeta
    lift = lambda t: (t, b.zero)
    consts = {}

    def lift_const(c):
        c = int(c) % P
        if c not in consts:
            consts[c] = lift(b.constant(c))
        return consts[c]

    g = cc.root_of_unity(db)
    zeta_pow = zeta
    for _ in range(db):
        zeta_pow = b.mul_ext(zeta_pow, zeta_pow)
    z_x = b.sub_ext(zeta_pow, b.one_ext)
    n_e = lift_const(1 << db)
    l_first = b.div_ext(z_x, b.mul_ext(n_e, b.sub_ext(zeta, b.one_ext)))
    g_zeta = (b.mul_const(zeta[0], g), b.mul_const(zeta[1], g))
    l_last = b.div_ext(z_x, b.mul_ext(n_e, b.sub_ext(g_zeta, b.one_ext)))
    z_last = b.sub_ext(zeta, lift_const(pow(g, P - 2, P)))
    mult = {cprog.EMIT: None, cprog.EMIT_TRANSITION: z_last, cprog.EMIT_FIRST_ROW: l_first, cprog.EMIT_LAST_ROW: l_last}
    ops_lift = lambda x: x if isinstance(x, tuple) else lift_const(x)
    out = program.evaluate(local_t, next_t, aux_t, aux_next_t, pi=[lift(t) for t in W.public_inputs], ch=[lift(t) for t in scalars],
                           add=lambda x, y: b.arith_ext(x, b.one_ext, y, 1, 1), sub=lambda x, y: b.sub_ext(x, y),
                           mul=lambda x, y: b.mul_ext(x, y), lift=ops_lift)
    terms = [c if mult[kind] is None else b.mul_ext(c, mult[kind]) for kind, c in out]
    for j in range(K):
        a = lift(alphas[j])
        acc = b.zero_ext
        for c in terms:  # ConstraintConsumer: acc <- acc * alpha + c
            acc = b.arith_ext(acc, a, c, 1, 1)
        reduced_q = b.reduce_ext(zeta_pow, quot_t[j * factor:(j + 1) * factor])
        b.connect_ext(b.mul_ext(z_x, reduced_q), acc)  # "Mismatch between evaluation and opening of quotient polynomial"

    # ---- verify_fri_proof_circuit over starky's instance
    oracle_caps_t = [W.trace_cap] + ([aux_cap_t] if n_aux else []) + [quot_cap_t]
    oracle_caps = [cap_rows(pr["trace_cap"])] + ([cap_rows(pr["auxiliary_polys_cap"])] if n_aux else []) + [cap_rows(pr["quotient_polys_cap"])]
    layer_caps = [cap_rows(c) for c in fp["commit_phase_merkle_caps"]]
    z_first_col = h["n_lookup_cols"] + h["n_ctl_helper_cols"]
    n_or = len(oracle_caps)

    def batches(leaf_t):
        trace_ev, aux_ev, quot_ev = leaf_t[0], (leaf_t[1] if n_aux else []), leaf_t[-1]
        bs = [(trace_ev + aux_ev + quot_ev, zeta, red[0]), (trace_ev + aux_ev, g_zeta, red[1])]
        if n_z:
            bs.append((aux_ev[z_first_col:], b.one_ext, red[2]))
        return bs

    red = [b.reduce_ext(fri_alpha, zeta_batch), b.reduce_ext(fri_alpha, next_batch)] + ([b.reduce_ext(fri_alpha, ctl_batch)] if n_z else [])
    shifts = [len(zeta_batch), len(next_batch)] + ([n_z] if n_z else [])
    shift_t = [None] + [pow_ext(b, fri_alpha, s) for s in shifts[1:]]
    nq = h["num_queries"] if max_queries is None else min(max_queries, h["num_queries"])
    n_layers, arity_bits = h["n_fri_layers"], h["arity_bits"]
    w_lde_t = b.constant(cc.root_of_unity(lde_bits))
    g_ar_inv_t = b.constant(pow(cc.root_of_unity(arity_bits), P - 2, P))
    quads = lambda sibs: [[int(x) for x in s["elements"]] for s in sibs]
    for q in range(nq):
        rp = fp["query_round_proofs"][q]
        x_t, bits = index(q)
        x_index = b.val(x_t)
        # fri_verify_initial_proof
        leaf_t = []
        for o in range(n_or):
            leaf, mp = rp["initial_trees_proof"]["evals_proofs"][o]
            it, lt = b.merkle.opening(leaf, x_index, quads(mp["siblings"]), oracle_caps[o], oracle_caps_t[o])
            b.connect(it, x_t)
            leaf_t.append(lt)
        sx_t = b.mul_const(b.exp_from_bits(w_lde_t, list(reversed(bits))), 7)  # subgroup_x = g * w^rev(x_index)
        # fri_combine_initial: sum <- sum * alpha^len(batch) + (reduce(alpha, evals) - reduced opening) / (x - point)
        old = None
        for bi, (evals, point, reduced) in enumerate(batches(leaf_t)):
            num = b.sub_ext(b.reduce_base(fri_alpha, evals), reduced)
            den = (b.sub(sx_t, point[0]), b.neg(point[1]))
            quo = b.div_ext(num, den)
            old = quo if old is None else b.arith_ext(old, shift_t[bi], quo, 1, 1)
        idx = x_index
        for layer in range(n_layers):
            step = rp["steps"][layer]
            lo = arity_bits * layer
            within_bits = bits[lo:lo + arity_bits]
            idx >>= arity_bits
            leaf = [c for e in step["evals"] for c in e]
            it, lt = b.merkle.opening(leaf, idx, quads(step["merkle_proof"]["siblings"]), layer_caps[layer], layer_caps_t[layer])
            b.connect(it, b.bits_to_target(bits[lo + arity_bits:]))
            ev = [(lt[2 * k], lt[2 * k + 1]) for k in range(1 << arity_bits)]
            b.connect_ext(b.random_access_ext(b.bits_to_target(within_bits), ev), old)  # evals[x_index mod 16] == the previous value
            coset_start = b.mul(sx_t, b.exp_from_bits(g_ar_inv_t, list(reversed(within_bits))))
            nat = [ev[_bitrev(k, arity_bits)] for k in range(1 << arity_bits)]
            old = b.interpolate_coset(coset_start, nat, betas[layer])
            for _ in range(arity_bits):
                sx_t = b.mul(sx_t, sx_t)
        b.connect_ext(b.reduce_ext((sx_t, b.zero), final_t), old)  # "Final polynomial evaluation is invalid."
    return W


def stark_wrapper_circuit(program: cprog.Program, words, init_challenger_state=None, ctl_challenges=None, max_queries: int = None,
                          min_degree_bits: int = 0):
    """recursive_stark_circuit: ONE outer circuit that verifies a table proof (module docstring).
    -> (Circuit, wires, public inputs).  The structure (gates, constants, sigmas) depends on the table and the degree only."""
    b = GadgetBuilder()
    W = verify_stark_proof_in_circuit(b, program, words, init_challenger_state, ctl_challenges, max_queries)
    _publish(b, W.flat())
    circuit, wires = b.build(min_degree_bits)
    return circuit, wires, list(b.public_inputs)


def verify_cross_table_lookups_circuit(b: GadgetBuilder, ctls: Sequence[tuple], ctl_zs_first: Sequence[Sequence[Target]], num_challenges: int = NUM_CHALLENGES):
    """cross_table_lookup.rs verify_cross_table_lookups_circuit: for every CTL and challenge the looking tables' Z(1) sum to the
    looked table's.  ctls: [(looking table indices (with repeats), looked table index)]; ctl_zs_first: per table, targets."""
    its = [iter(v) for v in ctl_zs_first]
    for looking, looked in ctls:
        uniq = []
        for t in looking:
            if t not in uniq:
                uniq.append(t)
        for _ in range(num_challenges):
            total = b.zero
            for t in uniq:
                total = b.arith(next(its[t]), b.one, total, 1, 1)
            b.connect(total, next(its[looked]))
    for it in its:
        if next(it, None) is not None:
            raise ValueError("unused ctl_zs_first openings")


def verify_circuit_proof_with_public_inputs(b: GadgetBuilder, prover, words, inner_pis: Sequence[int], max_queries: int = None) -> List[Target]:
    """verify_proof of ONE inner circuit proof with the proof body as advice and the inner circuit's constants / sigmas cap as
    constants of the outer circuit (verifier data); the inner public inputs become advice targets, re-hashed in-circuit and tied
    to the proof's public-input hash.  -> the targets of the inner public inputs (for the caller to publish or to link)."""
    d = fri_challenges_and_openings(prover, words, inner_pis)
    nq = d["header"]["num_queries"] if max_queries is None else min(max_queries, d["header"]["num_queries"])
    vals, assign, per_query, openings = _pi_layout(b, d, nq, False, True)
    T = assign(advice(b, vals))
    for t in T.caps[0]:
        b.connect(t, b.constant(b.val(t)))
    verify_circuit_proof_in_circuit(b, prover, d, T, nq, per_query, openings)
    pis_t = advice(b, inner_pis)
    r_h, entered = b.merkle.sponge([b.val(t) for t in pis_t])
    for e, t in zip(entered, pis_t):
        b.connect(e, t)
    for i in range(4):
        b.connect((r_h, 12 + i), T.pi_hash[i])
    return pis_t


def _publish(b: GadgetBuilder, targets: Sequence[Target]):
    pw = b.merkle.public_inputs([b.val(t) for t in targets])
    for w_, t in zip(pw, targets):
        b.connect(w_, t)


def shrink_circuit(inner: tuple, max_queries: int = None, min_degree_bits: int = 0):
    """One shrinking step (fixed_recursive_verifier.rs shrinking_config wrappers: `add_virtual_proof_with_pis`, `verify_proof`,
    `register_public_inputs(&proof_with_pis.public_inputs)`): verifies `(CircuitProver, proof words, public inputs)` and has the
    SAME public inputs, so the chain wrapper -> shrink -> ... -> root keeps exposing trace cap, CTL openings and challenger
    states.  Whatever the inner size, the result is a 2^13-row circuit (the recursion's fixed point)."""
    b = GadgetBuilder()
    _publish(b, verify_circuit_proof_with_public_inputs(b, inner[0], inner[1], inner[2], max_queries))
    circuit, wires = b.build(min_degree_bits)
    return circuit, wires, list(b.public_inputs)


def root_circuit(wrappers: Sequence[tuple], layouts: Sequence[dict], ctls: Sequence[tuple], public_values: Sequence[int] = (),
                 max_queries: int = None, min_degree_bits: int = 0):
    """create_root_circuit: verifies the (wrapped and possibly shrunk) proofs `[(CircuitProver, proof words, public inputs)]` of
    ALL tables of one transaction (in table order) and links them (module docstring).  layouts[t] =
    wrapper_public_input_layout of table t (multi-table form).  Public inputs of the root: every trace cap ++ public_values ++
    the CTL challenges.  -> (Circuit, wires, public inputs)."""
    b = GadgetBuilder()
    Ws = []
    for (prover, words, inner_pis), lay in zip(wrappers, layouts):
        if len(inner_pis) != lay["total"]:
            raise ValueError("public inputs do not match the wrapper layout of the table")
        pis_t = verify_circuit_proof_with_public_inputs(b, prover, words, inner_pis, max_queries)
        Ws.append({k: pis_t[v[0]:v[0] + v[1]] for k, v in lay.items() if k != "total"})
    # one transcript over all tables: trace caps, public values -> CTL challenges -> the first table's initial state
    pv_t = advice(b, public_values)
    ch = CircuitChallenger(b)
    for W in Ws:
        ch.observe(W["trace_cap"])
    ch.observe(pv_t)
    ctl_ch = ch.get_n(2 * NUM_CHALLENGES)
    state = ch.compact()
    for W in Ws:
        for a, c in zip(W["ctl_challenges"], ctl_ch):
            b.connect(a, c)
        for a, c in zip(W["state_in"], state):
            b.connect(a, c)
        state = W["state_out"]
    verify_cross_table_lookups_circuit(b, ctls, [W["ctl_zs_first"] for W in Ws])
    _publish(b, [t for W in Ws for t in W["trace_cap"]] + pv_t + ctl_ch)
    circuit, wires = b.build(min_degree_bits)
    return circuit, wires, list(b.public_inputs)


def transaction_recursion_plan(tables: Sequence[tuple], ctls: Sequence[tuple], all_proof, circuit_prove, threshold_degree_bits: int = 13,
                               max_queries: int = None, log=None, public_values: Sequence[int] = ()) -> List[dict]:
    """The recursion layers of ONE transaction, built over its real table proofs in the order the reference proves them
    (proof_gen::generate_txn_proof -> AllRecursiveCircuits::prove_root, /root/reference/ops/src/lib.rs:52): per table the
    wrapper circuit of its STARK proof, then shrinking steps until the proof's circuit has at most 2^threshold_degree_bits rows
    (at least one, as upstream's shrinking_config chain), then the root circuit over the seven shrunk proofs.
    tables: [(name, Program, trace)]; all_proof: prover.AllProof-shaped (stark_proofs, init_challenger_states,
    ctl_challenges); circuit_prove(circuit, wires, public_inputs) -> (prover-like with .c/.digest/.constants_sigmas_cap, proof
    words).  -> [{"name", "kind", "circuit", "wires", "public_inputs", "prover", "words"}] in proving order: every later circuit
    verifies the proof(s) produced before it, so proving the list again with the same witnesses reproduces the whole job."""
    plan, tops, layouts = [], [], []

    def step(name, kind, built):
        circuit, wires, pis = built
        prover, words = circuit_prove(circuit, wires, pis)
        plan.append({"name": name, "kind": kind, "circuit": circuit, "wires": wires, "public_inputs": pis, "prover": prover, "words": words})
        if log:
            log(f"{name}: {kind} circuit 2^{circuit.degree_bits} rows")
        return prover, words, pis

    for (name, prog, _), words, state in zip(tables, all_proof.stark_proofs, all_proof.init_challenger_states):
        cur = step(name, "wrapper", stark_wrapper_circuit(prog, words, state, all_proof.ctl_challenges, max_queries))
        first = True
        while first or cur[0].c.degree_bits > threshold_degree_bits:
            cur = step(name, "shrink", shrink_circuit(cur, max_queries))
            first = False
        tops.append(cur)
        layouts.append(wrapper_public_input_layout(prog, True))
    step("root", "root", root_circuit(tops, layouts, ctls, public_values=public_values, max_queries=max_queries))
    return plan


# ---- aggregation and block layers (/root/reference/ops/src/lib.rs:72 AggProof::combine -> generate_agg_proof, :95 BlockProof ->
# generate_block_proof; evm_arithmetization fixed_recursive_verifier.rs create_aggregation_circuit / create_block_circuit).
# Upstream's public values chain state: a child's "after" must be the next child's "before".  Here a transaction's
# public_values are [before (STATE_WORDS) ++ after (STATE_WORDS)] (root public inputs: trace caps ++ public_values ++ CTL
# challenges), an aggregation proof publishes [before ++ after] of its span, a block proof [before ++ after ++ block number].
# Upstream's circuits are cyclic (one aggregation circuit verifies a root OR an aggregation proof; the block circuit verifies its
# own previous proof); without cyclic recursion every level is its own circuit here — same checks, per-level verifier data.
STATE_WORDS = 4


def root_io(n_tables: int) -> tuple:
    """(offset of `before`, offset of `after`) inside a root proof's public inputs."""
    return 64 * n_tables, 64 * n_tables + STATE_WORDS


def aggregation_circuit(lhs: tuple, rhs: tuple, lhs_io: tuple = (0, STATE_WORDS), rhs_io: tuple = (0, STATE_WORDS), max_queries: int = None,
                        min_degree_bits: int = 0):
    """create_aggregation_circuit: verifies two child proofs `(CircuitProver, words, public inputs)` — roots or aggregations,
    `*_io` = (offset of before, offset of after) in the child's public inputs — and checks lhs.after == rhs.before.
    Public inputs: lhs.before ++ rhs.after."""
    b = GadgetBuilder()
    lt = verify_circuit_proof_with_public_inputs(b, *lhs, max_queries)
    rt = verify_circuit_proof_with_public_inputs(b, *rhs, max_queries)
    for k in range(STATE_WORDS):
        b.connect(lt[lhs_io[1] + k], rt[rhs_io[0] + k])  # "the state after the left child is the state before the right child"
    _publish(b, lt[lhs_io[0]:lhs_io[0] + STATE_WORDS] + rt[rhs_io[1]:rhs_io[1] + STATE_WORDS])
    circuit, wires = b.build(min_degree_bits)
    return circuit, wires, list(b.public_inputs)


def block_circuit(agg: tuple, prev_block: tuple = None, agg_io: tuple = (0, STATE_WORDS), genesis_number: int = 0, max_queries: int = None,
                  min_degree_bits: int = 0):
    """create_block_circuit: verifies the block's aggregation proof and, when there is one, the previous block proof (its `after`
    must be this block's `before`, the block number increases by one).  Public inputs: before ++ after ++ block number."""
    b = GadgetBuilder()
    at = verify_circuit_proof_with_public_inputs(b, *agg, max_queries)
    before, after = at[agg_io[0]:agg_io[0] + STATE_WORDS], at[agg_io[1]:agg_io[1] + STATE_WORDS]
    if prev_block is None:
        number = b.constant(genesis_number)
    else:
        pt = verify_circuit_proof_with_public_inputs(b, *prev_block, max_queries)
        for k in range(STATE_WORDS):
            b.connect(pt[STATE_WORDS + k], before[k])
        number = b.arith(pt[2 * STATE_WORDS], b.one, b.one, 1, 1)
        before = pt[:STATE_WORDS]  # the chain's public `before` stays the first block's (upstream: checkpoint state root)
    _publish(b, before + after + [number])
    circuit, wires = b.build(min_degree_bits)
    return circuit, wires, list(b.public_inputs)


def block_recursion_plan(root_step: dict, n_tables: int, circuit_prove, levels: int = 3, max_queries: int = None, log=None) -> List[dict]:
    """The layers above 2^levels identical, chainable transactions (public values before == after): per level ONE aggregation
    circuit over two proofs of the level below (AggProof::combine folds a binary tree, /root/reference/ops/src/lib.rs:64-76),
    then the block circuit over the top aggregation proof (:95).  root_step: the last step of transaction_recursion_plan.
    -> [{"name": "agg<level>" | "block", ...}] in proving order (one proof per circuit: the tree's 2^(levels - l) proofs of a
    level re-prove the same circuit with the same witness)."""
    plan = []
    cur = (root_step["prover"], root_step["words"], root_step["public_inputs"])
    io = root_io(n_tables)
    for level in range(1, levels + 1):
        circuit, wires, pis = aggregation_circuit(cur, cur, io, io, max_queries)
        prover, words = circuit_prove(circuit, wires, pis)
        plan.append({"name": f"agg{level}", "kind": "aggregation", "circuit": circuit, "wires": wires, "public_inputs": pis, "prover": prover, "words": words})
        if log:
            log(f"agg{level}: aggregation circuit 2^{circuit.degree_bits} rows")
        cur, io = (prover, words, pis), (0, STATE_WORDS)
    circuit, wires, pis = block_circuit(cur, max_queries=max_queries)
    prover, words = circuit_prove(circuit, wires, pis)
    plan.append({"name": "block", "kind": "block", "circuit": circuit, "wires": wires, "public_inputs": pis, "prover": prover, "words": words})
    if log:
        log(f"block: block circuit 2^{circuit.degree_bits} rows")
    return plan
