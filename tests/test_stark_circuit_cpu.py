"""CPU tests of the first recursion layer (eth_tx_proof_b200/stark_circuit.py; starky 0.4.0 src/recursive_verifier.rs,
src/get_challenges.rs, src/cross_table_lookup.rs verify_cross_table_lookups_circuit; evm_arithmetization 0.1.3
src/fixed_recursive_verifier.rs recursive_stark_circuit / create_root_circuit — /root/reference/Cargo.lock:4529,1675, reached
from /root/reference/ops/src/lib.rs:52): table proofs made by the oracle are verified IN-CIRCUIT, the wrapper circuits'
constraints hold, the oracle's circuit prover proves them and the independent Python verifier accepts; the root circuit links
the tables of one transaction (challenger chain, CTL challenges, cross-table lookup sums).  A builder's copy constraint asserts
equality of the connected values, so an invalid inner proof has no witness: building raises."""
import types

import numpy as np
import pytest

import oracle
import plonk_verifier
import stark_verifier as V
from eth_tx_proof_b200 import circuit as cc
from eth_tx_proof_b200 import cprog
from eth_tx_proof_b200 import stark_circuit as sc
from eth_tx_proof_b200 import synthetic as syn
from test_circuit_cpu import _violations, _words_from_oracle_proof

P = cprog.P
DIGEST = [9, 8, 7, 6]


def _constraints_hold(circuit, wires, pis):
    zs_pp = oracle.plonk_partial_products_and_zs(wires[:80], circuit.sigmas, circuit.k_is, 8, [3, 5], [7, 11])
    return _violations(circuit, wires, zs_pp, cc.hash_no_pad(pis), [3, 5], [7, 11]) == []


def _prove_and_verify(circuit, wires, pis):
    proof = oracle.circuit_prove(circuit, wires, pis, DIGEST)
    plonk_verifier.verify(proof, circuit, proof["constants_sigmas_cap"], DIGEST, max_queries=1)
    fake = types.SimpleNamespace(c=circuit, digest=DIGEST, constants_sigmas_cap=proof["constants_sigmas_cap"])
    return fake, _words_from_oracle_proof(circuit, proof, pis), pis


def test_fibonacci_proof_verified_in_circuit():
    """A stand-alone proof (own transcript: public inputs, trace cap): 6 rows of Fibonacci at 2^6, no FRI layer; and 2^9 with one
    arity-16 layer.  Public inputs of the wrapper: trace cap ++ the table's public inputs."""
    for log_n, nq in ((6, 2), (9, 1)):
        t, pi = syn.fibonacci_trace(log_n, seed=log_n)
        proof = oracle.stark_prove(oracle.TABLE_FIBONACCI, t, pi)
        pr = V.verify(proof)
        outer, w, pis = sc.stark_wrapper_circuit(cprog.fibonacci_program(), proof, max_queries=nq)
        assert pis[:64] == [x for d in pr["trace_cap"] for x in d] and pis[64:] == [int(x) for x in pi]
        assert _constraints_hold(outer, w, pis)
        if log_n == 6:
            _prove_and_verify(outer, w, pis)
    # the wrapper's structure depends on the table and the degree only
    t2, pi2 = syn.fibonacci_trace(9, seed=77)
    outer2, w2, _ = sc.stark_wrapper_circuit(cprog.fibonacci_program(), oracle.stark_prove(oracle.TABLE_FIBONACCI, t2, pi2), max_queries=1)
    assert (outer2.constants == outer.constants).all() and (outer2.sigmas == outer.sigmas).all() and not (w2 == w).all()


def test_memory_proof_with_lookups_verified_in_circuit():
    """The memory table (recalled constraint set + one logUp range check): lookup challenges drawn in-circuit, the lookup
    checks are part of the recorded program evaluated at zeta."""
    t = syn.memory_trace(7, seed=3)
    proof = oracle.stark_prove(oracle.TABLE_MEMORY, t)
    outer, w, pis = sc.stark_wrapper_circuit(cprog.memory_program(), proof, max_queries=1)
    assert len(pis) == 64
    assert _constraints_hold(outer, w, pis)
    _prove_and_verify(outer, w, pis)


@pytest.mark.parametrize("what", ["opening", "quotient", "final_poly", "pow", "query_leaf", "public_input"])
def test_tampered_table_proof_has_no_witness(what):
    t, pi = syn.fibonacci_trace(9, seed=4)
    proof = oracle.stark_prove(oracle.TABLE_FIBONACCI, t, pi)
    pr = V.parse_proof(proof)
    h = pr["h"]
    capw = 4 << h["cap_height"]
    off_open = V.HEADER_WORDS + 2 * capw  # no auxiliary cap
    n_open = 2 * (2 * h["n_trace"] + h["n_quot"])
    off_fri = off_open + n_open
    off = {"opening": off_open + 1, "quotient": off_open + 2 * 2 * h["n_trace"], "final_poly": proof.size - 1 - h["n_pi"] - 3,
           "pow": proof.size - 1 - h["n_pi"], "query_leaf": off_fri + h["n_layers"] * capw, "public_input": proof.size - 1}[what]
    bad = proof.copy()
    bad[off] = np.uint64((int(bad[off]) + 1) % P)
    with pytest.raises(V.VerifyError):
        V.verify(bad, max_queries=1)
    with pytest.raises(AssertionError):
        sc.stark_wrapper_circuit(cprog.fibonacci_program(), bad, max_queries=1)


def _prove_tables(tables, tamper=None):
    """prove_with_traces' shape on the oracle: -> (proofs, init challenger states, CTL challenges, final state)."""
    tids = [oracle.register_table_ex(p, p.aux_spec) for _, p, _ in tables]
    traces = [t.copy() for _, _, t in tables]
    if tamper:
        tamper(traces)
    batches = [oracle.Batch.from_values(t, 1, 4) for t in traces]
    ch = oracle.HostChallenger()
    for bb in batches:
        ch.observe(bb.cap)
    ctl_ch = ch.get_n(4)
    proofs, states = [], []
    for tid, t, bb in zip(tids, traces, batches):
        states.append(ch.compact())
        proofs.append(oracle.prove_with_commitment(tid, t, bb, ch, ctl_ch))
    return proofs, states, ctl_ch, ch.compact()


def _wrappers(tables, proofs, states, ctl_ch):
    out = []
    for (_, prog, _), proof, st in zip(tables, proofs, states):
        outer, w, pis = sc.stark_wrapper_circuit(prog, proof, st, ctl_ch, max_queries=1)
        out.append((outer, w, pis, sc.wrapper_public_input_layout(prog, True)))
    return out


def test_transaction_tables_wrapped_and_linked_by_the_root_circuit():
    """Three tables linked by cross-table lookups, proven on ONE transcript (prove_with_commitment): every table proof is
    verified by its wrapper circuit starting from the compacted challenger state before the table; the wrapper's exposed final
    state is the next table's initial state; the root circuit verifies the three wrapper proofs, re-derives the CTL challenges
    from the trace caps, chains the states and checks the CTL sums — its oracle proof is accepted."""
    tables, ctls = cprog.ctl_demo_tables(6, 5, 5)
    proofs, states, ctl_ch, final_state = _prove_tables(tables)
    wr = _wrappers(tables, proofs, states, ctl_ch)
    for (outer, w, pis, lay), st in zip(wr, states):
        assert len(pis) == lay["total"]
        o, n = lay["state_in"]
        assert pis[o:o + n] == [int(x) for x in st]
        o, n = lay["ctl_challenges"]
        assert pis[o:o + n] == [int(x) for x in ctl_ch]
        assert _constraints_hold(outer, w, pis)
    outs = [pis[lay["state_out"][0]:lay["state_out"][0] + 12] for _, _, pis, lay in wr]
    assert outs[:-1] == [[int(x) for x in s] for s in states[1:]] and outs[-1] == [int(x) for x in final_state]
    inner = [_prove_and_verify(outer, w, pis) for outer, w, pis, _ in wr]
    root, w, pis = sc.root_circuit(inner, [x[3] for x in wr], ctls, max_queries=1)
    assert len(pis) == 3 * 64 + 4 and pis[-4:] == [int(x) for x in ctl_ch]
    assert _constraints_hold(root, w, pis)
    _prove_and_verify(root, w, pis)
    # one shrinking step per table (same public inputs) and the root over the SHRUNK proofs: the chain the reference runs
    shrunk = []
    for x in inner:
        c, ws, ps = sc.shrink_circuit(x, max_queries=1)
        assert ps == x[2]
        shrunk.append(_prove_and_verify(c, ws, ps))
    root2, w2, pis2 = sc.root_circuit(shrunk, [x[3] for x in wr], ctls, max_queries=1)
    assert pis2 == pis and _constraints_hold(root2, w2, pis2)
    # tables in another order: the challenger chain breaks, the root has no witness
    with pytest.raises(AssertionError):
        sc.root_circuit([inner[1], inner[0], inner[2]], [wr[1][3], wr[0][3], wr[2][3]], ctls, max_queries=1)


def test_root_circuit_rejects_a_cross_table_lookup_mismatch():
    """A looked-table multiplicity that does not match: every table proof verifies, every wrapper builds — the root's
    verify_cross_table_lookups_circuit has no witness (upstream's division of labour)."""
    tables, ctls = cprog.ctl_demo_tables(6, 5, 5)

    def tamper(traces):
        traces[1][2, 3] = np.uint64(int(traces[1][2, 3]) + 1)  # rom MULT

    proofs, states, ctl_ch, _ = _prove_tables(tables, tamper)
    wr = _wrappers(tables, proofs, states, ctl_ch)
    inner = [_prove_and_verify(outer, w, pis) for outer, w, pis, _ in wr]
    with pytest.raises(AssertionError, match="copy constraint"):
        sc.root_circuit(inner, [x[3] for x in wr], ctls, max_queries=1)


def test_wrapper_rejects_a_proof_under_other_ctl_challenges():
    tables, _ = cprog.ctl_demo_tables(6, 5, 5)
    proofs, states, ctl_ch, _ = _prove_tables(tables)
    other = ctl_ch.copy()
    other[1] = np.uint64(int(other[1]) ^ 1)
    with pytest.raises(AssertionError):
        sc.stark_wrapper_circuit(tables[0][1], proofs[0], states[0], other, max_queries=1)
    with pytest.raises(AssertionError):  # another initial state: the in-circuit transcript diverges from the proof's
        sc.stark_wrapper_circuit(tables[0][1], proofs[0], states[1], ctl_ch, max_queries=1)


def test_aggregation_and_block_circuits_chain_the_public_values():
    """The layers above the root (/root/reference/ops/src/lib.rs:72,95): two transactions whose public values chain
    (s0 -> s1, s1 -> s2) are proven table by table, wrapped and rooted; the aggregation circuit verifies both root proofs and
    publishes (s0, s2); the block circuit verifies the aggregation proof; a second block verifies the first block proof as well.
    Two transactions that do NOT chain have no aggregation witness."""
    tables, ctls = cprog.ctl_demo_tables(5, 4, 4)
    s = [[11, 12, 13, 14], [21, 22, 23, 24], [31, 32, 33, 34]]

    def root_of(before, after):
        pv = before + after
        tids = [oracle.register_table_ex(p, p.aux_spec) for _, p, _ in tables]
        batches = [oracle.Batch.from_values(t, 1, 4) for _, _, t in tables]
        ch = oracle.HostChallenger()
        for bb in batches:
            ch.observe(bb.cap)
        ch.observe(pv)
        ctl_ch = ch.get_n(4)
        proofs, states = [], []
        for tid, (_, _, t), bb in zip(tids, tables, batches):
            states.append(ch.compact())
            proofs.append(oracle.prove_with_commitment(tid, t, bb, ch, ctl_ch))
        wr = _wrappers(tables, proofs, states, ctl_ch)
        inner = [_prove_and_verify(outer, w, pis) for outer, w, pis, _ in wr]
        root, w, pis = sc.root_circuit(inner, [x[3] for x in wr], ctls, public_values=pv, max_queries=1)
        assert pis[3 * 64:3 * 64 + 8] == pv
        return _prove_and_verify(root, w, pis)

    r01, r12 = root_of(s[0], s[1]), root_of(s[1], s[2])
    io = sc.root_io(3)
    agg, w, pis = sc.aggregation_circuit(r01, r12, io, io, max_queries=1)
    assert pis == s[0] + s[2] and _constraints_hold(agg, w, pis)
    a02 = _prove_and_verify(agg, w, pis)
    with pytest.raises(AssertionError, match="copy constraint"):
        sc.aggregation_circuit(r12, r01, io, io, max_queries=1)  # s2 != s0: the spans do not chain
    blk, w, pis = sc.block_circuit(a02, genesis_number=7, max_queries=1)
    assert pis == s[0] + s[2] + [7] and _constraints_hold(blk, w, pis)
    b0 = _prove_and_verify(blk, w, pis)
    # the next block: a span s2 -> s2 (the same transaction shape, chained), verified together with the previous block proof
    r22 = root_of(s[2], s[2])
    agg2, w, pis = sc.aggregation_circuit(r22, r22, io, io, max_queries=1)
    a22 = _prove_and_verify(agg2, w, pis)
    blk1, w, pis = sc.block_circuit(a22, prev_block=b0, max_queries=1)
    assert pis == s[0] + s[2] + [8] and _constraints_hold(blk1, w, pis)
    with pytest.raises(AssertionError, match="copy constraint"):
        sc.block_circuit(a02, prev_block=b0, max_queries=1)  # the block's span must start where the previous block ended


def test_transaction_and_block_plans_assemble_the_whole_job():
    """transaction_recursion_plan + block_recursion_plan (what bench.py's tx_real_recursion leg and tools/tx_recursion.py drive on
    the device) with the oracle as the circuit prover: wrapper + shrink per table, root, two aggregation levels, block — every later
    circuit is built from the proofs made before it; re-proving the list reproduces the proofs."""
    tables, ctls = cprog.ctl_demo_tables(5, 4, 4)
    pv = [5, 6, 7, 8, 5, 6, 7, 8]
    tids = [oracle.register_table_ex(p, p.aux_spec) for _, p, _ in tables]
    batches = [oracle.Batch.from_values(t, 1, 4) for _, _, t in tables]
    ch = oracle.HostChallenger()
    for bb in batches:
        ch.observe(bb.cap)
    ch.observe(pv)
    ctl_ch = ch.get_n(4)
    proofs, states = [], []
    for tid, (_, _, t), bb in zip(tids, tables, batches):
        states.append(ch.compact())
        proofs.append(oracle.prove_with_commitment(tid, t, bb, ch, ctl_ch))
    ap = types.SimpleNamespace(stark_proofs=proofs, init_challenger_states=states, ctl_challenges=ctl_ch)

    def circuit_prove(c, w, p):
        fake, words, _ = _prove_and_verify(c, w, p)
        return fake, words

    plan = sc.transaction_recursion_plan(tables, ctls, ap, circuit_prove, max_queries=1, public_values=pv)
    assert [(s["name"], s["kind"]) for s in plan] == [("ops", "wrapper"), ("ops", "shrink"), ("rom", "wrapper"), ("rom", "shrink"),
                                                      ("extra", "wrapper"), ("extra", "shrink"), ("root", "root")]
    assert plan[-1]["public_inputs"][3 * 64:3 * 64 + 8] == pv
    blocks = sc.block_recursion_plan(plan[-1], len(tables), circuit_prove, levels=2, max_queries=1)
    assert [s["name"] for s in blocks] == ["agg1", "agg2", "block"]
    assert blocks[0]["public_inputs"] == pv and blocks[-1]["public_inputs"] == pv + [0]
    again = _prove_and_verify(plan[-1]["circuit"], plan[-1]["wires"], plan[-1]["public_inputs"])[1]
    assert (again == plan[-1]["words"]).all()
