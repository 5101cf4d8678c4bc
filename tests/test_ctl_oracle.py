"""Cross-table lookups end to end on the CPU: the oracle's restatement of evm_arithmetization's prove_with_traces shape
(all trace caps observed by ONE challenger, CTL challenges, then starky::prover::prove_with_commitment per table with
ctl_data, sharing the challenger) on a synthetic three-table system (cprog.ctl_demo_tables), checked by the independent
Python verifier including verify_cross_table_lookups.  Upstream: starky 0.4.0 src/cross_table_lookup.rs, src/prover.rs;
evm_arithmetization 0.1.3 src/prover.rs (/root/reference/Cargo.lock:4529,1675; reached from /root/reference/ops/src/lib.rs:52)."""
import numpy as np
import pytest

import oracle
import stark_verifier as V
from eth_tx_proof_b200 import cprog

P = cprog.P


def prove_all(tables, tamper=None):
    """-> (proofs, trace caps)."""
    tids = [oracle.register_table_ex(p, p.aux_spec) for _, p, _ in tables]
    traces = [t.copy() for _, _, t in tables]
    if tamper:
        tamper(traces)
    batches = [oracle.Batch.from_values(t, 1, 4) for t in traces]
    ch = oracle.HostChallenger()
    for b in batches:
        ch.observe(b.cap)
    ctl_ch = ch.get_n(4)  # get_grand_product_challenge_set: (beta, gamma) x num_challenges
    proofs = []
    for tid, t, b in zip(tids, traces, batches):
        ch.compact()  # prove_single_table: init_challenger_state = challenger.compact()
        proofs.append(oracle.prove_with_commitment(tid, t, b, ch, ctl_ch))
    return proofs, [b.cap for b in batches]


def verify_all(tables, ctls, proofs, caps, max_queries=3):
    vch = V.new_challenger()
    for cap in caps:
        for d in cap:
            vch.observe([int(x) for x in d])
    raw = vch.get_n(4)
    ctl = [(raw[0], raw[1]), (raw[2], raw[3])]
    zs = []
    for (name, prog, _), proof, cap in zip(tables, proofs, caps):
        vch.compact()
        pr = V.verify(proof, program=prog, challenger=vch, ctl_challenges=ctl, max_queries=max_queries)
        assert pr["trace_cap"] == [[int(x) for x in d] for d in cap]
        zs.append(pr["ctl_zs_first"])
    V.verify_cross_table_lookups(ctls, zs)
    return zs


def test_ctl_aux_columns_definition():
    """partial_sums: Z[i] = sum_{j >= i} sum_sets filter_j / combine_j; helpers per chunk of two; Z(1) sums match across tables."""
    tables, ctls = cprog.ctl_demo_tables(5, 4, 4)
    beta0, gamma0, beta1, gamma1 = 11111, 22222, 33333, 44444
    first = []
    for name, prog, trace in tables:
        tid = oracle.register_table_ex(prog, prog.aux_spec)
        aux = oracle.aux_columns(tid, trace, [beta0, beta1], [beta0, gamma0, beta1, gamma1])
        assert aux.shape[0] == prog.n_aux
        n = trace.shape[1]
        t = [[int(x) for x in col] for col in trace]
        z_base = prog.n_lookup_cols + prog.n_ctl_helper_cols
        h_at = prog.n_lookup_cols
        zf = []
        for zi, (k, sets) in enumerate(prog.ctl_zs):
            beta, gamma = (beta0, gamma0) if k == 0 else (beta1, gamma1)
            hsum = [0] * n
            for i in range(n):
                for cols, filt in sets:
                    comb = (sum(c.eval_table(t, i) * pow(beta, j, P) for j, c in enumerate(cols)) + gamma) % P
                    hsum[i] = (hsum[i] + filt.eval_table(t, i) * pow(comb, P - 2, P)) % P
            if len(sets) > 1:
                assert [int(x) for x in aux[h_at]] == hsum  # one chunk of two sets -> the helper IS the row sum
                h_at += 1
            acc = 0
            z = [0] * n
            for i in reversed(range(n)):
                acc = (acc + hsum[i]) % P
                z[i] = acc
            assert [int(x) for x in aux[z_base + zi]] == z
            zf.append(z[0])
        first.append(zf)
    V.verify_cross_table_lookups(ctls, first)


def test_multi_table_ctl_proofs_verify():
    tables, ctls = cprog.ctl_demo_tables(6, 5, 5)
    proofs, caps = prove_all(tables)
    zs = verify_all(tables, ctls, proofs, caps)
    assert all(len(z) == 2 for z in zs)
    # the proofs are not valid stand-alone (their transcript starts from the shared challenger)
    with pytest.raises(V.VerifyError):
        V.verify(proofs[1], program=tables[1][1], max_queries=1)


def test_ctl_mismatch_is_rejected():
    """A looked-table multiplicity that does not match the looking tables: every single proof still verifies (each table's
    Z is internally consistent) but verify_cross_table_lookups fails — exactly upstream's division of labour."""
    tables, ctls = cprog.ctl_demo_tables(6, 5, 5)

    def tamper(traces):
        traces[1][2, 3] = np.uint64(int(traces[1][2, 3]) + 1)  # rom MULT

    proofs, caps = prove_all(tables, tamper)
    with pytest.raises(V.VerifyError, match="Cross-table lookup"):
        verify_all(tables, ctls, proofs, caps, max_queries=1)


def test_tampered_ctl_opening_is_rejected():
    tables, ctls = cprog.ctl_demo_tables(6, 5, 5)
    proofs, caps = prove_all(tables)
    pr = V.parse_proof(proofs[0])
    h = pr["h"]
    off = V.HEADER_WORDS + 3 * (4 << h["cap_height"]) + 2 * (2 * h["n_trace"] + 2 * h["n_aux"])  # first ctl_zs_first word
    bad = [p.copy() for p in proofs]
    bad[0][off] = np.uint64((int(bad[0][off]) + 1) % P)
    with pytest.raises(V.VerifyError):
        verify_all(tables, ctls, bad, caps, max_queries=2)


def test_evm_shaped_transaction_proves_and_verifies():
    """Seven tables of the evm_arithmetization shapes linked by the upstream CTL topology (cprog.evm_shaped_system): cpu ->
    arithmetic / byte packing / keccak sponge / logic / memory x3, keccak sponge -> keccak x2 / logic / memory, byte packing ->
    memory.  One shared transcript, CTL challenges, per-table prove_with_commitment, then verify_cross_table_lookups."""
    bits = {"arithmetic": 6, "byte_packing": 5, "cpu": 6, "keccak": 5, "keccak_sponge": 5, "logic": 5, "memory": 7}
    tables, ctls = cprog.evm_shaped_system(degree_bits=bits)
    assert [n for n, _, _ in tables] == list(cprog.EVM_TABLE_ORDER)
    proofs, caps = prove_all(tables)
    zs = verify_all(tables, ctls, proofs, caps, max_queries=1)
    assert [len(z) for z in zs] == [2, 4, 10, 4, 10, 2, 2]
