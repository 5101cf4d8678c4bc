"""GPU tests of the first recursion layer (eth_tx_proof_b200/stark_circuit.py; starky 0.4.0 src/recursive_verifier.rs,
evm_arithmetization 0.1.3 src/fixed_recursive_verifier.rs recursive_stark_circuit / create_root_circuit —
/root/reference/Cargo.lock:4529,1675, reached from /root/reference/ops/src/lib.rs:52): table proofs made ON THE DEVICE are
verified in-circuit by their wrapper circuits, the wrapper circuits are proven on the device (word for word the oracle's proof),
and the root circuit that links the tables of one transaction is proven on the device and accepted by the Python verifier."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

P = 0xFFFFFFFF00000001


@pytest.fixture(scope="module")
def ctx():
    import eth_tx_proof_b200 as etp

    c = etp.Context(0)
    yield c
    c.close()


def dev(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


def test_device_table_proofs_verified_in_circuit_and_wrapped_on_the_device(ctx):
    """Fibonacci (2^9 rows, one FRI layer) and the memory table with its logUp lookup (2^8 rows) proven by etp_stark_prove,
    ALL 84 queries verified by the wrapper circuit; the wrapper is proven by the device circuit prover: proof == the oracle's."""
    import oracle
    import plonk_verifier
    import eth_tx_proof_b200 as etp
    from eth_tx_proof_b200 import circuit as cc, cprog, stark_circuit as sc, synthetic as syn

    t, pi = syn.fibonacci_trace(9, seed=5)
    cases = [(cprog.fibonacci_program(), ctx.stark_prove(etp.TABLE_FIBONACCI, t, pi)),
             (cprog.memory_program(), ctx.stark_prove(etp.TABLE_MEMORY, syn.memory_trace(8, seed=2)))]
    for program, words in cases:
        outer, w, pis = sc.stark_wrapper_circuit(program, words)
        cp = cc.CircuitProver(ctx, outer)
        got = cp.prove(w, pis)
        plonk_verifier.verify(got, outer, cp.constants_sigmas_cap, cp.digest, max_queries=2)
        want = oracle.circuit_prove(outer, w, pis, cp.digest)
        assert (got["quotient_polys_cap"] == want["quotient_polys_cap"]).all() and (got["opening_proof"] == want["opening_proof"]).all()
        del cp
    bad = cases[0][1].copy()
    bad[bad.size - 1] = np.uint64((int(bad[bad.size - 1]) + 1) % P)  # the claimed last Fibonacci value
    with pytest.raises(AssertionError):
        sc.stark_wrapper_circuit(cases[0][0], bad, max_queries=1)


def test_transaction_wrappers_and_root_circuit_on_the_device(ctx):
    """prove_with_traces on the device (three tables, one transcript, CTLs) -> one wrapper circuit per table, started from the
    table's init_challenger_state -> wrapper proofs on the device -> the root circuit (CTL challenges re-derived from the trace
    caps, challenger chain, cross-table lookup sums) -> root proof on the device, accepted by the Python verifier."""
    import torch

    import plonk_verifier
    from eth_tx_proof_b200 import circuit as cc, cprog, prover, stark_circuit as sc

    tables, ctls = cprog.ctl_demo_tables(7, 6, 5)
    tids = [ctx.register_table(p) for _, p, _ in tables]
    devs = [dev(t) for _, _, t in tables]
    torch.cuda.synchronize()
    traces_dev = [(d.data_ptr(), t.shape[1], t.shape[0], int(t.shape[1]).bit_length() - 1) for d, (_, _, t) in zip(devs, tables)]
    got = prover.prove_with_traces(ctx, tids, traces_dev)
    inner, layouts, provers = [], [], []
    for (_, prog, _), words, state in zip(tables, got.stark_proofs, got.init_challenger_states):
        outer, w, pis = sc.stark_wrapper_circuit(prog, words, state, got.ctl_challenges, max_queries=4)
        lay = sc.wrapper_public_input_layout(prog, True)
        assert len(pis) == lay["total"]
        cp = cc.CircuitProver(ctx, outer)
        inner.append((cp, cp.prove_words(w, pis), pis))
        layouts.append(lay)
        provers.append(cp)
    for (_, _, pis), lay, nxt in zip(inner, layouts, got.init_challenger_states[1:]):
        o, n = lay["state_out"]
        assert pis[o:o + n] == [int(x) for x in nxt]  # the in-circuit transcript ends where the next table's begins
    root, w, pis = sc.root_circuit(inner, layouts, ctls, max_queries=2)
    assert pis[:64] == [int(x) for x in np.asarray(got.trace_caps[0]).ravel()] and pis[-4:] == [int(x) for x in got.ctl_challenges]
    rp = cc.CircuitProver(ctx, root)
    plonk_verifier.verify(rp.prove(w, pis), root, rp.constants_sigmas_cap, rp.digest, max_queries=1)
    with pytest.raises(AssertionError):  # tables in another order: the challenger chain breaks
        sc.root_circuit([inner[1], inner[0], inner[2]], [layouts[1], layouts[0], layouts[2]], ctls, max_queries=1)
