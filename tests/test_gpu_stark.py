"""GPU parity tests (through the C ABI) for the starky half of the path: lookup helper columns,
compute_quotient_polys, proof of work and the complete single-table proof — bit-identical to the
oracle's proof on the same trace, and accepted by the independent verifier (tests/stark_verifier.py).
Stands in for starky's fibonacci_stark.rs tests (test_fibonacci_stark) and evm_arithmetization's
per-table prove+verify tests, none of which is on disk."""
import random

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


def _dev(arr):
    import torch

    return torch.from_numpy(arr.view(np.int64)).cuda()


@pytest.mark.parametrize("log_n", [5, 9, 13])
def test_lookup_helper_columns(ctx, log_n):
    import torch

    import eth_tx_proof_b200 as etp
    import oracle
    from eth_tx_proof_b200 import synthetic as syn

    t = syn.memory_trace(log_n, seed=log_n)
    ch = [0x123456789ABCDEF0 % P, 0x0FEDCBA987654321]
    want = oracle.lookup_helper_columns(oracle.TABLE_MEMORY, t, ch)
    d = _dev(t)
    aux = torch.zeros((4, 1 << log_n), dtype=torch.int64, device="cuda")
    ctx.lookup_helper_columns_dev(etp.TABLE_MEMORY, log_n, d.data_ptr(), 1 << log_n, ch, aux.data_ptr())
    got = aux.cpu().numpy().view(np.uint64)
    assert (got == want).all()


@pytest.mark.parametrize("table,log_n", [(0, 5), (0, 10), (1, 5), (1, 8), (1, 12), (1, 13)])
def test_compute_quotient_polys(ctx, table, log_n):
    import eth_tx_proof_b200 as etp
    import oracle
    from eth_tx_proof_b200 import synthetic as syn

    rng = random.Random(log_n)
    alphas = [rng.randrange(P), rng.randrange(P)]
    if table == 0:
        t, pi = syn.fibonacci_trace(log_n, seed=3)
        ch, aux, oaux = [], None, None
        ob = oracle.Batch.from_values(t, 1, 4)
        gb = etp.PolynomialBatch.from_values(ctx, t, 1, False, 4)
    else:
        t, pi = syn.memory_trace(log_n, seed=4), []
        ch = [rng.randrange(P), rng.randrange(P)]
        a = oracle.lookup_helper_columns(oracle.TABLE_MEMORY, t, ch)
        ob = oracle.Batch.from_values(t, 1, 4)
        oaux = oracle.Batch.from_values(a, 1, 4)
        gb = etp.PolynomialBatch.from_values(ctx, t, 1, False, 4)
        aux = etp.PolynomialBatch.from_values(ctx, a, 1, False, 4)
    want = oracle.compute_quotient_polys(table, ob, oaux, ch, pi, alphas)
    got = ctx.compute_quotient_polys(table, gb, aux, ch, pi, alphas)
    assert got.shape == want.shape
    assert (got == want).all()


def test_pow_grind_smallest(ctx):
    import oracle

    rng = random.Random(2)
    for bits, pos in [(8, 3), (12, 0), (16, 5)]:
        st = [rng.randrange(P) for _ in range(12)]
        assert ctx.pow_grind(st, pos, bits) == oracle.pow_grind(st, pos, bits)


@pytest.mark.parametrize("table,log_n", [(0, 5), (0, 6), (0, 9), (0, 13), (1, 5), (1, 8), (1, 11), (1, 13)])
def test_stark_proof_bit_identical_and_verifies(ctx, table, log_n):
    import oracle
    import stark_verifier as V
    from eth_tx_proof_b200 import synthetic as syn

    if table == 0:
        t, pi = syn.fibonacci_trace(log_n, seed=log_n)
    else:
        t, pi = syn.memory_trace(log_n, seed=log_n), []
    got = ctx.stark_prove(table, t, pi)
    want = oracle.stark_prove(table, t, pi)
    assert got.shape == want.shape
    diff = np.nonzero(got != want)[0]
    assert diff.size == 0, f"first differing word {diff[:5]} of {got.size}"
    V.verify(got)
    tm = ctx.last_prove_timings()
    assert "compute quotient polys" in tm


def test_stark_proof_larger_size_verifies(ctx):
    """2^17-row memory-shaped table: beyond what the oracle proves in seconds, so check the
    size-independent property instead: the independent verifier accepts the proof."""
    import stark_verifier as V
    from eth_tx_proof_b200 import synthetic as syn

    t = syn.memory_trace(17, seed=99)
    proof = ctx.stark_prove(1, t, [])
    V.verify(proof)


def test_invalid_trace_yields_a_rejected_proof(ctx):
    """Like upstream in release mode (check_constraints is debug-only), the prover cannot notice a bad
    trace when quotient_degree_factor is a power of two: the quotient interpolant always exists.  The
    proof must then be REJECTED by the verifier, and must still equal the oracle's bytes."""
    import oracle
    import stark_verifier as V
    from eth_tx_proof_b200 import synthetic as syn

    t = syn.memory_trace(10, seed=5)
    t[syn.M_COUNTER, 17] += np.uint64(1)  # breaks the counter transition and the lookup table
    assert oracle.check_constraints(oracle.TABLE_MEMORY, t) != -1
    proof = ctx.stark_prove(1, t, [])
    assert (proof == oracle.stark_prove(oracle.TABLE_MEMORY, t)).all()
    with pytest.raises(V.VerifyError):
        V.verify(proof)


def test_prove_dev_matches_prove_host(ctx):
    from eth_tx_proof_b200 import synthetic as syn

    t = syn.memory_trace(12, seed=8)
    d = _dev(t)
    a = ctx.stark_prove(1, t, [])
    b = ctx.stark_prove_dev(1, 12, d.data_ptr(), 1 << 12, [])
    assert (a == b).all()


@pytest.mark.gpu
def test_prover_pool_proofs_are_identical_to_single_context(ctx):
    """Two prover contexts on one GPU (parallel.ProverPool) produce the same proofs as one context, job by job."""
    import torch

    import eth_tx_proof_b200 as etp
    from eth_tx_proof_b200 import parallel, synthetic as syn

    log_n = 10
    traces = [torch.from_numpy(syn.memory_trace(log_n, seed=s).view(np.int64)).cuda() for s in range(5)]
    torch.cuda.synchronize()
    want = [ctx.stark_prove_dev(etp.TABLE_MEMORY, log_n, t.data_ptr(), 1 << log_n) for t in traces]
    pool = parallel.ProverPool(0, 2)
    try:
        got = pool.stark_prove_dev(etp.TABLE_MEMORY, log_n, [(t.data_ptr(), 1 << log_n) for t in traces])
    finally:
        pool.close()
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert (g == w).all()
    with pytest.raises(ValueError):
        parallel.ProverPool(0, 0)
