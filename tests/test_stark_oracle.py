"""CPU tests: the oracle's single-table STARK prover against the independent Python verifier
(tests/stark_verifier.py) — the stand-in for starky's own fibonacci_stark.rs / stark_testing.rs tests
and evm_arithmetization's per-table prove+verify tests (none on disk)."""
import numpy as np
import pytest

import oracle
import stark_verifier as V
from eth_tx_proof_b200 import synthetic as syn
from oracle import pyref as R

P = R.P


@pytest.mark.parametrize("log_n", [5, 6, 9, 10, 13])
def test_fibonacci_prove_verify(log_n):
    t, pi = syn.fibonacci_trace(log_n, seed=log_n)
    assert oracle.check_constraints(oracle.TABLE_FIBONACCI, t, pi) == -1
    proof = oracle.stark_prove(oracle.TABLE_FIBONACCI, t, pi)
    pr = V.verify(proof)
    assert pr["h"]["n_layers"] == {5: 0, 6: 0, 9: 1, 10: 1, 13: 2}[log_n]


@pytest.mark.parametrize("log_n", [5, 8, 11])
def test_memory_prove_verify(log_n):
    t = syn.memory_trace(log_n, seed=log_n)
    assert oracle.check_constraints(oracle.TABLE_MEMORY, t) == -1
    proof = oracle.stark_prove(oracle.TABLE_MEMORY, t)
    V.verify(proof)


def test_pure_python_hash_path_agrees():
    t = syn.memory_trace(6)
    proof = oracle.stark_prove(oracle.TABLE_MEMORY, t)
    V.verify(proof, fast=False, max_queries=3)


def test_bad_trace_is_rejected():
    t = syn.memory_trace(7)
    t[syn.M_VALUE0, 40] ^= np.uint64(1)  # break a read-consistency or write row
    t[syn.M_IS_READ, 41] = 1
    t[syn.M_IS_READ, 40] = 1
    bad_row = oracle.check_constraints(oracle.TABLE_MEMORY, t)
    if bad_row == -1:
        pytest.skip("mutation happened to keep the trace valid")
    proof = oracle.stark_prove(oracle.TABLE_MEMORY, t)  # quotient is not a polynomial any more
    with pytest.raises(V.VerifyError):
        V.verify(proof)


@pytest.mark.parametrize("field", ["trace_cap", "opening", "fri_cap", "query_leaf", "final_poly", "pow"])
def test_tampered_proof_is_rejected(field):
    t = syn.memory_trace(9)
    proof = oracle.stark_prove(oracle.TABLE_MEMORY, t)
    pr = V.parse_proof(proof)
    h = pr["h"]
    capw = 4 << h["cap_height"]
    off_open = V.HEADER_WORDS + 3 * capw
    n_open = 2 * (2 * h["n_trace"] + 2 * h["n_aux"] + h["n_quot"])
    off_fri = off_open + n_open
    pos = {"trace_cap": V.HEADER_WORDS + 5, "opening": off_open + 7, "fri_cap": off_fri + 3,
           "query_leaf": off_fri + capw * h["n_layers"] + 2, "final_poly": len(proof) - 2 - 2 * h["final_len"] + 1,
           "pow": len(proof) - 1}[field]
    bad = proof.copy()
    bad[pos] = np.uint64((int(bad[pos]) + 1) % P)
    with pytest.raises(V.VerifyError):
        V.verify(bad)


def test_lookup_helper_columns_definition():
    t = syn.memory_trace(6)
    ch = [12345678901234567, 98765432109876543]
    aux = oracle.lookup_helper_columns(oracle.TABLE_MEMORY, t, ch)
    n = t.shape[1]
    for k, c in enumerate(ch):
        h, z = aux[2 * k], aux[2 * k + 1]
        for i in range(n):
            assert int(h[i]) * ((int(t[syn.M_RANGE_CHECK, i]) + c) % P) % P == 1
        assert int(z[0]) == 0
        acc = 0
        for i in range(n):
            assert int(z[i]) == acc
            acc = (acc + int(h[i]) - int(t[syn.M_FREQ, i]) * pow((int(t[syn.M_COUNTER, i]) + c) % P, P - 2, P)) % P
        assert acc == 0  # logUp sum closes: the wrap-around Z constraint holds


def test_quotient_is_low_degree_and_matches_constraints_at_random_point():
    t = syn.memory_trace(6)
    tb = oracle.Batch.from_values(t, 1, 4)
    ch = [3, 5]
    aux = oracle.lookup_helper_columns(oracle.TABLE_MEMORY, t, ch)
    ab = oracle.Batch.from_values(aux, 1, 4)
    alphas = [1111111111, 2222222222]
    q = oracle.compute_quotient_polys(oracle.TABLE_MEMORY, tb, ab, ch, [], alphas)
    assert q.shape == (4, 64)
    # evaluate everything at a base-field point outside H and compare t(x) * Z_H(x) with the constraints
    x = 123456789
    n = 64
    g = R.root_of_unity(6)
    lv = [(R.eval_poly([int(v) for v in c], x), 0) for c in tb.coeffs]
    nv = [(R.eval_poly([int(v) for v in c], x * g % P), 0) for c in tb.coeffs]
    al = [(R.eval_poly([int(v) for v in c], x), 0) for c in ab.coeffs]
    an = [(R.eval_poly([int(v) for v in c], x * g % P), 0) for c in ab.coeffs]
    zx = (pow(x, n, P) - 1) % P
    l_first = zx * pow(n * (x - 1) % P, P - 2, P) % P
    l_last = zx * pow(n * (g * x - 1) % P, P - 2, P) % P
    cons = V._Consumer(alphas, ((x - pow(g, P - 2, P)) % P, 0), (l_first, 0), (l_last, 0))
    V._eval_memory(lv, nv, [], cons)
    V._eval_memory_lookups(lv, al, an, ch, cons)
    for j in range(2):
        tq = (R.eval_poly([int(v) for v in q[2 * j]], x) + pow(x, n, P) * R.eval_poly([int(v) for v in q[2 * j + 1]], x)) % P
        assert cons.acc[j] == (tq * zx % P, 0)
