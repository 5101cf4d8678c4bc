"""Committed golden vectors of the hot path (tests/golden/path_vectors.json, written by tools/gen_golden.py): commit caps,
coefficients, digests, leaf rows, Merkle paths and whole proofs on seeded inputs, as SHA-256 digests.

They are regression anchors produced by this repository's oracle — the reference's tests hold no vector for this path
 — and are checked twice: the oracle must still reproduce them (CPU), and the CUDA library must reproduce
them WITHOUT the oracle in the loop (GPU), so a change that moved oracle and product together would be caught."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_golden as gg  # noqa: E402

with open(os.path.join(ROOT, "tests", "golden", "path_vectors.json")) as _f:
    GOLD = json.load(_f)


with open(os.path.join(ROOT, "tests", "golden", "circuit_vectors.json")) as _f:
    GOLD_CIRCUITS = json.load(_f)


@pytest.mark.parametrize("k", range(len(gg.CIRCUITS)))
def test_oracle_reproduces_golden_circuit_proofs(k):
    assert [tuple(c["shape"]) for c in GOLD_CIRCUITS["circuits"]] == gg.CIRCUITS
    assert gg.circuit_case(*gg.CIRCUITS[k]) == GOLD_CIRCUITS["circuits"][k]


def test_golden_file_covers_the_generator_cases():
    assert [tuple(c["shape"]) for c in GOLD["commits"]] == gg.COMMITS
    assert [(p["table"], p["log_n"], p["seed"]) for p in GOLD["proofs"]] == gg.PROOFS


@pytest.mark.parametrize("k", range(len(gg.COMMITS)))
def test_oracle_reproduces_golden_commits(k):
    assert gg.commit_case(*gg.COMMITS[k]) == GOLD["commits"][k]


@pytest.mark.parametrize("k", range(len(gg.PROOFS)))
def test_oracle_reproduces_golden_proofs(k):
    assert gg.proof_case(*gg.PROOFS[k]) == GOLD["proofs"][k]


@pytest.fixture(scope="module")
def ctx():
    import eth_tx_proof_b200 as etp

    c = etp.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("k", range(len(gg.COMMITS)))
def test_cuda_reproduces_golden_commits(ctx, k):
    import eth_tx_proof_b200 as etp
    from eth_tx_proof_b200 import synthetic as syn

    g = GOLD["commits"][k]
    n_cols, log_n, rate_bits, cap_height, seed = g["shape"]
    b = etp.PolynomialBatch.from_values(ctx, syn.random_columns(n_cols, log_n, seed=seed), rate_bits, False, cap_height)
    assert gg.sha(b.cap) == g["cap_sha256"]
    assert [f"{int(x):016x}" for x in b.cap.reshape(-1)[:4]] == g["cap_first"]
    assert gg.sha(b.polynomials) == g["coeffs_sha256"]
    assert gg.sha(b.digests) == g["digests_sha256"]
    assert gg.sha(b.leaves_at(g["leaf_rows"])) == g["leaf_rows_sha256"]
    if (1 << (log_n + rate_bits)) > (1 << cap_height):
        assert gg.sha(np.concatenate([np.asarray(b.prove(i)).reshape(-1) for i in g["leaf_rows"]])) == g["paths_sha256"]


@pytest.mark.gpu
@pytest.mark.parametrize("k", range(len(gg.PROOFS)))
def test_cuda_reproduces_golden_proofs(ctx, k):
    import eth_tx_proof_b200 as etp

    g = GOLD["proofs"][k]
    name, prog, trace, pi = gg.proof_inputs(g["table"], g["log_n"], g["seed"])
    if prog is None:
        tid = etp.TABLE_FIBONACCI if name == "fibonacci" else etp.TABLE_MEMORY
    else:
        tid = ctx.register_table(prog, prog.lookups)
    proof = ctx.stark_prove(tid, trace, pi) if len(pi) else ctx.stark_prove(tid, trace)
    assert int(proof.size) == g["words"]
    assert gg.sha(np.concatenate([proof[:1], proof[2:]])) == g["proof_sha256_without_table_id"]


@pytest.mark.gpu
@pytest.mark.parametrize("k", range(len(gg.CIRCUITS)))
def test_cuda_reproduces_golden_circuit_proofs(ctx, k):
    """etp_circuit_create + etp_circuit_prove_host against the committed digests, the oracle not in the loop."""
    from eth_tx_proof_b200 import circuit as cc

    g = GOLD_CIRCUITS["circuits"][k]
    circ, wires, pis = gg.circuit_inputs(*g["shape"])
    assert len(circ.program.ops) == g["program_ops"] and circ.num_constants == g["num_constants"]
    prover = cc.CircuitProver(ctx, circ)
    assert [f"{x:016x}" for x in prover.digest] == g["digest"]
    assert gg.sha(prover.constants_sigmas_cap) == g["constants_sigmas_cap_sha256"]
    pr = prover.prove(wires, pis)
    op = pr["openings"]
    assert gg.sha(pr["wires_cap"]) == g["wires_cap_sha256"]
    assert gg.sha(pr["plonk_zs_partial_products_cap"]) == g["zs_partial_products_cap_sha256"]
    assert gg.sha(pr["quotient_polys_cap"]) == g["quotient_polys_cap_sha256"]
    assert gg.sha(np.concatenate([np.asarray(op[k2]).reshape(-1) for k2 in
                                  ("constants_sigmas", "wires", "zs_partial_products", "quotient_polys", "plonk_zs_next")])) == g["openings_sha256"]
    assert int(pr["opening_proof"].size) == g["opening_proof_words"] and gg.sha(pr["opening_proof"]) == g["opening_proof_sha256"]
