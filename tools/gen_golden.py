#!/usr/bin/env python
"""Writes tests/golden/path_vectors.json: outputs of the ORACLE on seeded inputs of the hot path (commit caps, digests, LDE rows,
Merkle paths, quotient coefficients, whole proofs), as SHA-256 digests plus a few literal words.

These are regression anchors, not upstream vectors: the reference's own tests hold none for this path, so the
file pins what THIS repository's oracle computed when it was written.  tests/test_golden_vectors.py checks that the oracle
still reproduces them (CPU) and that the CUDA library reproduces them without consulting the oracle at all (GPU) — a change
that moves the oracle and the product together no longer goes unnoticed.  Inputs come from eth_tx_proof_b200.synthetic /
cprog (counter-based generators), so nothing but the seeds is stored.

    python tools/gen_golden.py          # rewrite the file (review the diff!)
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402  (dev-time generator of TEST fixtures)
from eth_tx_proof_b200 import cprog, synthetic as syn  # noqa: E402

COMMITS = [  # (n_cols, log_n, rate_bits, cap_height, seed)
    (3, 4, 1, 2, 101), (9, 5, 1, 4, 102), (4, 6, 1, 4, 103), (21, 8, 1, 4, 104), (128, 10, 1, 4, 105), (17, 7, 2, 3, 106), (135, 6, 1, 0, 107),
]
PROOFS = [  # (table, log_n, seed)
    ("fibonacci", 5, 1), ("fibonacci", 9, 2), ("memory", 6, 3), ("memory", 10, 4), ("shape:37:3", 7, 5), ("shape:80:8", 8, 6), ("logic:1", 6, 7),
]


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(np.asarray(a, dtype=np.uint64)).tobytes()).hexdigest()


def commit_case(n_cols, log_n, rate_bits, cap_height, seed):
    vals = syn.random_columns(n_cols, log_n, seed=seed)
    b = oracle.Batch.from_values(vals, rate_bits, cap_height)
    n_leaves = 1 << (log_n + rate_bits)
    idx = sorted({0, 1, n_leaves // 3, n_leaves - 1})
    return {"shape": [n_cols, log_n, rate_bits, cap_height, seed], "cap_sha256": sha(b.cap), "cap_first": [f"{int(x):016x}" for x in b.cap.reshape(-1)[:4]],
            "coeffs_sha256": sha(b.coeffs), "digests_sha256": sha(b.digests), "leaf_rows": idx, "leaf_rows_sha256": sha(b.leaves[idx]),
            "paths_sha256": sha(np.concatenate([np.asarray(oracle.merkle_prove(b.digests, n_leaves, cap_height, i)).reshape(-1) for i in idx]))
            if n_leaves > (1 << cap_height) else sha(np.zeros(0, dtype=np.uint64))}


def proof_inputs(table, log_n, seed):
    """-> (oracle table id or name, program or None, trace, public inputs)"""
    if table == "fibonacci":
        t, pi = syn.fibonacci_trace(log_n, seed=seed)
        return "fibonacci", None, t, pi
    if table == "memory":
        return "memory", None, syn.memory_trace(log_n, seed=seed), ()
    if table.startswith("shape:"):
        _, cols, lk = table.split(":")
        return table, cprog.shape_program(int(cols), int(lk)), cprog.shape_trace(log_n, int(cols), int(lk), seed=seed), ()
    if table.startswith("logic:"):
        limbs = int(table.split(":")[1])
        return table, cprog.logic_program(limbs), cprog.logic_trace(log_n, limbs, seed=seed), ()
    raise ValueError(table)


def proof_case(table, log_n, seed):
    name, prog, trace, pi = proof_inputs(table, log_n, seed)
    if prog is None:
        tid = oracle.TABLE_FIBONACCI if name == "fibonacci" else oracle.TABLE_MEMORY
    else:
        tid = oracle.register_table(prog, prog.lookups)
    proof = oracle.stark_prove(tid, trace, pi)
    # word 1 is the table id (differs for registered tables): excluded, as in the parity tests
    return {"table": table, "log_n": log_n, "seed": seed, "words": int(proof.size), "proof_sha256_without_table_id": sha(np.concatenate([proof[:1], proof[2:]])),
            "pow_witness": f"{int(proof[-1 - len(pi)]):016x}" if len(pi) else f"{int(proof[-1]):016x}"}


CIRCUITS = [  # (degree_bits, seed, all fourteen gates?)
    (5, 1, False), (7, 2, False), (6, 3, True),
]


def circuit_inputs(degree_bits, seed, all_gates):
    from eth_tx_proof_b200 import circuit as cc

    return cc.hash_chain_circuit(degree_bits, seed=seed, all_gates=all_gates)


def circuit_case(degree_bits, seed, all_gates):
    """A circuit proof (eth_tx_proof_b200/circuit.py; plonk::prover::prove restated): the oracle's proof with the stand-in digest
    the library derives (hash_no_pad(constants_sigmas cap ++ degree_bits))."""
    from eth_tx_proof_b200 import circuit as cc

    circ, wires, pis = circuit_inputs(degree_bits, seed, all_gates)
    cs = oracle.Batch.from_values(np.concatenate([circ.constants, circ.sigmas]), cc.RATE_BITS, cc.CAP_HEIGHT)
    digest = [int(x) for x in oracle.hash_no_pad(np.concatenate([cs.cap.reshape(-1), np.array([degree_bits], dtype=np.uint64)]))]
    pr = oracle.circuit_prove(circ, wires, pis, digest)
    op = pr["openings"]
    return {"shape": [degree_bits, seed, all_gates], "program_ops": len(circ.program.ops), "num_constants": circ.num_constants,
            "digest": [f"{x:016x}" for x in digest], "constants_sigmas_cap_sha256": sha(pr["constants_sigmas_cap"]),
            "wires_cap_sha256": sha(pr["wires_cap"]), "zs_partial_products_cap_sha256": sha(pr["plonk_zs_partial_products_cap"]),
            "quotient_polys_cap_sha256": sha(pr["quotient_polys_cap"]),
            "openings_sha256": sha(np.concatenate([np.asarray(op[k]).reshape(-1) for k in
                                                   ("constants_sigmas", "wires", "zs_partial_products", "quotient_polys", "plonk_zs_next")])),
            "opening_proof_sha256": sha(pr["opening_proof"]), "opening_proof_words": int(pr["opening_proof"].size),
            "pow_witness": f"{int(pr['opening_proof'][-1]):016x}"}


def generate_circuits():
    return {"note": "oracle.circuit_prove outputs on seeded synthetic circuits; regression anchors, NOT upstream vectors",
            "circuits": [circuit_case(*c) for c in CIRCUITS]}


def generate():
    return {"note": "oracle outputs on seeded inputs; regression anchors, NOT upstream vectors (see tools/gen_golden.py)",
            "commits": [commit_case(*c) for c in COMMITS], "proofs": [proof_case(*p) for p in PROOFS]}


if __name__ == "__main__":
    out = os.path.join(ROOT, "tests", "golden", "path_vectors.json")
    with open(out, "w") as f:
        json.dump(generate(), f, indent=1)
    print("wrote", out)
    out = os.path.join(ROOT, "tests", "golden", "circuit_vectors.json")
    with open(out, "w") as f:
        json.dump(generate_circuits(), f, indent=1)
    print("wrote", out)
