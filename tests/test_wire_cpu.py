"""Proof wire formats (eth_tx_proof_b200/wire.py): the flat "B200STK2" layout <-> the serde-JSON form of starky 0.4.0's
StarkProofWithPublicInputs that the reference's leader / workers exchange (/root/reference/leader/src/main.rs:56-60)."""
import json

import numpy as np

import oracle
from eth_tx_proof_b200 import cprog, synthetic as syn, wire
from test_ctl_oracle import prove_all


def test_flat_to_serde_json_round_trip_fibonacci_and_memory():
    t, pi = syn.fibonacci_trace(6, seed=3)
    for table, trace, pis, n_lookup in ((oracle.TABLE_FIBONACCI, t, pi, 0), (oracle.TABLE_MEMORY, syn.memory_trace(7), (), 4)):
        proof = oracle.stark_prove(table, trace, pis)
        text = wire.to_serde_json(proof)
        d = json.loads(text)
        assert set(d) == {"proof", "public_inputs"}
        assert set(d["proof"]) == {"trace_cap", "auxiliary_polys_cap", "quotient_polys_cap", "openings", "opening_proof"}
        assert set(d["proof"]["openings"]) == {"local_values", "next_values", "auxiliary_polys", "auxiliary_polys_next", "ctl_zs_first",
                                               "quotient_polys"}
        fri = d["proof"]["opening_proof"]
        assert set(fri) == {"commit_phase_merkle_caps", "query_round_proofs", "final_poly", "pow_witness"}
        assert len(fri["query_round_proofs"]) == 84 and len(d["proof"]["trace_cap"]) == 16
        # FriQueryStep.evals of the uncompressed FriProof: all 2^arity_bits values (the verifier indexes evals[x & 15])
        assert all(len(s["evals"]) == 16 for r in fri["query_round_proofs"] for s in r["steps"])
        assert d["public_inputs"] == [int(x) for x in pis]
        back = wire.from_serde_json(text, table, int(trace.shape[1]).bit_length() - 1, n_lookup_cols=n_lookup)
        assert (back == proof).all()


def test_ctl_proof_round_trip():
    tables, _ = cprog.ctl_demo_tables(5, 4, 4)
    proofs, _ = prove_all(tables)
    p = wire.parse(proofs[0])
    h = p["header"]
    assert h["n_ctl_zs"] == 2 and len(p["proof"]["openings"]["ctl_zs_first"]) == 2
    back = wire.from_serde_json(wire.to_serde_json(proofs[0]), h["table"], h["degree_bits"], n_lookup_cols=h["n_lookup_cols"],
                                n_ctl_helper_cols=h["n_ctl_helper_cols"])
    assert (back == proofs[0]).all()


def test_circuit_proof_serde_json_round_trip():
    """A "B200PLK1" circuit proof <-> serde-JSON ProofWithPublicInputs (plonky2 0.2.2 plonk/proof.rs field names): the flat
    words survive the round trip; built here from the oracle's circuit proof (no GPU)."""
    import json

    import numpy as np

    import oracle
    from eth_tx_proof_b200 import circuit as cc, wire

    circuit, wires, public_inputs = cc.hash_chain_circuit(5, seed=4)
    pr = oracle.circuit_prove(circuit, wires, public_inputs, [1, 2, 3, 4])
    op = pr["openings"]
    nc = circuit.num_constants
    n_fri = 0  # 2^5 rows: ConstantArityBits(4, 5) stops at degree_bits <= 5
    hdr = np.zeros(wire.HEADER_WORDS, dtype=np.uint64)
    body = np.concatenate([np.asarray(x, dtype=np.uint64).reshape(-1) for x in (
        pr["wires_cap"], pr["plonk_zs_partial_products_cap"], pr["quotient_polys_cap"], op["constants_sigmas"], op["wires"],
        op["zs_partial_products"][:2], op["plonk_zs_next"], op["zs_partial_products"][2:], op["quotient_polys"], pr["opening_proof"],
        cc.hash_no_pad(public_inputs))])
    hdr[:16] = [wire.CIRCUIT_MAGIC, 5, nc, 80, 135, 2, 9, 8, 3, 4, n_fri, 4, 32, 28, 16, wire.HEADER_WORDS + body.size]
    words = np.concatenate([hdr, body])
    parsed = wire.parse_circuit_proof(words)
    assert (parsed["openings"]["plonk_sigmas"] == np.asarray(op["constants_sigmas"]).reshape(-1, 2)[nc:]).all()
    text = wire.circuit_to_serde_json(words, public_inputs)
    d = json.loads(text)
    assert set(d) == {"proof", "public_inputs"}
    assert set(d["proof"]) == {"wires_cap", "plonk_zs_partial_products_cap", "quotient_polys_cap", "openings", "opening_proof"}
    assert set(d["proof"]["openings"]) == {"constants", "plonk_sigmas", "wires", "plonk_zs", "plonk_zs_next", "partial_products",
                                           "quotient_polys", "lookup_zs", "lookup_zs_next"}
    assert len(d["proof"]["opening_proof"]["query_round_proofs"]) == 28
    assert len(d["proof"]["opening_proof"]["query_round_proofs"][0]["initial_trees_proof"]["evals_proofs"]) == 4
    back = wire.circuit_from_serde_json(text, 5, cc.hash_no_pad(public_inputs))
    assert back.shape == words.shape and (back == words).all()
