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
