"""Proof wire formats.

* the flat u64 layout "B200STK2" the C ABI writes: `parse` turns it into plonky2's field structure;
* `to_serde_json` / `from_serde_json`: the JSON that serde derives for starky 0.4.0's
  `StarkProofWithPublicInputs<GoldilocksField, PoseidonGoldilocksConfig, 2>` — the form in which the reference's leader
  and workers exchange proofs (`serde_json`, /root/reference/leader/src/main.rs:56-60; proofs travel through Paladin as
  serde values, /root/reference/ops/src/lib.rs:46-52).  Field names and nesting follow the upstream structs:
    StarkProofWithPublicInputs { proof: StarkProof, public_inputs: Vec<F> }
    StarkProof { trace_cap, auxiliary_polys_cap: Option, quotient_polys_cap: Option, openings: StarkOpeningSet, opening_proof: FriProof }
    StarkOpeningSet { local_values, next_values, auxiliary_polys: Option, auxiliary_polys_next: Option, ctl_zs_first: Option, quotient_polys: Option }
    FriProof { commit_phase_merkle_caps, query_round_proofs: [FriQueryRound], final_poly: PolynomialCoeffs { coeffs }, pow_witness }
    FriQueryRound { initial_trees_proof: FriInitialTreeProof { evals_proofs: [(Vec<F>, MerkleProof { siblings })] }, steps: [FriQueryStep { evals, merkle_proof }] }
  with GoldilocksField -> u64 number, QuadraticExtension -> [c0, c1], HashOut -> {"elements": [4]}, MerkleCap -> [HashOut].
Host side only (numpy / json): nothing here touches the GPU.
"""
from __future__ import annotations

import json
from typing import Any, Dict

import numpy as np

MAGIC = 0x4232303053544B32  # "B200STK2"
HEADER_WORDS = 24
HEADER_FIELDS = ("magic", "table", "degree_bits", "n_trace", "n_aux", "n_quot", "cap_height", "n_fri_layers", "arity_bits", "final_poly_len",
                 "num_queries", "n_public_inputs", "rate_bits", "pow_bits", "num_challenges", "total_words", "n_ctl_zs", "n_lookup_cols",
                 "n_ctl_helper_cols")


def parse(words) -> Dict[str, Any]:
    """Flat proof words -> {"header": {...}, "proof": StarkProof-shaped dict of Python ints, "public_inputs": [...]}."""
    w = [int(x) for x in np.asarray(words, dtype=np.uint64)]
    if len(w) < HEADER_WORDS or w[0] != MAGIC:
        raise ValueError("not a B200STK2 proof")
    h = dict(zip(HEADER_FIELDS, w[:len(HEADER_FIELDS)]))
    if h["total_words"] != len(w):
        raise ValueError("length mismatch")
    pos = HEADER_WORDS
    capw = 4 << h["cap_height"]

    def take(n):
        nonlocal pos
        out = w[pos:pos + n]
        if len(out) != n:
            raise ValueError("truncated proof")
        pos += n
        return out

    def hashes(ws):
        return [{"elements": ws[4 * i:4 * i + 4]} for i in range(len(ws) // 4)]

    def cap():
        return hashes(take(capw))

    def exts(n):
        c = take(2 * n)
        return [[c[2 * i], c[2 * i + 1]] for i in range(n)]

    trace_cap = cap()
    aux_cap = cap() if h["n_aux"] else None
    quot_cap = cap()
    openings = {"local_values": exts(h["n_trace"]), "next_values": exts(h["n_trace"]),
                "auxiliary_polys": exts(h["n_aux"]) if h["n_aux"] else None,
                "auxiliary_polys_next": exts(h["n_aux"]) if h["n_aux"] else None}
    openings["ctl_zs_first"] = take(h["n_ctl_zs"]) if h["n_ctl_zs"] else None
    openings["quotient_polys"] = exts(h["n_quot"])
    caps = [cap() for _ in range(h["n_fri_layers"])]
    log_lde = h["degree_bits"] + h["rate_bits"]
    rounds = []
    for _ in range(h["num_queries"]):
        evals_proofs = []
        for ncols in (h["n_trace"], h["n_aux"], h["n_quot"]):
            if ncols == 0:
                continue
            leaf = take(ncols)
            evals_proofs.append([leaf, {"siblings": hashes(take(4 * (log_lde - h["cap_height"])))}])
        steps = []
        bits = log_lde
        for _l in range(h["n_fri_layers"]):
            bits -= h["arity_bits"]
            ev = exts(1 << h["arity_bits"])
            steps.append({"evals": ev, "merkle_proof": {"siblings": hashes(take(4 * (bits - h["cap_height"])))}})
        rounds.append({"initial_trees_proof": {"evals_proofs": evals_proofs}, "steps": steps})
    final_poly = {"coeffs": exts(h["final_poly_len"])}
    pow_witness = take(1)[0]
    public_inputs = take(h["n_public_inputs"])
    if pos != len(w):
        raise ValueError("trailing words")
    proof = {"trace_cap": trace_cap, "auxiliary_polys_cap": aux_cap, "quotient_polys_cap": quot_cap, "openings": openings,
             "opening_proof": {"commit_phase_merkle_caps": caps, "query_round_proofs": rounds, "final_poly": final_poly,
                               "pow_witness": pow_witness}}
    return {"header": h, "proof": proof, "public_inputs": public_inputs}


def to_serde_json(words) -> str:
    """serde_json::to_string(&StarkProofWithPublicInputs) for the proof in `words`."""
    p = parse(words)
    return json.dumps({"proof": p["proof"], "public_inputs": p["public_inputs"]}, separators=(",", ":"))


def from_serde_json(text: str, table: int, degree_bits: int, rate_bits: int = 1, pow_bits: int = 16, num_challenges: int = 2,
                    n_lookup_cols: int = 0, n_ctl_helper_cols: int = 0) -> np.ndarray:
    """The inverse: a serde-JSON StarkProofWithPublicInputs -> flat proof words (header rebuilt from the structure plus the
    few facts JSON does not carry: table id, degree_bits, rate/pow bits and the split of the auxiliary columns)."""
    d = json.loads(text)
    pr, pi = d["proof"], d["public_inputs"]
    op, fri = pr["openings"], pr["opening_proof"]
    flat_cap = lambda c: [x for hsh in c for x in hsh["elements"]]
    flat_ext = lambda v: [x for e in (v or []) for x in e]
    n_trace = len(op["local_values"])
    n_aux = len(op["auxiliary_polys"] or [])
    n_quot = len(op["quotient_polys"] or [])
    n_zs = len(op["ctl_zs_first"] or [])
    cap_height = (len(pr["trace_cap"]) - 1).bit_length()
    steps0 = fri["query_round_proofs"][0]["steps"] if fri["query_round_proofs"] else []
    arity_bits = (len(steps0[0]["evals"]) - 1).bit_length() if steps0 else 4
    body = flat_cap(pr["trace_cap"])
    if pr["auxiliary_polys_cap"] is not None:
        body += flat_cap(pr["auxiliary_polys_cap"])
    body += flat_cap(pr["quotient_polys_cap"])
    body += flat_ext(op["local_values"]) + flat_ext(op["next_values"]) + flat_ext(op["auxiliary_polys"]) + flat_ext(op["auxiliary_polys_next"])
    body += list(op["ctl_zs_first"] or []) + flat_ext(op["quotient_polys"])
    for c in fri["commit_phase_merkle_caps"]:
        body += flat_cap(c)
    for r in fri["query_round_proofs"]:
        for leaf, mp in r["initial_trees_proof"]["evals_proofs"]:
            body += list(leaf) + flat_cap(mp["siblings"])
        for st in r["steps"]:
            body += flat_ext(st["evals"]) + flat_cap(st["merkle_proof"]["siblings"])
    body += flat_ext(fri["final_poly"]["coeffs"]) + [fri["pow_witness"]] + list(pi)
    hdr = [0] * HEADER_WORDS
    hdr[:19] = [MAGIC, table, degree_bits, n_trace, n_aux, n_quot, cap_height, len(fri["commit_phase_merkle_caps"]), arity_bits,
                len(fri["final_poly"]["coeffs"]), len(fri["query_round_proofs"]), len(pi), rate_bits, pow_bits, num_challenges,
                HEADER_WORDS + len(body), n_zs, n_lookup_cols, n_ctl_helper_cols]
    return np.array(hdr + body, dtype=np.uint64)


# ---- circuit proofs ("B200PLK1", include/etp_b200.h etp_circuit_prove_*) ------------------------------------------------------
CIRCUIT_MAGIC = 0x42323030504C4B31  # "B200PLK1"
CIRCUIT_HEADER_FIELDS = ("magic", "degree_bits", "num_constants", "num_routed_wires", "num_wires", "num_challenges", "num_partial_products",
                         "quotient_degree_factor", "rate_bits", "cap_height", "n_fri_layers", "arity_bits", "final_poly_len", "num_queries",
                         "pow_bits", "total_words")


def parse_circuit_proof(words) -> Dict[str, Any]:
    """Flat circuit proof -> the field structure of plonky2's ProofWithPublicInputs as numpy arrays:
    {"header", "wires_cap", "plonk_zs_partial_products_cap", "quotient_polys_cap" ((2^cap, 4)),
     "openings": {"constants", "plonk_sigmas", "wires", "plonk_zs", "plonk_zs_next", "partial_products", "quotient_polys"} ((k, 2)),
     "opening_proof" (flat FriProof words), "public_inputs_hash"}."""
    w = np.asarray(words, dtype=np.uint64)
    if w.size < HEADER_WORDS or int(w[0]) != CIRCUIT_MAGIC:
        raise ValueError("not a B200PLK1 proof")
    h = {k: int(v) for k, v in zip(CIRCUIT_HEADER_FIELDS, w[:len(CIRCUIT_HEADER_FIELDS)])}
    if h["total_words"] != w.size:
        raise ValueError("length mismatch")
    pos = HEADER_WORDS
    capw = 4 << h["cap_height"]

    def take(n):
        nonlocal pos
        if pos + n > w.size:
            raise ValueError("truncated proof")
        out = w[pos:pos + n]
        pos += n
        return out

    out = {"header": h}
    for k in ("wires_cap", "plonk_zs_partial_products_cap", "quotient_polys_cap"):
        out[k] = take(capw).reshape(-1, 4).copy()
    K = h["num_challenges"]
    counts = (("constants", h["num_constants"]), ("plonk_sigmas", h["num_routed_wires"]), ("wires", h["num_wires"]), ("plonk_zs", K),
              ("plonk_zs_next", K), ("partial_products", K * h["num_partial_products"]), ("quotient_polys", K * h["quotient_degree_factor"]))
    out["openings"] = {k: take(2 * c).reshape(-1, 2).copy() for k, c in counts}
    out["opening_proof"] = take(w.size - pos - 4).copy()
    out["public_inputs_hash"] = take(4).copy()
    return out


def _fri_to_serde(words, oracle_cols, h) -> Dict[str, Any]:
    """Flat FriProof words -> plonky2's FriProof structure (as `parse` builds it for the STARK proofs)."""
    w = [int(x) for x in words]
    pos = 0
    capw = 4 << h["cap_height"]

    def take(n):
        nonlocal pos
        out = w[pos:pos + n]
        if len(out) != n:
            raise ValueError("truncated FRI proof")
        pos += n
        return out

    hashes = lambda ws: [{"elements": ws[4 * i:4 * i + 4]} for i in range(len(ws) // 4)]
    exts = lambda n: (lambda c: [[c[2 * i], c[2 * i + 1]] for i in range(n)])(take(2 * n))
    log_lde = h["degree_bits"] + h["rate_bits"]
    caps = [hashes(take(capw)) for _ in range(h["n_fri_layers"])]
    rounds = []
    for _ in range(h["num_queries"]):
        evals_proofs = []
        for ncols in oracle_cols:
            leaf = take(ncols)
            evals_proofs.append([leaf, {"siblings": hashes(take(4 * (log_lde - h["cap_height"])))}])
        steps, bits = [], log_lde
        for _l in range(h["n_fri_layers"]):
            bits -= h["arity_bits"]
            ev = exts(1 << h["arity_bits"])
            steps.append({"evals": ev, "merkle_proof": {"siblings": hashes(take(4 * (bits - h["cap_height"])))}})
        rounds.append({"initial_trees_proof": {"evals_proofs": evals_proofs}, "steps": steps})
    final_poly = {"coeffs": exts(h["final_poly_len"])}
    pow_witness = take(1)[0]
    if pos != len(w):
        raise ValueError("trailing words in the FRI proof")
    return {"commit_phase_merkle_caps": caps, "query_round_proofs": rounds, "final_poly": final_poly, "pow_witness": pow_witness}


def circuit_to_serde_json(words, public_inputs) -> str:
    """serde_json::to_string(&ProofWithPublicInputs<GoldilocksField, PoseidonGoldilocksConfig, 2>) for a circuit proof — the
    form in which the reference passes shrink / aggregation / block proofs around (PlonkyProofIntern,
    /root/reference/ops/src/lib.rs:63-101).  Field names follow plonky2 0.2.2 plonk/proof.rs:
      ProofWithPublicInputs { proof: Proof, public_inputs }
      Proof { wires_cap, plonk_zs_partial_products_cap, quotient_polys_cap, openings: OpeningSet, opening_proof: FriProof }
      OpeningSet { constants, plonk_sigmas, wires, plonk_zs, plonk_zs_next, partial_products, quotient_polys,
                   lookup_zs, lookup_zs_next }"""
    p = parse_circuit_proof(words)
    h = p["header"]
    cap = lambda c: [{"elements": [int(x) for x in row]} for row in c]
    ext = lambda a: [[int(x[0]), int(x[1])] for x in a]
    op = {k: ext(v) for k, v in p["openings"].items()}
    op["lookup_zs"], op["lookup_zs_next"] = [], []
    K = h["num_challenges"]
    oracle_cols = [h["num_constants"] + h["num_routed_wires"], h["num_wires"], K * (1 + h["num_partial_products"]), K * h["quotient_degree_factor"]]
    proof = {"wires_cap": cap(p["wires_cap"]), "plonk_zs_partial_products_cap": cap(p["plonk_zs_partial_products_cap"]),
             "quotient_polys_cap": cap(p["quotient_polys_cap"]), "openings": op, "opening_proof": _fri_to_serde(p["opening_proof"], oracle_cols, h)}
    return json.dumps({"proof": proof, "public_inputs": [int(x) for x in public_inputs]}, separators=(",", ":"))


def circuit_from_serde_json(text: str, degree_bits: int, public_inputs_hash, rate_bits: int = 3, pow_bits: int = 16) -> np.ndarray:
    """The inverse of circuit_to_serde_json: flat "B200PLK1" words (the header facts JSON does not carry are arguments)."""
    d = json.loads(text)
    pr = d["proof"]
    op, fri = pr["openings"], pr["opening_proof"]
    flat_cap = lambda c: [x for hsh in c for x in hsh["elements"]]
    flat_ext = lambda v: [x for e in v for x in e]
    K = len(op["plonk_zs"])
    cap_height = (len(pr["wires_cap"]) - 1).bit_length()
    steps0 = fri["query_round_proofs"][0]["steps"] if fri["query_round_proofs"] else []
    arity_bits = (len(steps0[0]["evals"]) - 1).bit_length() if steps0 else 4
    body = flat_cap(pr["wires_cap"]) + flat_cap(pr["plonk_zs_partial_products_cap"]) + flat_cap(pr["quotient_polys_cap"])
    for k in ("constants", "plonk_sigmas", "wires", "plonk_zs", "plonk_zs_next", "partial_products", "quotient_polys"):
        body += flat_ext(op[k])
    for c in fri["commit_phase_merkle_caps"]:
        body += flat_cap(c)
    for r in fri["query_round_proofs"]:
        for leaf, mp in r["initial_trees_proof"]["evals_proofs"]:
            body += list(leaf) + flat_cap(mp["siblings"])
        for st in r["steps"]:
            body += flat_ext(st["evals"]) + flat_cap(st["merkle_proof"]["siblings"])
    body += flat_ext(fri["final_poly"]["coeffs"]) + [fri["pow_witness"]] + [int(x) for x in public_inputs_hash]
    hdr = [0] * HEADER_WORDS
    hdr[:16] = [CIRCUIT_MAGIC, degree_bits, len(op["constants"]), len(op["plonk_sigmas"]), len(op["wires"]), K, len(op["partial_products"]) // K,
                len(op["quotient_polys"]) // K, rate_bits, cap_height, len(fri["commit_phase_merkle_caps"]), arity_bits,
                len(fri["final_poly"]["coeffs"]), len(fri["query_round_proofs"]), pow_bits, HEADER_WORDS + len(body)]
    return np.array(hdr + body, dtype=np.uint64)
