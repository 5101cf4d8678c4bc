#!/usr/bin/env python
"""Diffs the output of rust/parity_dump (REAL plonky2 0.2.2 / starky 0.4.0 on the seeds of tools/gen_golden.py) against
tests/golden/path_vectors.json (this repository's oracle == its CUDA library, tests/test_golden_vectors.py).

    RAYON_NUM_THREADS=1 cargo run --release --manifest-path rust/parity_dump/Cargo.toml > upstream_vectors.json
    python tools/compare_parity_dump.py upstream_vectors.json

Exit status 0 = every compared entry is byte-identical: "parity unpinned"  is then closed for the
commit path (cap, coefficients, digests in plonky2's layout, leaf rows, Merkle paths) and for starky::prove on the
Fibonacci table (whole proof incl. transcript order, FRI schedule and the smallest PoW witness)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def compare(upstream: dict, golden: dict):
    """-> list of (case, field, upstream value, golden value) mismatches; entries are matched by shape / (table, log_n, seed)."""
    bad = []
    g_commits = {tuple(c["shape"]): c for c in golden["commits"]}
    for c in upstream.get("commits", []):
        g = g_commits.get(tuple(c["shape"]))
        if g is None:
            bad.append((tuple(c["shape"]), "missing in golden", None, None))
            continue
        for k in ("cap_sha256", "cap_first", "coeffs_sha256", "digests_sha256", "leaf_rows", "leaf_rows_sha256", "paths_sha256"):
            if c.get(k) != g.get(k):
                bad.append((tuple(c["shape"]), k, c.get(k), g.get(k)))
    g_proofs = {(p["table"], p["log_n"], p["seed"]): p for p in golden["proofs"]}
    for p in upstream.get("proofs", []):
        key = (p["table"], p["log_n"], p["seed"])
        g = g_proofs.get(key)
        if g is None:
            bad.append((key, "missing in golden", None, None))
            continue
        for k in ("words", "pow_witness", "proof_sha256_without_table_id"):
            if p.get(k) != g.get(k):
                bad.append((key, k, p.get(k), g.get(k)))
    return bad


def main():
    with open(sys.argv[1]) as f:
        upstream = json.load(f)
    with open(os.path.join(ROOT, "tests", "golden", "path_vectors.json")) as f:
        golden = json.load(f)
    bad = compare(upstream, golden)
    n = len(upstream.get("commits", [])) + len(upstream.get("proofs", []))
    for case, field, u, g in bad:
        print(f"MISMATCH {case} {field}: upstream {u} != repo {g}")
    print(f"{n} cases compared, {len(bad)} mismatching fields")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
