"""Verifier for the circuit proofs of eth_tx_proof_b200/circuit.py (test infrastructure): an independent Python
restatement of the VERIFIER side of plonky2 0.2.2 —
    src/plonk/verifier.rs       verify_with_challenges: vanishing polynomial at zeta == Z_H(zeta) * reduced quotient chunks
    src/plonk/get_challenges.rs get_challenges: circuit digest, public-input hash, wires cap -> betas, gammas; Z cap -> alphas;
                                quotient cap -> zeta; openings -> FRI challenges
    src/plonk/vanishing_poly.rs eval_vanishing_poly (through the circuit's recorded program, interpreted over F_{p^2})
    src/fri/verifier.rs         verify_fri_proof for a general FriInstanceInfo (four oracles, two batches)
(/root/reference/Cargo.lock:3441; the reference proves at /root/reference/ops/src/lib.rs:52,72,95).  Reuses the field /
hash / Merkle / interpolation helpers of stark_verifier.py."""
from __future__ import annotations

import numpy as np

import stark_verifier as V
from oracle import pyref as R

P = R.P
VerifyError = V.VerifyError


def parse_fri_proof(words, oracle_cols, degree_bits, rate_bits, cap_height, n_layers, arity_bits, num_queries, final_len):
    """Flat FriProof (etp_fri_proof_words layout) -> dict(fri_caps, queries, final_poly, pow_witness)."""
    w = [int(x) for x in words]
    pos = 0
    capw = 4 << cap_height

    def take(n):
        nonlocal pos
        out = w[pos:pos + n]
        if len(out) != n:
            raise VerifyError("truncated FRI proof")
        pos += n
        return out

    def take_cap():
        c = take(capw)
        return [c[4 * i:4 * i + 4] for i in range(capw // 4)]

    def take_ext(n):
        c = take(2 * n)
        return [(c[2 * i], c[2 * i + 1]) for i in range(n)]

    log_lde = degree_bits + rate_bits
    out = {"fri_caps": [take_cap() for _ in range(n_layers)], "queries": []}
    for _ in range(num_queries):
        q = dict(initial=[], steps=[])
        for ncols in oracle_cols:
            leaf = take(ncols)
            sib = take(4 * (log_lde - cap_height))
            q["initial"].append((leaf, [sib[4 * i:4 * i + 4] for i in range(len(sib) // 4)]))
        bits = log_lde
        for _l in range(n_layers):
            bits -= arity_bits
            ev = take_ext(1 << arity_bits)
            sib = take(4 * (bits - cap_height))
            q["steps"].append((ev, [sib[4 * i:4 * i + 4] for i in range(len(sib) // 4)]))
        out["queries"].append(q)
    out["final_poly"] = take_ext(final_len)
    out["pow_witness"] = take(1)[0]
    if pos != len(w):
        raise VerifyError("trailing words in the FRI proof")
    return out


def verify_fri(fri, caps, batches, ch, degree_bits, rate_bits, arity_bits, pow_bits, num_queries, fast=True, max_queries=None, folds=None, merkle=None):
    """verify_fri_proof.  batches: [(point (ext), [(oracle, poly)], [opened values (ext)])]; `ch`: the challenger after it
    observed the openings.  folds: a list that receives every compute_evaluation instance
    (values in natural order, coset start, beta, interpolated value) — inputs for the in-circuit fold check; merkle: a list that
    receives every Merkle check (leaf, index, siblings, cap) — inputs for the in-circuit Merkle verifications."""
    hash_or_noop, two_to_one, _ = V._hashers(fast)
    fri_alpha = ch.get_ext()
    betas = []
    for cap in fri["fri_caps"]:
        for d in cap:
            ch.observe(d)
        betas.append(ch.get_ext())
    for e in fri["final_poly"]:
        ch.observe(e)
    ch.observe([fri["pow_witness"]])
    pow_response = ch.get()
    if pow_bits and (pow_response >> (64 - pow_bits)) != 0:
        raise VerifyError("Invalid proof of work witness")
    lde_bits = degree_bits + rate_bits
    query_indices = [ch.get() % (1 << lde_bits) for _ in range(num_queries)]
    reduced = [V._reduce(fri_alpha, vals) for _, _, vals in batches]
    w_lde = R.root_of_unity(lde_bits)
    arity = 1 << arity_bits
    nq = num_queries if max_queries is None else min(max_queries, num_queries)
    for qi in range(nq):
        x_index = query_indices[qi]
        q = fri["queries"][qi]
        for (leaf, sib), cap in zip(q["initial"], caps):
            V._merkle_verify(leaf, x_index, sib, cap, hash_or_noop, two_to_one)
            if merkle is not None:
                merkle.append((leaf, x_index, sib, cap))
        subgroup_x = R.GENERATOR * pow(w_lde, R.bitrev(x_index, lde_bits), P) % P
        leaves = [lf for lf, _ in q["initial"]]
        sx = R.e_from(subgroup_x)
        total = (0, 0)
        for (point, polys, _), red in zip(batches, reduced):  # fri_combine_initial
            evals = [R.e_from(leaves[o][k]) for o, k in polys]
            numerator = R.e_sub(V._reduce(fri_alpha, evals), red)
            denominator = R.e_sub(sx, point)
            total = R.e_mul(total, R.e_pow(fri_alpha, len(evals)))
            total = R.e_add(total, R.e_mul(numerator, R.e_inv(denominator)))
        old_eval = total
        for i, (evals, sib) in enumerate(q["steps"]):
            coset_index = x_index >> arity_bits
            within = x_index & (arity - 1)
            if evals[within] != old_eval:
                raise VerifyError(f"FRI consistency check failed (query {qi}, layer {i})")
            ga = R.root_of_unity(arity_bits)
            ev = [evals[R.bitrev(k, arity_bits)] for k in range(arity)]
            rev_within = R.bitrev(within, arity_bits)
            coset_start = subgroup_x * pow(ga, arity - rev_within, P) % P
            pts = [(R.e_from(coset_start * pow(ga, k, P) % P), ev[k]) for k in range(arity)]
            old_eval = V._interpolate(pts, betas[i])
            if folds is not None:
                folds.append((ev, coset_start, betas[i], old_eval))
            flat = [c for e in evals for c in e]
            V._merkle_verify(flat, coset_index, sib, fri["fri_caps"][i], hash_or_noop, two_to_one)
            if merkle is not None:
                merkle.append((flat, coset_index, sib, fri["fri_caps"][i]))
            subgroup_x = pow(subgroup_x, arity, P)
            x_index = coset_index
        acc = (0, 0)
        for c in reversed(fri["final_poly"]):
            acc = R.e_add(R.e_mul(acc, R.e_from(subgroup_x)), c)
        if acc != old_eval:
            raise VerifyError("Final polynomial evaluation is invalid.")


def verify(proof, circuit, constants_sigmas_cap, circuit_digest, fast=True, max_queries=None, folds=None, merkle=None):
    """Raises VerifyError unless `proof` (CircuitProver.prove) is a valid proof for `circuit`."""
    from eth_tx_proof_b200 import circuit as cc

    db = circuit.degree_bits
    n = 1 << db
    _, _, perm = V._hashers(fast)
    ext = lambda a: [(int(x[0]), int(x[1])) for x in np.asarray(a).reshape(-1, 2)]
    caps = lambda c: [[int(v) for v in row] for row in np.asarray(c).reshape(-1, 4)]
    op = proof["openings"]
    cs, wires, zs_pp, quot, zs_next = (ext(op[k]) for k in ("constants_sigmas", "wires", "zs_partial_products", "quotient_polys", "plonk_zs_next"))
    shapes = [circuit.num_constants + cc.NUM_ROUTED, cc.NUM_WIRES, cc.NUM_CHALLENGES * (1 + cc.NUM_PARTIAL_PRODUCTS),
              cc.NUM_CHALLENGES * cc.QUOTIENT_DEGREE_FACTOR]
    if [len(cs), len(wires), len(zs_pp), len(quot), len(zs_next)] != shapes + [cc.NUM_CHALLENGES]:
        raise VerifyError("shape")
    # ---- get_challenges
    pi_hash = cc.hash_no_pad(proof["public_inputs"])
    ch = V._Challenger(perm)
    ch.observe([int(x) for x in circuit_digest])
    ch.observe(pi_hash)
    for d in caps(proof["wires_cap"]):
        ch.observe(d)
    betas, gammas = ch.get_n(cc.NUM_CHALLENGES), ch.get_n(cc.NUM_CHALLENGES)
    for d in caps(proof["plonk_zs_partial_products_cap"]):
        ch.observe(d)
    alphas = ch.get_n(cc.NUM_CHALLENGES)
    for d in caps(proof["quotient_polys_cap"]):
        ch.observe(d)
    zeta = ch.get_ext()
    for batch in (cs, wires, zs_pp, quot, zs_next):
        for e in batch:
            ch.observe(e)
    # ---- vanishing polynomial at zeta (the recorded program over F_{p^2}): lv = openings, nv = openings at g*zeta (only Z is read)
    lv = cs + wires + zs_pp + [zeta]
    nv = [None] * len(lv)
    for i in range(cc.NUM_CHALLENGES):
        nv[circuit.col_z(i)] = zs_next[i]
    zeta_pow = R.e_pow(zeta, n)
    z_h = R.e_sub(zeta_pow, (1, 0))
    l_0 = R.e_mul(z_h, R.e_inv(R.e_scalar(R.e_sub(zeta, (1, 0)), n)))
    cons = V._Consumer(alphas, None, l_0, None)
    V._eval_program(circuit.program, lv, nv, (), (), pi_hash, list(betas) + list(gammas), cons)
    f = cc.QUOTIENT_DEGREE_FACTOR
    for i in range(cc.NUM_CHALLENGES):
        if cons.acc[i] != R.e_mul(z_h, V._reduce(zeta_pow, quot[i * f:(i + 1) * f])):
            raise VerifyError("Mismatch between evaluation and opening of quotient polynomial")
    # ---- FRI over the four oracles
    g = R.root_of_unity(db)
    zeta_next = R.e_scalar(zeta, g)
    all_polys = [(o, k) for o, cnt in enumerate(shapes) for k in range(cnt)]
    batches = [(zeta, all_polys, cs + wires + zs_pp + quot), (zeta_next, [(2, k) for k in range(cc.NUM_CHALLENGES)], zs_next)]
    n_layers, bits = 0, db
    while bits > 5:  # ConstantArityBits(4, 5)
        n_layers, bits = n_layers + 1, bits - 4
    fri = parse_fri_proof(proof["opening_proof"], shapes, db, cc.RATE_BITS, cc.CAP_HEIGHT, n_layers, 4, cc.NUM_QUERIES, 1 << bits)
    all_caps = [caps(constants_sigmas_cap), caps(proof["wires_cap"]), caps(proof["plonk_zs_partial_products_cap"]), caps(proof["quotient_polys_cap"])]
    verify_fri(fri, all_caps, batches, ch, db, cc.RATE_BITS, 4, cc.POW_BITS, cc.NUM_QUERIES, fast, max_queries, folds, merkle)
    return True
