"""CPU tests of the circuit-prover host logic (eth_tx_proof_b200/circuit.py; plonky2 0.2.2 plonk/{prover,vanishing_poly}.rs,
gates/*.rs, gates/selectors.rs — /root/reference/Cargo.lock:3441, reached from /root/reference/ops/src/lib.rs:52,72,95):
gates, selector groups, copy constraints -> sigmas, and the vanishing polynomial recorded as ONE constraint program, which
must vanish on every row of a valid witness (with the oracle's Z / partial products) and must not on a corrupted one."""
import numpy as np
import pytest

P = 0xFFFFFFFF00000001


def _eval_rows(circuit, wires, zs_pp, pi_hash, betas, gammas):
    """The program on every trace-domain row at once (numpy object arrays): [(kind, values[n])]."""
    t = circuit.virtual_trace(wires, zs_pp).astype(object)
    lv = [t[c] for c in range(t.shape[0])]
    nv = [np.roll(t[c], -1) for c in range(t.shape[0])]
    return circuit.program.evaluate(lv, nv, pi=pi_hash, ch=[int(x) for x in betas] + [int(x) for x in gammas])


def _violations(circuit, wires, zs_pp, pi_hash, betas, gammas):
    from eth_tx_proof_b200 import cprog

    bad = []
    for idx, (kind, vals) in enumerate(_eval_rows(circuit, wires, zs_pp, pi_hash, betas, gammas)):
        vals = np.asarray(vals, dtype=object) % P if not np.isscalar(vals) else np.array([vals % P] * circuit.n, dtype=object)
        rows = [0] if kind == cprog.EMIT_FIRST_ROW else range(circuit.n)
        bad += [(idx, r) for r in rows if vals[r] != 0]
    return bad


def test_poseidon_gate_witness_matches_the_permutation():
    import ctypes as C

    from eth_tx_proof_b200 import circuit as cc
    from eth_tx_proof_b200 import load_library

    L = load_library()
    rng = np.random.default_rng(3)
    for swap in (0, 1):
        ins = [int(x) % P for x in rng.integers(0, 2**63, 12)]
        w = cc.poseidon_gate_wires(ins, swap)
        st = list(ins)
        if swap:
            st[0:4], st[4:8] = ins[4:8], ins[0:4]
        a = (C.c_uint64 * 12)(*st)
        L.etp_host_poseidon_permute(a)
        assert w[12:24] == [int(x) for x in a]
    assert cc.hash_no_pad(list(range(8))) == cc.poseidon_gate_wires(list(range(8)) + [0] * 4, 0)[12:16]
    # the library's witness helper (etp_host_poseidon_gate_wires) == the gate's definition in Python integers, edge lanes included
    for k in range(40):
        ins = [int(x) % P for x in rng.integers(0, 2**64, 12, dtype=np.uint64)]
        ins[k % 12] = (0, P - 1, ins[(k + 4) % 12])[k % 3]
        for swap in (0, 1):
            assert cc.poseidon_gate_wires(ins, swap) == cc.poseidon_gate_wires_py(ins, swap)


def test_selector_groups_follow_the_degree_rule():
    from eth_tx_proof_b200 import circuit as cc

    gates = sorted([cc.NoopGate(), cc.ConstantGate(2), cc.PublicInputGate(), cc.ArithmeticGate(20), cc.PoseidonGate()], key=lambda g: (g.degree, g.id()))
    groups = cc.selector_groups(gates, 8)
    assert [list(g) for g in groups] == [[0, 1, 2, 3], [4]]
    assert [list(g) for g in cc.selector_groups(gates[:4], 8)] == [[0, 1, 2, 3]]  # 3 + 4 - 1 <= 8: one selector
    for grp in groups:  # filter degree + gate degree <= quotient degree factor
        assert max(len(grp) - 1 + (len(groups) > 1) + gates[i].degree for i in grp) <= 8


@pytest.mark.parametrize("degree_bits", [5, 6])
def test_vanishing_program_vanishes_on_a_valid_witness_only(degree_bits):
    import oracle
    from eth_tx_proof_b200 import circuit as cc

    circuit, wires, public_inputs = cc.hash_chain_circuit(degree_bits, seed=degree_bits)
    assert circuit.num_constants == 4 and circuit.num_selectors == 2 and circuit.num_gate_constraints == 123
    assert circuit.program.n_trace == 4 + 80 + 135 + 20 + 1 and circuit.program.degree == 9
    assert circuit.program.n_constraints == 2 + 20 + 123 == circuit.num_vanishing_terms
    pi_hash = cc.hash_no_pad(public_inputs)
    betas, gammas = [0x1234567, 0x7654321], [0xABCDEF, 0xFEDCBA]
    zs_pp = oracle.plonk_partial_products_and_zs(wires[:80], circuit.sigmas, circuit.k_is, 8, betas, gammas)
    assert (zs_pp[0] != 1).any() or True
    assert _violations(circuit, wires, zs_pp, pi_hash, betas, gammas) == []
    # sigma is a permutation of the identity values that respects the copy constraints
    ident = circuit.virtual_trace(wires, zs_pp)[-1][None, :].astype(object) * np.array(circuit.k_is, dtype=object)[:, None] % P
    assert sorted(ident.reshape(-1).tolist()) == sorted(circuit.sigmas.astype(object).reshape(-1).tolist())
    assert (circuit.sigmas.astype(object) != ident).sum() > 100
    # a corrupted S-box wire breaks a Poseidon constraint; a corrupted copy breaks the permutation argument
    bad = wires.copy()
    row = int(np.nonzero(circuit.gate_of_row == 4)[0][3])
    bad[cc.PoseidonGate.wire_partial_sbox(7), row] ^= np.uint64(1)
    assert _violations(circuit, bad, zs_pp, pi_hash, betas, gammas)
    bad = wires.copy()
    bad[0, row] = (int(bad[0, row]) + 1) % P  # an input wired to the previous row's output
    zs_bad = oracle.plonk_partial_products_and_zs(bad[:80], circuit.sigmas, circuit.k_is, 8, betas, gammas)
    v = _violations(circuit, bad, zs_bad, pi_hash, betas, gammas)
    assert v and any(idx >= 123 for idx, _ in v)  # a permutation / Z term (emitted after the reversed gate slots)
    # wrong public inputs: the PublicInputGate row fails
    assert _violations(circuit, wires, zs_pp, [1, 2, 3, 4], betas, gammas)


def test_vanishing_program_compiles_for_sm_100a():
    from eth_tx_proof_b200 import circuit as cc
    import eth_tx_proof_b200 as etp
    import ctypes as C

    circuit, _, _ = cc.hash_chain_circuit(5, seed=1)
    L = etp.load_library()
    w = np.ascontiguousarray(circuit.program.words)
    size = C.c_size_t()
    err = C.create_string_buffer(512)
    rc = L.etp_cprog_compile_check(w.ctypes.data_as(C.POINTER(C.c_uint64)), w.size, C.byref(size), err, 512)
    assert rc == 0, err.value
    assert size.value > 10000


@pytest.mark.parametrize("degree_bits", [5, 7])
def test_oracle_circuit_proof_is_accepted_and_tampering_is_not(degree_bits):
    """The CPU restatement of plonk::prover::prove (oracle.circuit_prove) against the independent Python verifier
    (tests/plonk_verifier.py): transcript, vanishing identity at zeta, FRI."""
    import oracle
    import plonk_verifier
    from eth_tx_proof_b200 import circuit as cc

    circuit, wires, public_inputs = cc.hash_chain_circuit(degree_bits, seed=2 + degree_bits)
    digest = [11, 22, 33, 44]
    proof = oracle.circuit_prove(circuit, wires, public_inputs, digest)
    plonk_verifier.verify(proof, circuit, proof["constants_sigmas_cap"], digest, max_queries=3)
    bad = dict(proof, public_inputs=[(public_inputs[0] + 1) % P] + list(public_inputs[1:]))
    with pytest.raises(plonk_verifier.VerifyError):
        plonk_verifier.verify(bad, circuit, proof["constants_sigmas_cap"], digest, max_queries=1)
    op = dict(proof["openings"])
    op["wires"] = op["wires"].copy()
    op["wires"][3, 0] ^= np.uint64(1)
    with pytest.raises(plonk_verifier.VerifyError):
        plonk_verifier.verify(dict(proof, openings=op), circuit, proof["constants_sigmas_cap"], digest, max_queries=1)
    fri = proof["opening_proof"].copy()
    fri[-3] ^= np.uint64(1)  # a final-polynomial coefficient
    with pytest.raises(plonk_verifier.VerifyError):
        plonk_verifier.verify(dict(proof, opening_proof=fri), circuit, proof["constants_sigmas_cap"], digest, max_queries=1)
    # a witness that violates a gate: the oracle's quotient is not a polynomial and the proof is rejected
    w = wires.copy()
    row = int(np.nonzero(circuit.gate_of_row == 3)[0][0])  # an ArithmeticGate row
    w[3, row] = (int(w[3, row]) + 1) % P
    with pytest.raises(plonk_verifier.VerifyError):
        plonk_verifier.verify(oracle.circuit_prove(circuit, w, public_inputs, digest), circuit, proof["constants_sigmas_cap"], digest, max_queries=1)


def test_full_gate_set_constraints_hold_and_every_gate_bites():
    """All fourteen gates (five selector groups): the vanishing program vanishes on the witness; corrupting one wire of a row of
    each gate type makes that gate's filtered constraints non-zero on that row."""
    import oracle
    from eth_tx_proof_b200 import circuit as cc

    circuit, wires, public_inputs = cc.hash_chain_circuit(6, seed=5, all_gates=True)
    assert [list(g) for g in circuit.groups] == [[0, 1, 2, 3, 4, 5], [6, 7, 8, 9], [10, 11], [12], [13]]
    assert circuit.num_constants == 7 and len(circuit.gates) == 14
    pi_hash = cc.hash_no_pad(public_inputs)
    betas, gammas = [3, 5], [7, 11]
    zs_pp = oracle.plonk_partial_products_and_zs(wires[:80], circuit.sigmas, circuit.k_is, 8, betas, gammas)
    assert _violations(circuit, wires, zs_pp, pi_hash, betas, gammas) == []
    # advice (non-routed or unconnected) wires: changing them cannot be absorbed by the permutation argument
    victims = {"ArithmeticExtensionGate": 15, "MulExtensionGate": 10, "BaseSumGate": 63, "ReducingExtensionGate": 80, "ReducingGate": 60,
               "ExponentiationGate": 100, "RandomAccessGate": 75, "PoseidonMdsGate": 30, "ConstantGate": 1, "ArithmeticGate": 7,
               "CosetInterpolationGate": 40}
    for gi, gate in enumerate(circuit.gates):
        key = gate.name.split(" ")[0]
        if key not in victims:
            continue
        row = int(np.nonzero(circuit.gate_of_row == gi)[0][0])
        bad = wires.copy()
        bad[victims[key], row] = (int(bad[victims[key], row]) + 1) % P
        zs_bad = oracle.plonk_partial_products_and_zs(bad[:80], circuit.sigmas, circuit.k_is, 8, betas, gammas)
        v = _violations(circuit, bad, zs_bad, pi_hash, betas, gammas)
        assert v and all(r == row or idx >= 123 for idx, r in v), key


def test_oracle_proof_of_the_full_gate_set_verifies():
    import oracle
    import plonk_verifier
    from eth_tx_proof_b200 import circuit as cc

    circuit, wires, public_inputs = cc.hash_chain_circuit(6, seed=8, all_gates=True)
    digest = [5, 6, 7, 8]
    proof = oracle.circuit_prove(circuit, wires, public_inputs, digest)
    plonk_verifier.verify(proof, circuit, proof["constants_sigmas_cap"], digest, max_queries=2)


def test_fast_partial_rounds_equal_the_plain_round_function():
    """The derived sparse-matrix form of the 22 partial rounds (poseidon.rs partial_first_constant_layer / mds_partial_layer_init /
    mds_partial_layer_fast) == the plain rounds, as a permutation and as POLYNOMIALS in the gate's wires: both forms of the
    PoseidonGate program give the same 123 constraint values on random (non-witness) wire assignments."""
    from eth_tx_proof_b200 import circuit as cc, cprog

    rc, _, _ = cc._pc()
    rng = np.random.default_rng(9)
    for _ in range(4):
        st = [int(x) % P for x in rng.integers(0, 2**63, 12)]
        plain = list(st)
        for r in range(22):
            plain = [(s + c) % P for s, c in zip(plain, rc[4 + r])]
            plain[0] = pow(plain[0], 7, P)
            plain = cc.mds_layer(plain)
        assert cc.partial_rounds_fast(list(st), cc._IntRing) == plain
    progs = []
    for fast in (True, False):
        gate = cc.PoseidonGate()
        gate.fast_partial = fast
        b = cprog.ProgramBuilder(cc.NUM_WIRES, 0, 8)
        for c in gate.eval(b, b.lv, None, None):
            b.constraint(c)
        progs.append(b.build())
    assert len(progs[0].ops) < 0.6 * len(progs[1].ops)
    for _ in range(3):
        lv = [int(x) % P for x in rng.integers(0, 2**63, cc.NUM_WIRES)]
        a, b_ = (p.evaluate(lv, lv) for p in progs)
        assert [v % P for _, v in a] == [v % P for _, v in b_]


def test_coset_interpolation_gate_computes_the_lagrange_interpolant():
    """The gate's evaluation_value == the value at evaluation_point of the degree-15 polynomial through the 16 extension
    values on shift * H (independent O(n^2) Lagrange formula), for the degree with_max_degree(4, 8) picks (6)."""
    from eth_tx_proof_b200 import circuit as cc

    g = cc.CosetInterpolationGate.with_max_degree(4, 8)
    assert (g.deg, g.num_intermediates, g.num_constraints) == (6, 2, 12)
    rng = np.random.default_rng(3)
    g.witness(lambda: int(rng.integers(0, 2**63)) % P)
    shift, values, point, got = g._last
    emul = lambda a, b: ((a[0] * b[0] + 7 * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
    xs = [shift * d % P for d in g.domain]
    total = (0, 0)
    for i, xi in enumerate(xs):
        num, den = (1, 0), 1
        for j, xj in enumerate(xs):
            if i != j:
                num, den = emul(num, ((point[0] - xj) % P, point[1])), den * (xi - xj) % P
        t = emul(values[i], emul(num, (pow(den, P - 2, P), 0)))
        total = ((total[0] + t[0]) % P, (total[1] + t[1]) % P)
    assert total == got


@pytest.mark.parametrize("cols,log_n,idx", [(12, 6, 77), (3, 5, 5), (20, 7, 255)])
def test_merkle_proof_circuit_on_a_real_opening(cols, log_n, idx):
    """verify_merkle_proof_to_cap as a circuit (the gadget every FRI query of a recursive proof runs), fed with a real opening
    of a committed batch: constraints hold, the oracle's proof verifies, a tampered path has no witness."""
    import oracle
    import plonk_verifier
    from eth_tx_proof_b200 import circuit as cc, synthetic as syn

    b = oracle.Batch.from_values(syn.random_columns(cols, log_n, seed=cols), 1, 4)
    sib = oracle.merkle_prove(b.digests, 2 << log_n, 4, idx)
    leaf = b.leaves[idx]
    assert oracle.merkle_verify(leaf, idx, sib, b.cap)
    circuit, wires, public_inputs = cc.merkle_proof_circuit(leaf, idx, sib, b.cap)
    assert public_inputs == [int(x) for x in b.cap.reshape(-1)]
    zs_pp = oracle.plonk_partial_products_and_zs(wires[:80], circuit.sigmas, circuit.k_is, 8, [3, 5], [7, 11])
    assert _violations(circuit, wires, zs_pp, cc.hash_no_pad(public_inputs), [3, 5], [7, 11]) == []
    proof = oracle.circuit_prove(circuit, wires, public_inputs, [1, 2, 3, 4])
    plonk_verifier.verify(proof, circuit, proof["constants_sigmas_cap"], [1, 2, 3, 4], max_queries=2)
    bad = sib.copy()
    bad[1, 2] ^= np.uint64(1)
    with pytest.raises(AssertionError, match="copy constraint"):
        cc.merkle_proof_circuit(leaf, idx, bad, b.cap)
    with pytest.raises(AssertionError, match="copy constraint"):
        cc.merkle_proof_circuit(leaf, idx ^ 1, sib, b.cap)  # the sibling's index: the swaps go the wrong way


def test_fri_fold_check_circuit_on_real_fri_data():
    """compute_evaluation of the FRI verifier as a circuit (CosetInterpolationGate), fed with fold instances captured while
    verifying a real circuit proof (16 opened values of a query step, the coset start, beta, the next layer's value)."""
    import oracle
    import plonk_verifier
    from eth_tx_proof_b200 import circuit as cc

    inner, wires, public_inputs = cc.hash_chain_circuit(7, seed=4)  # 2^7 rows: one FRI reduction layer
    proof = oracle.circuit_prove(inner, wires, public_inputs, [9, 9, 9, 9])
    folds = []
    plonk_verifier.verify(proof, inner, proof["constants_sigmas_cap"], [9, 9, 9, 9], max_queries=2, folds=folds)
    assert len(folds) == 2
    for values, coset_start, beta, expected in folds:
        circuit, w, pis = cc.fri_fold_check_circuit(values, coset_start, beta, expected)
        zs_pp = oracle.plonk_partial_products_and_zs(w[:80], circuit.sigmas, circuit.k_is, 8, [3, 5], [7, 11])
        assert _violations(circuit, w, zs_pp, cc.hash_no_pad(pis), [3, 5], [7, 11]) == []
        outer = oracle.circuit_prove(circuit, w, pis, [1, 1, 1, 1])
        plonk_verifier.verify(outer, circuit, outer["constants_sigmas_cap"], [1, 1, 1, 1], max_queries=1)
        with pytest.raises(AssertionError, match="copy constraint"):
            cc.fri_fold_check_circuit(values, coset_start, beta, (expected[0] ^ 1, expected[1]))


def _words_from_oracle_proof(circuit, proof, public_inputs):
    """oracle.circuit_prove's dict -> flat "B200PLK1" words (the layout etp_circuit_prove_* writes)."""
    from eth_tx_proof_b200 import circuit as cc, wire

    op, nc, db = proof["openings"], circuit.num_constants, circuit.degree_bits
    n_layers, bits = 0, db
    while bits > 5:
        n_layers, bits = n_layers + 1, bits - 4
    body = np.concatenate([np.asarray(x, dtype=np.uint64).reshape(-1) for x in (
        proof["wires_cap"], proof["plonk_zs_partial_products_cap"], proof["quotient_polys_cap"], op["constants_sigmas"], op["wires"],
        op["zs_partial_products"][:2], op["plonk_zs_next"], op["zs_partial_products"][2:], op["quotient_polys"], proof["opening_proof"],
        cc.hash_no_pad(public_inputs))])
    hdr = np.zeros(wire.HEADER_WORDS, dtype=np.uint64)
    hdr[:16] = [wire.CIRCUIT_MAGIC, db, nc, 80, 135, 2, 9, 8, 3, 4, n_layers, 4, 1 << bits, 28, 16, wire.HEADER_WORDS + body.size]
    return np.concatenate([hdr, body])


def test_product_transcript_replay_finds_the_openings_the_verifier_checks():
    """circuit.fri_query_openings (product host logic: witness generation for a recursive verifier circuit) on a CPU-made proof:
    the same (leaf, index, path, cap) list, in the same order, as the independent verifier's Merkle checks; and the outer
    circuit built from it is satisfiable and its oracle proof verifies — a recursion step without a GPU."""
    import types

    import oracle
    import plonk_verifier
    from eth_tx_proof_b200 import circuit as cc

    inner, wires, public_inputs = cc.hash_chain_circuit(7, seed=12)
    digest = [4, 3, 2, 1]
    proof = oracle.circuit_prove(inner, wires, public_inputs, digest)
    seen = []
    plonk_verifier.verify(proof, inner, proof["constants_sigmas_cap"], digest, merkle=seen)
    fake_prover = types.SimpleNamespace(c=inner, digest=digest, constants_sigmas_cap=proof["constants_sigmas_cap"])
    words = _words_from_oracle_proof(inner, proof, public_inputs)
    ops = cc.fri_query_openings(fake_prover, words, public_inputs)
    assert len(ops) == len(seen) == 28 * 5
    for (l1, i1, s1, c1), (l2, i2, s2, c2) in zip(ops, seen):
        assert [int(x) for x in l1] == [int(x) for x in l2] and i1 == i2 and s1 == s2 and c1 == c2
    outer, w, pis = cc.merkle_openings_circuit(ops[:10])  # two queries' worth: keeps the pure-Python checks short
    zs_pp = oracle.plonk_partial_products_and_zs(w[:80], outer.sigmas, outer.k_is, 8, [3, 5], [7, 11])
    assert _violations(outer, w, zs_pp, cc.hash_no_pad(pis), [3, 5], [7, 11]) == []
    proof2 = oracle.circuit_prove(outer, w, pis, digest)
    plonk_verifier.verify(proof2, outer, proof2["constants_sigmas_cap"], digest, max_queries=1)


def test_fri_verifier_circuit_on_a_real_proof():
    """verify_fri_proof as a circuit (fri_circuit.py: Merkle openings, subgroup_x, fri_combine_initial, per-layer consistency +
    compute_evaluation, final polynomial) over a CPU-made inner proof: it builds only because every in-circuit value meets the
    proof's (the copy constraints assert equality while building), its constraints hold, the oracle proves it and the verifier
    accepts; a tampered inner proof has no witness."""
    import types

    import oracle
    import plonk_verifier
    from eth_tx_proof_b200 import circuit as cc, fri_circuit as fc

    inner, wires, public_inputs = cc.hash_chain_circuit(7, seed=12)
    digest = [4, 3, 2, 1]
    proof = oracle.circuit_prove(inner, wires, public_inputs, digest)
    fake_prover = types.SimpleNamespace(c=inner, digest=digest, constants_sigmas_cap=proof["constants_sigmas_cap"])
    words = _words_from_oracle_proof(inner, proof, public_inputs)
    outer, w, pis = fc.fri_verifier_circuit([(fake_prover, words, public_inputs)], max_queries=2)
    names = {g.name.split(" ")[0] for g in outer.gates}
    assert {"PoseidonGate", "CosetInterpolationGate", "RandomAccessGate", "ExponentiationGate", "ReducingGate", "ReducingExtensionGate",
            "ArithmeticExtensionGate", "ArithmeticGate", "BaseSumGate", "PublicInputGate"} <= names
    zs_pp = oracle.plonk_partial_products_and_zs(w[:80], outer.sigmas, outer.k_is, 8, [3, 5], [7, 11])
    assert _violations(outer, w, zs_pp, cc.hash_no_pad(pis), [3, 5], [7, 11]) == []
    proof2 = oracle.circuit_prove(outer, w, pis, digest)
    plonk_verifier.verify(proof2, outer, proof2["constants_sigmas_cap"], digest, max_queries=1)
    # the structure does not depend on the data: another inner proof of the same circuit gives the same constants and sigmas
    wires_b, pis_b = cc.hash_chain_circuit(7, seed=13)[1:]
    proof_b = oracle.circuit_prove(inner, wires_b, pis_b, digest)
    outer_b, w_b, _ = fc.fri_verifier_circuit([(fake_prover, _words_from_oracle_proof(inner, proof_b, pis_b), pis_b)], max_queries=2)
    assert (outer_b.constants == outer.constants).all() and (outer_b.sigmas == outer.sigmas).all() and not (w_b == w).all()
    # tampering: a final-polynomial coefficient, an opened row value, a layer value -> no witness
    from eth_tx_proof_b200 import wire

    fri_start = words.size - 4 - proof["opening_proof"].size
    for offset in (words.size - 4 - 3,            # final polynomial
                   fri_start + 64 + 5,            # a value of the first opened row (after the one layer cap)
                   fri_start + 64 + 84 + 4 * 6 + 135 + 4 * 6 + 20 + 4 * 6 + 16 + 4 * 6 + 3):  # a value of the layer row
        bad = words.copy()
        bad[offset] ^= np.uint64(1)
        with pytest.raises(AssertionError):
            fc.fri_verifier_circuit([(fake_prover, bad, public_inputs)], max_queries=1)


def test_verifier_circuit_with_the_vanishing_polynomial_check():
    """fri_verifier_circuit(vanishing=True): besides the FRI verification, the inner circuit's vanishing polynomial is evaluated
    at zeta IN-CIRCUIT (its recorded program over extension targets), combined with the alphas and checked against
    Z_H(zeta) * the reduced quotient chunks; the reduced openings the FRI part uses are computed from the openings in-circuit.
    A changed opening or a changed quotient chunk has no witness."""
    import types

    import oracle
    import plonk_verifier
    from eth_tx_proof_b200 import circuit as cc, fri_circuit as fc

    inner, wires, public_inputs = cc.hash_chain_circuit(6, seed=21)
    digest = [7, 7, 7, 7]
    proof = oracle.circuit_prove(inner, wires, public_inputs, digest)
    fake_prover = types.SimpleNamespace(c=inner, digest=digest, constants_sigmas_cap=proof["constants_sigmas_cap"])
    words = _words_from_oracle_proof(inner, proof, public_inputs)
    outer, w, pis = fc.fri_verifier_circuit([(fake_prover, words, public_inputs)], max_queries=1, vanishing=True)
    zs_pp = oracle.plonk_partial_products_and_zs(w[:80], outer.sigmas, outer.k_is, 8, [3, 5], [7, 11])
    assert _violations(outer, w, zs_pp, cc.hash_no_pad(pis), [3, 5], [7, 11]) == []
    proof2 = oracle.circuit_prove(outer, w, pis, digest)
    plonk_verifier.verify(proof2, outer, proof2["constants_sigmas_cap"], digest, max_queries=1)
    # openings start after the header and the three caps: [constants 4 | sigmas 80 | wires 135 | zs 2 | zs_next 2 | pps 18 | quotient 16] ext
    first_opening = 24 + 3 * 64
    for ext_index in (4 + 80 + 17, 4 + 80 + 135 + 2 + 2 + 18 + 3):  # a wire opening, a quotient chunk opening
        bad = words.copy()
        bad[first_opening + 2 * ext_index] ^= np.uint64(1)
        with pytest.raises(AssertionError):
            fc.fri_verifier_circuit([(fake_prover, bad, public_inputs)], max_queries=1, vanishing=True)


def test_recursive_verifier_circuit_with_the_in_circuit_challenger():
    """recursive_verifier_circuit: the whole verifier — transcript by an in-circuit challenger (its alpha and zeta equal the
    host's, its query-index bits open the right leaves), proof-of-work check, vanishing-polynomial check, FRI — over a CPU-made
    inner proof; constraints hold, the oracle proves the outer circuit and the verifier accepts it; a different proof-of-work
    witness or inner public-input hash has no witness."""
    import types

    import oracle
    import plonk_verifier
    from eth_tx_proof_b200 import circuit as cc, fri_circuit as fc

    inner, wires, public_inputs = cc.hash_chain_circuit(6, seed=31)
    digest = [2, 4, 6, 8]
    proof = oracle.circuit_prove(inner, wires, public_inputs, digest)
    fake_prover = types.SimpleNamespace(c=inner, digest=digest, constants_sigmas_cap=proof["constants_sigmas_cap"])
    words = _words_from_oracle_proof(inner, proof, public_inputs)
    outer, w, pis = fc.recursive_verifier_circuit([(fake_prover, words, public_inputs)], max_queries=2)
    # public inputs: caps | final polynomial | pi_hash | pow witness | openings — no challenge among them
    assert len(pis) == 64 * 5 + 2 * 4 + 4 + 1 + 2 * (4 + 80 + 135 + 2 + 18 + 16 + 2)
    zs_pp = oracle.plonk_partial_products_and_zs(w[:80], outer.sigmas, outer.k_is, 8, [3, 5], [7, 11])
    assert _violations(outer, w, zs_pp, cc.hash_no_pad(pis), [3, 5], [7, 11]) == []
    proof2 = oracle.circuit_prove(outer, w, pis, digest)
    plonk_verifier.verify(proof2, outer, proof2["constants_sigmas_cap"], digest, max_queries=1)
    bad = words.copy()
    bad[words.size - 5] ^= np.uint64(1)  # the proof-of-work witness: the response loses its leading zeros (and the indices move)
    with pytest.raises(AssertionError):
        fc.recursive_verifier_circuit([(fake_prover, bad, public_inputs)], max_queries=1)
    with pytest.raises(AssertionError):  # another circuit digest: the in-circuit transcript diverges from the proof's
        other = types.SimpleNamespace(c=inner, digest=[1, 1, 1, 1], constants_sigmas_cap=proof["constants_sigmas_cap"])
        fc.recursive_verifier_circuit([(other, words, public_inputs)], max_queries=1)


def test_vectorised_goldilocks_product_and_sigma_construction():
    """circuit.gl_mul_vec (numpy, 32-bit limbs) against Python integers on random and edge operands, and the sigma polynomials of a
    built circuit against the definition: every position of a copy set maps to the next one (ascending, cyclic), every other
    position to itself."""
    from eth_tx_proof_b200 import circuit as cc

    rng = np.random.default_rng(9)
    edge = [0, 1, 2, P - 1, P - 2, 2**32, 2**32 - 1, 2**32 + 1, 2**63, P // 2, 0xFFFFFFFF00000000]
    a = np.array([int(x) % P for x in rng.integers(0, 2**64, 4000, dtype=np.uint64)] + edge * len(edge), dtype=np.uint64)
    b = np.array([int(x) % P for x in rng.integers(0, 2**64, 4000, dtype=np.uint64)] + [y for y in edge for _ in edge], dtype=np.uint64)
    assert (cc.gl_mul_vec(a, b) == np.array([int(x) * int(y) % P for x, y in zip(a, b)], dtype=np.uint64)).all()
    for k in (0, 1, 7, P - 1, 2**32):
        assert (cc.gl_mul_vec(a, k) == np.array([int(x) * k % P for x in a], dtype=np.uint64)).all()
    b_ = cc.CircuitBuilder()
    rows = [b_.add_gate(cc.NoopGate(), wires=[5, 5, 5, 9, 9, 1]) for _ in range(3)]
    b_.connect((rows[2], 1), (rows[0], 0))
    b_.connect((rows[0], 0), (rows[1], 2))
    b_.connect((rows[1], 3), (rows[1], 4))
    circuit, _ = b_.build(2)
    n, g, k_is = circuit.n, cc.root_of_unity(circuit.degree_bits), cc.coset_shifts(cc.NUM_ROUTED)
    ident = lambda row, col: k_is[col] * pow(g, row, P) % P
    want = {(r, c): ident(r, c) for r in range(n) for c in range(cc.NUM_ROUTED)}
    # position = column * n + row: (0,0)=0*n+0 < (2,1)=1*n+2 < (1,2)=2*n+1
    want[(rows[0], 0)], want[(rows[2], 1)], want[(rows[1], 2)] = ident(rows[2], 1), ident(rows[1], 2), ident(rows[0], 0)
    want[(rows[1], 3)], want[(rows[1], 4)] = ident(rows[1], 4), ident(rows[1], 3)
    for (r, c), v in want.items():
        assert int(circuit.sigmas[c, r]) == v, (r, c)
