"""GPU tests of the circuit prover (eth_tx_proof_b200/circuit.py — plonky2 0.2.2 plonk/prover.rs prove: wires commitment,
permutation Z / partial products, QUOTIENT of the vanishing polynomial on the device, openings, four-oracle FRI;
/root/reference/Cargo.lock:3441, reached from /root/reference/ops/src/lib.rs:52,72,95).  The quotient is compared with a
pure-Python evaluation of the same vanishing polynomial (coset evaluation, division by Z_H, coset iNTT) and the proofs are
checked by the Python verifier (tests/plonk_verifier.py: transcript, vanishing identity at zeta, FRI)."""
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


def _ntt(a, w):
    """Recursive radix-2 DFT over Python ints: out[k] = sum_j a[j] w^(jk)."""
    n = len(a)
    if n == 1:
        return list(a)
    even, odd = _ntt(a[0::2], w * w % P), _ntt(a[1::2], w * w % P)
    out, t = [0] * n, 1
    for k in range(n // 2):
        u = odd[k] * t % P
        out[k], out[k + n // 2] = (even[k] + u) % P, (even[k] - u) % P
        t = t * w % P
    return out


def _intt(v, w):
    n_inv = pow(len(v), P - 2, P)
    return [x * n_inv % P for x in _ntt(v, pow(w, P - 2, P))]


def _python_quotient(circuit, wires, zs_pp, pi_hash, betas, gammas, alphas):
    """compute_quotient_polys in Python ints: every virtual column interpolated and evaluated on the coset 7 * H_{8n}, the
    program evaluated point-wise, reduced with the alphas (term i * alpha^i), divided by Z_H, coset-iNTT'd: (2, 8n) coeffs."""
    from oracle import pyref as R
    from eth_tx_proof_b200 import cprog

    n, db = circuit.n, circuit.degree_bits
    big = 8 * n
    t = circuit.virtual_trace(wires, zs_pp)
    g_big = R.root_of_unity(db + 3)
    xs = [7 * pow(g_big, k, P) % P for k in range(big)]
    lde = []
    for col in t:
        co = _intt([int(v) for v in col], R.root_of_unity(db))
        lde.append(np.array(_ntt([c * pow(7, i, P) % P for i, c in enumerate(co)] + [0] * (big - n), g_big), dtype=object))
    lv = lde
    nv = [np.roll(c, -8) for c in lde]
    out = circuit.program.evaluate(lv, nv, pi=pi_hash, ch=[int(x) for x in betas] + [int(x) for x in gammas])
    xs_o = np.array(xs, dtype=object)
    zh = np.array([(pow(x, n, P) - 1) % P for x in xs], dtype=object)
    zh_inv = np.array([pow(int(z), P - 2, P) for z in zh], dtype=object)
    l0 = zh * np.array([pow(n * (x - 1) % P, P - 2, P) for x in xs], dtype=object) % P
    res = []
    N = len(out)
    for a in alphas:
        acc = np.zeros(big, dtype=object)
        for idx, (kind, vals) in enumerate(out):
            vals = np.asarray(vals, dtype=object) if not np.isscalar(vals) else np.array([vals] * big, dtype=object)
            if kind == cprog.EMIT_FIRST_ROW:
                vals = vals * l0 % P
            acc = (acc + vals * pow(int(a), N - 1 - idx, P)) % P
        q_vals = [int(v) for v in acc * zh_inv % P]
        co = _intt(q_vals, g_big)
        inv7 = pow(7, P - 2, P)
        res.append([c * pow(inv7, i, P) % P for i, c in enumerate(co)])
    return res


def test_quotient_matches_python_and_proof_verifies_small(ctx):
    """etp_program_register + etp_compute_quotient_polys_cols_dev on the committed oracles' LDE columns == the pure-Python
    quotient, coefficient for coefficient; then the whole proof (one etp_circuit_prove_host call) is verified."""
    import ctypes as C

    import torch

    import eth_tx_proof_b200 as etp
    import oracle
    import plonk_verifier
    from eth_tx_proof_b200 import circuit as cc

    circuit, wires, public_inputs = cc.hash_chain_circuit(5, seed=11)
    n = circuit.n
    pi_hash = cc.hash_no_pad(public_inputs)
    betas, gammas, alphas = [0x1111, 0x2222], [0x3333, 0x4444], [0x5555, 0x6666]
    zs_pp = oracle.plonk_partial_products_and_zs(wires[:80], circuit.sigmas, circuit.k_is, 8, betas, gammas)
    x_coeffs = np.zeros((1, n), dtype=np.uint64)
    x_coeffs[0, 1] = 1
    batches = [etp.PolynomialBatch.from_values(ctx, np.concatenate([circuit.constants, circuit.sigmas]), 3, False, 4),
               etp.PolynomialBatch.from_values(ctx, wires, 3, False, 4), etp.PolynomialBatch.from_values(ctx, zs_pp, 3, False, 4),
               etp.PolynomialBatch.from_coeffs(ctx, x_coeffs, 3, False, 0)]
    cols = []
    for b in batches:
        stride = C.c_size_t()
        base = ctx.L.etp_batch_lde_dev(b.h, C.byref(stride))
        cols += [int(base) + 8 * k * stride.value for k in range(b.n_cols)]
    table = ctx.register_program(circuit.program)
    d_q = torch.zeros((16, n), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ctx.compute_quotient_polys_cols_dev(table, cols, circuit.degree_bits, 3, betas + gammas, pi_hash, alphas, d_q.data_ptr())
    got = d_q.cpu().numpy().view(np.uint64)  # (16, n): challenge j, chunk k = coefficients [k n, (k+1) n)
    want = _python_quotient(circuit, wires, zs_pp, pi_hash, betas, gammas, alphas)
    for j in range(2):
        assert all(c == 0 for c in want[j][8 * n - 8:]), "degree bound: deg(vanishing) <= 9 (n - 1), minus deg Z_H"
        for k in range(8):
            assert [int(v) for v in got[8 * j + k]] == want[j][k * n:(k + 1) * n], f"quotient chunk {j}/{k}"
    # a standalone program is not a starky table
    with pytest.raises(etp.EtpError):
        ctx.stark_prove(table, np.zeros((circuit.program.n_trace, 32), dtype=np.uint64), [0, 0, 0, 0])
    prover = cc.CircuitProver(ctx, circuit)
    proof = prover.prove(wires, public_inputs)
    assert proof["words"].size == prover.proof_words
    plonk_verifier.verify(proof, circuit, prover.constants_sigmas_cap, prover.digest, max_queries=None)


@pytest.mark.parametrize("degree_bits", [7, 10, 12])
def test_circuit_proof_is_accepted_by_the_verifier(ctx, degree_bits):
    import plonk_verifier
    from eth_tx_proof_b200 import circuit as cc

    circuit, wires, public_inputs = cc.hash_chain_circuit(degree_bits, seed=degree_bits)
    prover = cc.CircuitProver(ctx, circuit)
    proof = prover.prove(wires, public_inputs)
    plonk_verifier.verify(proof, circuit, prover.constants_sigmas_cap, prover.digest, max_queries=4 if degree_bits > 7 else None)
    # the same prover proves again (circuit state is reused): same words; and the serde-JSON form round-trips
    proof2 = prover.prove(wires, public_inputs)
    assert (proof2["words"] == proof["words"]).all()
    from eth_tx_proof_b200 import wire

    text = wire.circuit_to_serde_json(proof["words"], public_inputs)
    assert (wire.circuit_from_serde_json(text, degree_bits, cc.hash_no_pad(public_inputs)) == proof["words"]).all()


@pytest.mark.parametrize("degree_bits", [6, 9, 12])
def test_circuit_proof_equals_the_oracles_bit_for_bit(ctx, degree_bits):
    """Every cap, opening, quotient coefficient and FRI word of the device proof == the CPU restatement's (oracle.circuit_prove)."""
    import oracle
    from eth_tx_proof_b200 import circuit as cc

    circuit, wires, public_inputs = cc.hash_chain_circuit(degree_bits, seed=40 + degree_bits)
    prover = cc.CircuitProver(ctx, circuit)
    got = prover.prove(wires, public_inputs)
    want = oracle.circuit_prove(circuit, wires, public_inputs, prover.digest)
    assert (np.asarray(prover.constants_sigmas_cap) == want["constants_sigmas_cap"]).all()
    for k in ("wires_cap", "plonk_zs_partial_products_cap", "quotient_polys_cap"):
        assert (np.asarray(got[k]) == want[k]).all(), k
    for k, v in want["openings"].items():
        assert (np.asarray(got["openings"][k]).reshape(-1) == np.asarray(v).reshape(-1)).all(), k
    assert got["opening_proof"].shape == want["opening_proof"].shape and (got["opening_proof"] == want["opening_proof"]).all()


@pytest.mark.parametrize("degree_bits", [6, 10])
def test_full_gate_set_proof_equals_oracle_and_verifies(ctx, degree_bits):
    """Fourteen gates in five selector groups (extension arithmetic, base sum, reducing, random access, exponentiation, MDS,
    coset interpolation, ...): same parity and acceptance as the Poseidon / arithmetic circuit."""
    import oracle
    import plonk_verifier
    from eth_tx_proof_b200 import circuit as cc

    circuit, wires, public_inputs = cc.hash_chain_circuit(degree_bits, seed=70 + degree_bits, all_gates=True, extra_rows=1 if degree_bits < 8 else 12)
    prover = cc.CircuitProver(ctx, circuit)
    got = prover.prove(wires, public_inputs)
    want = oracle.circuit_prove(circuit, wires, public_inputs, prover.digest)
    for k in ("wires_cap", "plonk_zs_partial_products_cap", "quotient_polys_cap"):
        assert (np.asarray(got[k]) == want[k]).all(), k
    assert (got["opening_proof"] == want["opening_proof"]).all()
    plonk_verifier.verify(got, circuit, prover.constants_sigmas_cap, prover.digest, max_queries=3)


def test_invalid_witness_is_rejected(ctx):
    """A wrong S-box wire / a broken copy constraint / wrong public inputs: the quotient is no longer a polynomial of the
    right degree — the verifier's identity at zeta fails."""
    import plonk_verifier
    from eth_tx_proof_b200 import circuit as cc

    circuit, wires, public_inputs = cc.hash_chain_circuit(7, seed=3)
    prover = cc.CircuitProver(ctx, circuit)
    row = int(np.nonzero(circuit.gate_of_row == 4)[0][5])
    for mutate in ("sbox", "copy", "pi"):
        w, pi = wires.copy(), list(public_inputs)
        if mutate == "sbox":
            w[cc.PoseidonGate.wire_partial_sbox(3), row] ^= np.uint64(1)
        elif mutate == "copy":
            w[1, row] = (int(w[1, row]) + 5) % P
        else:
            pi[0] = (pi[0] + 1) % P
        proof = prover.prove(w, pi)
        with pytest.raises(plonk_verifier.VerifyError):
            plonk_verifier.verify(proof, circuit, prover.constants_sigmas_cap, prover.digest, max_queries=1)


def test_circuit_prove_dev_equals_prove_host_and_misuse_is_refused(ctx):
    """etp_circuit_prove_dev (witness already on the device, any column stride) == etp_circuit_prove_host; malformed circuits
    (wrong virtual column count, wrong degree, FRI parameters of another size) are refused at etp_circuit_create."""
    import ctypes as C

    import torch

    import eth_tx_proof_b200 as etp
    from eth_tx_proof_b200 import circuit as cc, cprog

    circuit, wires, public_inputs = cc.hash_chain_circuit(6, seed=21)
    prover = cc.CircuitProver(ctx, circuit)
    want = prover.prove_words(wires, public_inputs)
    n, stride = circuit.n, circuit.n + 24
    padded = np.zeros((cc.NUM_WIRES, stride), dtype=np.uint64)
    padded[:, :n] = wires
    d = torch.from_numpy(padded.view(np.int64)).cuda()
    torch.cuda.synchronize()
    out = np.zeros(prover.proof_words, dtype=np.uint64)
    ptr = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint64))
    pi_hash = np.array(cc.hash_no_pad(public_inputs), dtype=np.uint64)
    ctx.check(ctx.L.etp_circuit_prove_dev(prover.h, C.c_void_p(d.data_ptr()), stride, ptr(pi_hash), ptr(out)))
    assert (out == want).all()
    with pytest.raises(etp.EtpError):
        ctx.check(ctx.L.etp_circuit_prove_dev(prover.h, C.c_void_p(d.data_ptr()), n - 1, ptr(pi_hash), ptr(out)))
    # an explicit digest changes the transcript (and only that)
    other = cc.CircuitProver(ctx, circuit, circuit_digest=[1, 2, 3, 4])
    assert other.digest == [1, 2, 3, 4] and (other.constants_sigmas_cap == prover.constants_sigmas_cap).all()
    assert not (other.prove_words(wires, public_inputs) == want).all()

    def create(program, consts=circuit.constants, fri=None, qdf=8):
        u64 = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.uint64))
        prog, cs, sg, ks = u64(program.words), u64(consts), u64(circuit.sigmas), u64(circuit.k_is)
        fp = fri or etp.FriParams.make(circuit.degree_bits, 3, 4, 16, 28)
        h = C.c_void_p()
        return ctx.L.etp_circuit_create(ctx.h, ptr(prog), prog.size, ptr(cs), cs.shape[0], ptr(sg), ptr(ks), 80, 135, circuit.degree_bits, qdf, 2,
                                        C.byref(fp), None, C.byref(h)), h

    b = cprog.ProgramBuilder(circuit.num_virtual_columns - 1, 4, 9)  # one virtual column short
    b.constraint(b.lv(0))
    assert create(b.build())[0] == -1 and "virtual columns" in ctx.L.etp_last_error(ctx.h).decode()
    b = cprog.ProgramBuilder(circuit.num_virtual_columns, 4, 4)      # constraint degree 4 for a quotient degree factor of 8
    b.constraint(b.lv(0))
    assert create(b.build())[0] == -1
    assert create(circuit.program, fri=etp.FriParams.make(circuit.degree_bits + 1, 3, 4, 16, 28))[0] == -1
    assert create(circuit.program, fri=etp.FriParams.make(circuit.degree_bits, 1, 4, 16, 28))[0] == -1  # 2^rate_bits < quotient degree factor
    rc, h = create(circuit.program)
    assert rc == 0
    ctx.L.etp_circuit_free(h)


def test_merkle_proof_circuit_over_a_device_commitment(ctx):
    """The recursive verifier's Merkle gadget as a circuit, fed with an opening of a batch committed ON THE DEVICE (row, path,
    cap), proved on the device and verified; a prover that swaps in a wrong sibling afterwards is rejected."""
    import eth_tx_proof_b200 as etp
    import oracle
    import plonk_verifier
    from eth_tx_proof_b200 import circuit as cc, synthetic as syn

    b = etp.PolynomialBatch.from_values(ctx, syn.random_columns(21, 10, seed=5), 1, False, 4)
    idx = 1234
    leaf, sib, cap = b.leaves_at([idx])[0], b.prove(idx), b.cap
    circuit, wires, public_inputs = cc.merkle_proof_circuit(leaf, idx, sib, cap)
    prover = cc.CircuitProver(ctx, circuit)
    proof = prover.prove(wires, public_inputs)
    plonk_verifier.verify(proof, circuit, prover.constants_sigmas_cap, prover.digest, max_queries=3)
    want = oracle.circuit_prove(circuit, wires, public_inputs, prover.digest)
    assert (proof["opening_proof"] == want["opening_proof"]).all()
    bad = wires.copy()
    row = int(np.nonzero(circuit.gate_of_row == len(circuit.gates) - 1)[0][-1])  # the last PoseidonGate row: a path level
    bad[5, row] = (int(bad[5, row]) + 1) % P  # a sibling element, the other wires of the row left as they were
    with pytest.raises(plonk_verifier.VerifyError):
        plonk_verifier.verify(prover.prove(bad, public_inputs), circuit, prover.constants_sigmas_cap, prover.digest, max_queries=1)


def test_fri_fold_check_circuit_over_a_device_proof(ctx):
    """A proof ABOUT a proof: fold instances of the FRI verification of a device-made circuit proof (compute_evaluation: 16
    opened values, coset start, beta, next value) go through the CosetInterpolationGate circuit, proved on the device."""
    import plonk_verifier
    from eth_tx_proof_b200 import circuit as cc

    inner, wires, public_inputs = cc.hash_chain_circuit(7, seed=6)
    p_in = cc.CircuitProver(ctx, inner)
    proof = p_in.prove(wires, public_inputs)
    folds = []
    plonk_verifier.verify(proof, inner, p_in.constants_sigmas_cap, p_in.digest, max_queries=1, folds=folds)
    values, coset_start, beta, expected = folds[0]
    circuit, w, pis = cc.fri_fold_check_circuit(values, coset_start, beta, expected)
    p_out = cc.CircuitProver(ctx, circuit)
    plonk_verifier.verify(p_out.prove(w, pis), circuit, p_out.constants_sigmas_cap, p_out.digest, max_queries=2)


def test_recursive_merkle_verifier_chain(ctx):
    """A recursion chain on the device: proof 0 of a base circuit; circuit 1 checks EVERY Merkle opening of proof 0's FRI
    verification (transcript replayed by the product to get the query indices: the openings must be the ones the independent
    verifier checks) and is proved; circuit 2 does the same for proof 1.  Each outer proof is verified; the outer circuit's
    structure does not depend on the inner proof's data (same constants and sigmas for another inner proof)."""
    import plonk_verifier
    from eth_tx_proof_b200 import circuit as cc

    base, wires, pis = cc.hash_chain_circuit(7, seed=8)
    p0 = cc.CircuitProver(ctx, base)
    w0 = p0.prove_words(wires, pis)
    seen = []
    plonk_verifier.verify(p0.prove(wires, pis), base, p0.constants_sigmas_cap, p0.digest, merkle=seen)
    ops = cc.fri_query_openings(p0, w0, pis)
    assert len(ops) == len(seen) == 28 * 5
    for (l1, i1, s1, c1), (l2, i2, s2, c2) in zip(ops, seen):
        assert list(l1) == list(l2) and i1 == i2 and s1 == s2 and c1 == c2
    c1_, w1, pi1 = cc.recursive_merkle_verifier_circuit([(p0, w0, pis)])
    p1 = cc.CircuitProver(ctx, c1_)
    proof1 = p1.prove(w1, pi1)
    plonk_verifier.verify(proof1, c1_, p1.constants_sigmas_cap, p1.digest, max_queries=2)
    # the same outer circuit for ANOTHER inner proof of the same shape: only the witness changes
    wires_b, pis_b = cc.hash_chain_circuit(7, seed=9)[1:]
    c1b, w1b, pi1b = cc.recursive_merkle_verifier_circuit([(p0, p0.prove_words(wires_b, pis_b), pis_b)])
    assert (c1b.constants == c1_.constants).all() and (c1b.sigmas == c1_.sigmas).all() and not (w1b == w1).all()
    plonk_verifier.verify(p1.prove(w1b, pi1b), c1_, p1.constants_sigmas_cap, p1.digest, max_queries=1)
    # one more layer: the outer proof is itself the inner proof of the next circuit
    c2_, w2, pi2 = cc.recursive_merkle_verifier_circuit([(p1, proof1["words"], pi1)])
    p2 = cc.CircuitProver(ctx, c2_)
    plonk_verifier.verify(p2.prove(w2, pi2), c2_, p2.constants_sigmas_cap, p2.digest, max_queries=1)


def test_fri_verifier_circuit_over_a_device_proof(ctx):
    """The whole FRI verification (28 queries: Merkle openings, combine, per-layer consistency + interpolation, final
    polynomial) of a device-made inner proof as an outer circuit (~2.7 k rows -> 2^12), proved on the device and verified;
    and one more layer: the FRI verifier circuit of THAT proof."""
    import plonk_verifier
    from eth_tx_proof_b200 import circuit as cc, fri_circuit as fc

    inner, wires, public_inputs = cc.hash_chain_circuit(7, seed=15)
    p0 = cc.CircuitProver(ctx, inner)
    w0 = p0.prove_words(wires, public_inputs)
    c1, w1, pi1 = fc.fri_verifier_circuit([(p0, w0, public_inputs)])
    assert c1.degree_bits == 12
    p1 = cc.CircuitProver(ctx, c1)
    proof1 = p1.prove(w1, pi1)
    plonk_verifier.verify(proof1, c1, p1.constants_sigmas_cap, p1.digest, max_queries=2)
    c2, w2, pi2 = fc.fri_verifier_circuit([(p1, proof1["words"], pi1)], max_queries=4)
    p2 = cc.CircuitProver(ctx, c2)
    plonk_verifier.verify(p2.prove(w2, pi2), c2, p2.constants_sigmas_cap, p2.digest, max_queries=1)


def test_recursive_verifier_chain_reaches_a_fixed_point_on_the_device(ctx):
    """A complete recursion chain on the device: P0 proves a base circuit; C1 = the recursive verifier of P0 (in-circuit
    challenger, proof-of-work, vanishing-polynomial check, all 28 FRI queries) is proved (P1); C2 = the recursive verifier of P1
    is proved (P2) — C1 and C2 are both 2^13-row circuits (the size is a fixed point of the recursion).  Every proof is
    accepted by the independent Python verifier."""
    import plonk_verifier
    from eth_tx_proof_b200 import circuit as cc, fri_circuit as fc

    base, wires, public_inputs = cc.hash_chain_circuit(12, seed=41, witness_seed=1)
    p0 = cc.CircuitProver(ctx, base)
    w0 = p0.prove_words(wires, public_inputs)
    c1, w1, pi1 = fc.recursive_verifier_circuit([(p0, w0, public_inputs)])
    assert c1.degree_bits == 13
    p1 = cc.CircuitProver(ctx, c1)
    proof1 = p1.prove(w1, pi1)
    plonk_verifier.verify(proof1, c1, p1.constants_sigmas_cap, p1.digest, max_queries=1)
    c2, w2, pi2 = fc.recursive_verifier_circuit([(p1, proof1["words"], pi1)])
    assert c2.degree_bits == 13
    p2 = cc.CircuitProver(ctx, c2)
    proof2 = p2.prove(w2, pi2)
    plonk_verifier.verify(proof2, c2, p2.constants_sigmas_cap, p2.digest, max_queries=1)
    # the layer's circuit is data-independent: the verifier of ANOTHER proof of C1 is the same circuit, so p2 proves it too
    base_b, wires_b, pis_b = cc.hash_chain_circuit(12, seed=41, witness_seed=2)  # another witness of the SAME base circuit
    assert (base_b.constants == base.constants).all() and (base_b.sigmas == base.sigmas).all() and not (wires_b == wires).all()
    c1b, w1b, pi1b = fc.recursive_verifier_circuit([(p0, p0.prove_words(wires_b, pis_b), pis_b)])
    assert (c1b.constants == c1.constants).all() and (c1b.sigmas == c1.sigmas).all()
    c2b, w2b, pi2b = fc.recursive_verifier_circuit([(p1, p1.prove_words(w1b, pi1b), pi1b)])
    assert (c2b.constants == c2.constants).all() and (c2b.sigmas == c2.sigmas).all()
    plonk_verifier.verify(p2.prove(w2b, pi2b), c2, p2.constants_sigmas_cap, p2.digest, max_queries=1)
    # and an INVALID inner proof (a witness of a different circuit) is rejected by the recursive verifier: no witness
    wrong = cc.hash_chain_circuit(12, seed=43)
    with pytest.raises(AssertionError, match="copy constraint"):
        fc.recursive_verifier_circuit([(p0, p0.prove_words(wrong[1], wrong[2]), wrong[2])])
