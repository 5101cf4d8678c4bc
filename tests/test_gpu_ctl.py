"""GPU parity of the round-2 boundary: prove_with_commitment with the challenger state in/out and CTL data, the general
auxiliary columns (Lookups with linear-combination Columns / Filters, CTL helper columns and Z running sums), the general
prove_openings / FRI instance (plonky2's four-oracle shape under standard_recursion_config), the stepwise FRI API and
eval_at_ext_point — all through the C ABI, bit for bit against the oracle, and accepted by the Python verifier.
Upstream: starky 0.4.0 src/{prover,cross_table_lookup,lookup}.rs, plonky2 0.2.2 src/fri/{oracle,prover}.rs
(/root/reference/Cargo.lock:4529,3441; reached from /root/reference/ops/src/lib.rs:52,72,95)."""
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


def dev(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


@pytest.mark.parametrize("sizes", [(5, 4, 4), (7, 6, 5), (10, 8, 9)])
def test_general_aux_columns_match_oracle(ctx, sizes):
    import torch

    import oracle
    from eth_tx_proof_b200 import cprog

    tables, _ = cprog.ctl_demo_tables(*sizes)
    lookup_ch = [0x1234567890ABCDEF % P, 0x0FEDCBA987654321]
    ctl_ch = [lookup_ch[0], 777777777777, lookup_ch[1], 31337]
    for name, prog, trace in tables:
        tid = ctx.register_table(prog)
        oid = oracle.register_table_ex(prog, prog.aux_spec)
        L = ctx.L
        assert L.etp_table_num_aux_columns(ctx.h, tid, 2) == prog.n_aux
        assert L.etp_table_num_lookup_columns(ctx.h, tid, 2) == prog.n_lookup_cols
        assert L.etp_table_num_ctl_helper_columns(ctx.h, tid) == prog.n_ctl_helper_cols
        assert L.etp_table_num_ctl_zs(ctx.h, tid) == len(prog.ctl_zs)
        want = oracle.aux_columns(oid, trace, lookup_ch, ctl_ch)
        n = trace.shape[1]
        d = dev(trace)
        out = torch.zeros((prog.n_aux, n), dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        ctx.aux_columns_dev(tid, int(n).bit_length() - 1, d.data_ptr(), n, lookup_ch, ctl_ch, out.data_ptr())
        got = out.cpu().numpy().view(np.uint64)
        assert (got == want).all(), name


def test_zero_denominator_is_reported_like_upstreams_panic(ctx):
    """batch_multiplicative_inverse panics on a zero upstream ("Tried to invert zero"); here: ETP_ERR_PROOF."""
    import torch

    import eth_tx_proof_b200 as etp
    from eth_tx_proof_b200 import cprog

    tables, _ = cprog.ctl_demo_tables(5, 4, 4)
    name, prog, trace = tables[1]  # rom: combine = KEY + beta * VAL + gamma; choose gamma = -(KEY + beta * VAL) at row 3
    tid = ctx.register_table(prog)
    beta = 5
    gamma = (-(int(trace[0, 3]) + beta * int(trace[1, 3]))) % P
    n = trace.shape[1]
    d = dev(trace)
    out = torch.zeros((prog.n_aux, n), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    with pytest.raises(etp.EtpError) as e:
        ctx.aux_columns_dev(tid, 4, d.data_ptr(), n, [beta, 9], [beta, gamma, 9, 10], out.data_ptr())
    assert e.value.code == -4
    ctx.aux_columns_dev(tid, 4, d.data_ptr(), n, [beta, 9], [beta, gamma + 1, 9, 10], out.data_ptr())  # and the context still works


@pytest.mark.parametrize("sizes", [(6, 5, 5), (9, 7, 8), (12, 10, 11)])
def test_multi_table_ctl_proofs_match_oracle_and_verify(ctx, sizes):
    """evm_arithmetization's prove_with_traces shape on the synthetic three-table CTL system: every table's proof, the CTL
    challenges and the challenger state after every table equal the oracle's; the verifier accepts incl. the CTL sums."""
    import oracle
    import stark_verifier as V
    from eth_tx_proof_b200 import cprog, prover
    from test_ctl_oracle import verify_all

    tables, ctls = cprog.ctl_demo_tables(*sizes)
    tids = [ctx.register_table(p) for _, p, _ in tables]
    devs = [dev(t) for _, _, t in tables]
    import torch

    torch.cuda.synchronize()
    traces_dev = [(d.data_ptr(), t.shape[1], t.shape[0], int(t.shape[1]).bit_length() - 1) for d, (_, _, t) in zip(devs, tables)]
    got = prover.prove_with_traces(ctx, tids, traces_dev)
    # the oracle, step by step, with the challenger state compared after every table
    oids = [oracle.register_table_ex(p, p.aux_spec) for _, p, _ in tables]
    batches = [oracle.Batch.from_values(t, 1, 4) for _, _, t in tables]
    och = oracle.HostChallenger()
    for b, cap in zip(batches, got.trace_caps):
        assert (b.cap == cap).all()
        och.observe(b.cap)
    octl = och.get_n(4)
    assert (octl == got.ctl_challenges).all()
    for k, (oid, (_, prog, t), b) in enumerate(zip(oids, tables, batches)):
        assert (och.compact() == got.init_challenger_states[k]).all()
        want = oracle.prove_with_commitment(oid, t, b, och, octl)
        assert got.stark_proofs[k].size == want.size
        assert (np.delete(got.stark_proofs[k], 1) == np.delete(want, 1)).all(), f"table {k} proof differs from the oracle"
    zs = verify_all(tables, ctls, got.stark_proofs, got.trace_caps, max_queries=2)
    assert [len(z) for z in zs] == [2, 2, 2]


def test_prove_with_commitment_challenger_state_in_out(ctx):
    """Stand-alone table through prove_with_commitment: equals etp_stark_prove when the caller replays prove()'s
    prologue, and the challenger handed back equals the oracle's after its proof."""
    import torch

    import eth_tx_proof_b200 as etp
    import oracle
    from eth_tx_proof_b200 import synthetic as syn

    t = syn.memory_trace(9, seed=21)
    d = dev(t)
    torch.cuda.synchronize()
    com = etp.PolynomialBatch.from_values_dev(ctx, d.data_ptr(), t.shape[1], 21, 9, 1, False, 4)
    ch = etp.Challenger()
    ch.observe_cap(com.cap)
    proof = ctx.prove_with_commitment(etp.TABLE_MEMORY, com, d.data_ptr(), t.shape[1], ch)
    assert (proof == ctx.stark_prove(etp.TABLE_MEMORY, t)).all()
    ob = oracle.Batch.from_values(t, 1, 4)
    och = oracle.HostChallenger()
    och.observe(ob.cap)
    want = oracle.prove_with_commitment(oracle.TABLE_MEMORY, t, ob, och)
    assert (proof == want).all()
    assert (ch.words() == och.words()).all()
    # a different incoming transcript gives a different proof
    ch2 = etp.Challenger()
    ch2.observe([1])
    ch2.observe_cap(com.cap)
    assert not (ctx.prove_with_commitment(etp.TABLE_MEMORY, com, d.data_ptr(), t.shape[1], ch2) == proof).all()


def test_ctl_table_needs_ctl_challenges(ctx):
    import eth_tx_proof_b200 as etp
    from eth_tx_proof_b200 import cprog

    tables, _ = cprog.ctl_demo_tables(5, 4, 4)
    _, prog, trace = tables[1]
    tid = ctx.register_table(prog)
    with pytest.raises(etp.EtpError):
        ctx.stark_prove(tid, trace)  # starky::prove has no ctl_data: a CTL table must go through prove_with_commitment


@pytest.mark.parametrize("degree_bits,rate_bits,queries", [(8, 3, 28), (12, 3, 28), (10, 1, 84), (5, 3, 28)])
def test_prove_openings_general_instance_matches_oracle(ctx, degree_bits, rate_bits, queries):
    """plonky2's circuit-prover shape: four oracles (constants+sigmas, wires, Z+partial products, quotient), the zeta batch
    over all of them and the g*zeta batch over a NON-prefix subset (the Zs), under standard_recursion_config
    (rate_bits 3, cap_height 4, 28 queries) — plonky2/src/plonk/prover.rs, /root/reference/ops/src/lib.rs:72,95."""
    import eth_tx_proof_b200 as etp
    import oracle
    from eth_tx_proof_b200 import synthetic as syn

    shapes = [9, 20, 6, 4]
    cols = [syn.random_columns(c, degree_bits, seed=900 + i) for i, c in enumerate(shapes)]
    gb = [etp.PolynomialBatch.from_values(ctx, v, rate_bits, False, 4) for v in cols]
    ob = [oracle.Batch.from_values(v, rate_bits, 4) for v in cols]
    zeta = [123456789123456789 % P, 987654321987654321 % P]
    from oracle import pyref

    g = pyref.root_of_unity(degree_bits)
    zeta_next = [zeta[0] * g % P, zeta[1] * g % P]
    all_polys = [(o, c) for o, n in enumerate(shapes) for c in range(n)]
    zs = [(2, 0), (2, 1)]
    batches = [(zeta, all_polys), (zeta_next, zs)]
    fp = etp.FriParams.make(degree_bits, rate_bits, 4, 16, queries)
    ofp = oracle.fri_params(degree_bits, rate_bits, 4, 16, queries)
    gch, och = etp.Challenger(), oracle.HostChallenger()
    for b, o in zip(gb, ob):
        assert (b.cap == o.cap).all()
        gch.observe_cap(b.cap)
        och.observe(o.cap)
    got = ctx.prove_openings(batches, gb, gch, fp)
    want = oracle.prove_openings(batches, ob, och, ofp)
    assert got.size == want.size
    assert (got == want).all()
    assert (gch.words() == och.words()).all()


def test_eval_at_ext_point_matches_oracle(ctx):
    import eth_tx_proof_b200 as etp
    import oracle
    from eth_tx_proof_b200 import synthetic as syn

    for log_n, cols in ((4, 3), (11, 21), (13, 5)):
        v = syn.random_columns(cols, log_n, seed=5)
        b = etp.PolynomialBatch.from_values(ctx, v, 1, False, 4)
        o = oracle.Batch.from_values(v, 1, 4)
        for z in ([3, 0], [P - 1, P - 2], [0x1111111111111111, 0x2222222222222222]):
            assert (b.eval_at_ext_point(z) == oracle.batch_eval_at_ext_point(o, z)).all()


def test_stepwise_fri_equals_fused_and_oracle(ctx):
    """etp_fri_begin / commit_layer / fold / final_poly / query_rounds driven by a caller-side challenger == the fused
    commit phase == the FriProof inside etp_prove_openings."""
    import torch

    import eth_tx_proof_b200 as etp
    import oracle
    from eth_tx_proof_b200 import synthetic as syn

    degree_bits, rate_bits = 10, 1
    # a low-degree ext polynomial on the coset: c0/c1 parts from two LDE columns (bit-reversed order, as committed)
    v = syn.random_columns(2, degree_bits, seed=77)
    b = etp.PolynomialBatch.from_values(ctx, v, rate_bits, False, 4)
    lde_n = 1 << (degree_bits + rate_bits)
    leaves = b.leaves  # (lde_n, 2): row p = point bitrev(p) -> interleaved ext values in bit-reversed order
    vals = dev(leaves.reshape(-1))
    torch.cuda.synchronize()
    fp = etp.FriParams.make(degree_bits, rate_bits, 4, 16, 84)
    ch1, ch2 = etp.Challenger(), etp.Challenger()
    for c in (ch1, ch2):
        c.observe([42])
    s1 = etp.FriState(ctx, vals.data_ptr(), fp)
    caps = []
    for _ in range(fp.n_reductions):
        cap = s1.commit_layer()
        caps.append(cap)
        ch1.observe_cap(cap)
        s1.fold(ch1.get_extension_challenge())
    fin = s1.final_poly()
    ch1.observe(fin)
    s2 = etp.FriState(ctx, vals.data_ptr(), fp)
    caps2, fin2 = s2.commit_phase(ch2)
    assert (np.array(caps) == caps2).all() and (fin == fin2).all() and (ch1.words() == ch2.words()).all()
    idx = [0, 1, lde_n - 1, 1234 % lde_n]
    q1, q2 = s1.query_rounds([b], idx), s2.query_rounds([b], idx)
    assert (q1 == q2).all()
    # the same layers through the oracle's coefficient-form fold: final polynomial equal
    co = oracle.Batch.from_values(v, rate_bits, 4).coeffs  # (2, n)
    coeffs = np.zeros((lde_n, 2), dtype=np.uint64)
    coeffs[: 1 << degree_bits, 0], coeffs[: 1 << degree_bits, 1] = co[0], co[1]
    och = oracle.HostChallenger()
    och.observe([42])
    for cap in caps:
        och.observe(cap)
        beta = och.get_n(2)
        coeffs = oracle.fri_fold_coeffs(coeffs, 4, beta)
    assert (coeffs[: fin.shape[0]] == fin).all() and not coeffs[fin.shape[0]:].any()
    with pytest.raises(etp.EtpError):
        s1.fold([1, 2])  # nothing left to fold


@pytest.mark.parametrize("degree_bits", [6, 9, 12])
def test_recursion_proof_skeleton_matches_oracle(ctx, degree_bits):
    """plonky2's circuit-prover skeleton under standard_recursion_config (eth_tx_proof_b200/recursion.py): the four caps, the
    openings, the FriProof and the final transcript state equal the oracle's on the same stand-in polynomials."""
    import oracle
    from eth_tx_proof_b200 import recursion as rec

    polys = rec.stand_in_polys(degree_bits, seed=degree_bits)
    got = rec.prove_skeleton(ctx, degree_bits, polys)
    ob = [oracle.Batch.from_values(polys["constants_sigmas"], 3, 4), oracle.Batch.from_values(polys["wires"], 3, 4), None,
          oracle.Batch.from_coeffs(polys["quotient"], 3, 4)]
    och = oracle.HostChallenger()
    och.observe(ob[0].cap)
    och.observe(ob[1].cap)
    betas, gammas = och.get_n(2), och.get_n(2)
    assert (betas == got["betas"]).all() and (gammas == got["gammas"]).all()
    # Z and partial products from the device == plonky2's all_wires_permutation_partial_products as restated by the oracle,
    # and the permutation argument closes: Z(x_{n-1}) times the last row's quotient product is 1
    want_z = oracle.plonk_partial_products_and_zs(polys["wires"][:rec.NUM_ROUTED], polys["constants_sigmas"][rec.NUM_CONSTANTS:], polys["k_is"],
                                                  rec.QUOTIENT_DEGREE_FACTOR, betas, gammas)
    assert (got["zs_partial_products"] == want_z).all()
    assert (want_z[:2, 0] == 1).all()
    ob[2] = oracle.Batch.from_values(want_z, 3, 4)
    och.observe(ob[2].cap)
    assert (och.get_n(2) == got["alphas"]).all()
    och.observe(ob[3].cap)
    zeta = och.get_n(2)
    assert (zeta == got["zeta"]).all()
    for o, cap, op in zip(ob, got["caps"], got["openings"]):
        assert (o.cap == cap).all()
        want = oracle.batch_eval_at_ext_point(o, zeta)
        assert (want == op).all()
        och.observe(want)
    g = rec.root_of_unity(degree_bits)
    zn = [int(zeta[0]) * g % P, int(zeta[1]) * g % P]
    och.observe(oracle.batch_eval_at_ext_point(ob[2], zn)[:2])
    want = oracle.prove_openings(rec.fri_instance(zeta, degree_bits), ob, och, oracle.fri_params(degree_bits, 3, 4, 16, 28))
    assert (got["fri_proof"] == want).all()
    assert (got["challenger"].words() == och.words()).all()


def test_evm_shaped_transaction_matches_oracle(ctx):
    """The seven-table transaction with upstream's CTL topology (2400-column keccak shape, bit-decomposed logic, memory with its
    range-check lookup): every proof equals the oracle's and the CTL sums close."""
    import torch

    import oracle
    from eth_tx_proof_b200 import cprog, prover
    from test_ctl_oracle import prove_all, verify_all

    bits = {"arithmetic": 7, "byte_packing": 5, "cpu": 6, "keccak": 6, "keccak_sponge": 5, "logic": 6, "memory": 8}
    tables, ctls = cprog.evm_shaped_system(degree_bits=bits, seed=5)
    tids = [ctx.register_table(p) for _, p, _ in tables]
    devs = [dev(t) for _, _, t in tables]
    torch.cuda.synchronize()
    traces_dev = [(d.data_ptr(), t.shape[1], t.shape[0], int(t.shape[1]).bit_length() - 1) for d, (_, _, t) in zip(devs, tables)]
    got = prover.prove_with_traces(ctx, tids, traces_dev)
    want, caps = prove_all(tables)
    for k in range(7):
        assert (np.delete(got.stark_proofs[k], 1) == np.delete(want[k], 1)).all(), tables[k][0]
    verify_all(tables, ctls, got.stark_proofs, got.trace_caps, max_queries=1)


def test_prove_openings_edge_instances(ctx):
    """One batch, four batches, a batch over a single polynomial, degree too small for any FRI reduction — against the oracle;
    and the misuse the C ABI must reject (upstream: panics)."""
    import eth_tx_proof_b200 as etp
    import oracle
    from eth_tx_proof_b200 import synthetic as syn

    for degree_bits, rate_bits in ((4, 1), (7, 2), (9, 1)):
        cols = [syn.random_columns(c, degree_bits, seed=40 + i) for i, c in enumerate((3, 5))]
        gb = [etp.PolynomialBatch.from_values(ctx, v, rate_bits, False, 2) for v in cols]
        ob = [oracle.Batch.from_values(v, rate_bits, 2) for v in cols]
        pts = [[5, 6], [8, 0], [P - 3, 11], [1, 0]]  # ([7, 0] would be a point OF the LDE coset: see the next test)
        instances = [
            [(pts[0], [(0, 0), (0, 1), (0, 2), (1, 0), (1, 1), (1, 2), (1, 3), (1, 4)])],
            [(pts[0], [(1, 4)]), (pts[1], [(0, 2)])],
            [(pts[0], [(0, 0), (1, 0)]), (pts[1], [(0, 0)]), (pts[2], [(1, 0), (0, 0)]), (pts[3], [(1, 3), (1, 2), (0, 1)])],
        ]
        fp = etp.FriParams.make(degree_bits, rate_bits, 2, 3, 5)
        ofp = oracle.fri_params(degree_bits, rate_bits, 2, 3, 5)
        for inst in instances:
            gch, och = etp.Challenger(), oracle.HostChallenger()
            gch.observe([degree_bits])
            och.observe([degree_bits])
            got = ctx.prove_openings(inst, gb, gch, fp)
            want = oracle.prove_openings(inst, ob, och, ofp)
            assert (got == want).all() and (gch.words() == och.words()).all(), (degree_bits, len(inst))
    b = etp.PolynomialBatch.from_values(ctx, syn.random_columns(2, 6, seed=1), 1, False, 4)
    fp = etp.FriParams.make(6, 1, 4, 16, 84)
    with pytest.raises(etp.EtpError):  # polynomial index out of range
        ctx.prove_openings([([1, 2], [(0, 2)])], [b], etp.Challenger(), fp)
    with pytest.raises(etp.EtpError):  # oracle committed with another rate
        ctx.prove_openings([([1, 2], [(0, 0)])], [b], etp.Challenger(), etp.FriParams.make(6, 2, 4, 16, 84))
    with pytest.raises(etp.EtpError):  # five batches
        ctx.prove_openings([([k, 2], [(0, 0)]) for k in range(5)], [b], etp.Challenger(), fp)
    with pytest.raises(etp.EtpError):  # the same polynomial twice in one batch
        ctx.prove_openings([([1, 2], [(0, 0), (0, 0)])], [b], etp.Challenger(), fp)
    bad = etp.Challenger()
    bad.input_len = 9
    with pytest.raises(etp.EtpError):  # corrupt transcript state
        ctx.prove_openings([([1, 2], [(0, 0)])], [b], bad, fp)


def test_opening_point_on_the_lde_coset_is_reported(ctx):
    """(x - z) vanishes at an LDE point: upstream's divide_by_linear is fine in coefficient form, the evaluation-form path
    must say so instead of returning garbage."""
    import eth_tx_proof_b200 as etp
    from eth_tx_proof_b200 import synthetic as syn

    b = etp.PolynomialBatch.from_values(ctx, syn.random_columns(2, 5, seed=2), 1, False, 4)
    with pytest.raises(etp.EtpError) as e:
        ctx.prove_openings([([7, 0], [(0, 0), (0, 1)])], [b], etp.Challenger(), etp.FriParams.make(5, 1, 4, 4, 3))  # 7 = the coset shift
    assert e.value.code == -4


def test_register_ex_rejects_malformed_specs(ctx):
    import eth_tx_proof_b200 as etp
    from eth_tx_proof_b200 import cprog

    tables, _ = cprog.ctl_demo_tables(5, 4, 4)
    _, prog, _ = tables[0]
    spec = prog.aux_spec.copy()
    ok = ctx.register_table_ex(prog, spec)
    assert ok >= 16
    for mutate in (lambda s: s[:-1], lambda s: np.concatenate([s, [0]]).astype(np.uint64),
                   lambda s: np.concatenate([[1], s[1:]]).astype(np.uint64),                    # bad magic
                   lambda s: np.concatenate([s[:1], [s[1] + 1], s[2:]]).astype(np.uint64)):     # one more lookup than described
        with pytest.raises(etp.EtpError):
            ctx.register_table_ex(prog, mutate(spec))
    col = spec.copy()
    col[5] = 1000  # first looking column reads trace column 1000
    with pytest.raises(etp.EtpError):
        ctx.register_table_ex(prog, col)
    with pytest.raises(etp.EtpError):  # program reads CTL aux columns but the spec has none
        ctx.register_table_ex(prog, np.array([cprog.AUXSPEC_MAGIC, 0, 0], dtype=np.uint64))


@pytest.mark.parametrize("log_n", [1, 2, 3, 5])
def test_tiny_tables_through_prove_with_commitment(ctx, log_n):
    """Degrees below every FRI reduction and below the cap height's comfort zone (2^log_n << rate_bits vs 2^cap_height)."""
    import torch

    import eth_tx_proof_b200 as etp
    import oracle
    from eth_tx_proof_b200 import synthetic as syn

    if log_n + 1 < 4:
        with pytest.raises(etp.EtpError):  # cap_height 4 > log2(leaves): upstream asserts in MerkleTree::new
            ctx.stark_prove(etp.TABLE_FIBONACCI, *syn.fibonacci_trace(log_n))
        return
    t, pi = syn.fibonacci_trace(log_n, seed=log_n)
    assert (ctx.stark_prove(etp.TABLE_FIBONACCI, t, pi) == oracle.stark_prove(oracle.TABLE_FIBONACCI, t, pi)).all()
