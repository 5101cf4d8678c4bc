"""GPU tests of the tables with real semantics (eth_tx_proof_b200/evm_tables.py) and of the segmented code generation for large
constraint programs (csrc/cprog.h: above 16384 ops a program becomes a chain of __noinline__ segment functions; forced here on
small programs with ETP_CPROG_SEGMENT_OPS so that the segmented kernels RUN and are compared with the oracle and with the
one-function kernels).  evm_arithmetization 0.1.3's own table sources are not available offline (/root/reference/Cargo.lock:1675)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ctx():
    import eth_tx_proof_b200 as etp

    return etp.Context(0)


@pytest.mark.parametrize("n_limbs,limb_bits,log_n,segment_ops", [(4, 5, 6, 0), (16, 6, 7, 0), (4, 5, 7, 64), (16, 6, 8, 256)])
def test_arithmetic_table_on_the_device_equals_the_oracle(n_limbs, limb_bits, log_n, segment_ops, monkeypatch):
    """ADD / SUB / LT / GT / MUL on limb words with 80 range-checked limbs (one logUp lookup, chunked helpers): the device proof
    equals the oracle's word for word and the verifier accepts — with the one-function kernel and with forced segments."""
    import oracle
    import stark_verifier as V
    from eth_tx_proof_b200 import evm_tables as et

    if segment_ops:
        monkeypatch.setenv("ETP_CPROG_SEGMENT_OPS", str(segment_ops))
    else:
        monkeypatch.delenv("ETP_CPROG_SEGMENT_OPS", raising=False)
    ctx = _ctx()  # a fresh context: the per-context kernel cache is keyed by the program, not by the generated source
    try:
        prog = et.arithmetic_program(n_limbs, limb_bits)
        t, _ = et.arithmetic_trace(log_n, n_limbs, limb_bits, seed=31 + segment_ops)
        tid = ctx.register_table(prog)
        oid = oracle.register_table_ex(prog, prog.aux_spec)
        got, want = ctx.stark_prove(tid, t), oracle.stark_prove(oid, t)
        assert (np.delete(got, 1) == np.delete(want, 1)).all()
        V.verify(got, program=prog, max_queries=3)
    finally:
        ctx.close()


def test_segmented_memory_program_equals_the_builtin_table(monkeypatch):
    """The memory table's program cut into ~8 segments: proofs equal the built-in table's (and therefore the oracle's)."""
    import eth_tx_proof_b200 as etp
    from eth_tx_proof_b200 import cprog, synthetic as syn

    monkeypatch.setenv("ETP_CPROG_SEGMENT_OPS", "24")
    ctx = _ctx()
    try:
        prog = cprog.memory_program()
        tid = ctx.register_table(prog, prog.lookups)
        for log_n in (6, 10):
            t = syn.memory_trace(log_n, seed=5 + log_n)
            assert (ctx.stark_prove(tid, t)[2:] == ctx.stark_prove(etp.TABLE_MEMORY, t)[2:]).all()
    finally:
        ctx.close()
