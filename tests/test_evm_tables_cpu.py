"""CPU tests of the tables with real semantics (eth_tx_proof_b200/evm_tables.py: own layouts of an arithmetic table — ADD, SUB,
LT, GT, MUL on 256-bit words through 16-bit limbs, addcy without carry columns, logUp range checks — and of a bit-level
Keccak-f[1600] round table; evm_arithmetization 0.1.3's own sources are not available offline, /root/reference/Cargo.lock:1675).
The traces are checked against independent computations (Python integers, hashlib's SHA-3), the constraint programs hold on them
and only on them, the oracle proves the tables and the independent verifier accepts the proofs; the arithmetic table is the looked
side of a cross-table lookup whose looking table lists operations and results."""
import hashlib

import numpy as np
import pytest

import oracle
import stark_verifier as V
from eth_tx_proof_b200 import cprog
from eth_tx_proof_b200 import evm_tables as et

P = cprog.P


@pytest.mark.parametrize("n_limbs,limb_bits,log_n", [(4, 5, 6), (16, 6, 7)])
def test_arithmetic_table_semantics_constraints_and_proof(n_limbs, limb_bits, log_n):
    t, ops = et.arithmetic_trace(log_n, n_limbs, limb_bits)
    M = 1 << (n_limbs * limb_bits)
    seen = set()
    for op, a, b, r in ops:
        want = {et.OP_ADD: (a + b) % M, et.OP_SUB: (a - b) % M, et.OP_LT: int(a < b), et.OP_GT: int(a > b), et.OP_MUL: (a * b) % M}[op]
        assert r == want
        seen.add(op)
    assert seen == {0, 1, 2, 3, 4}
    own = et.arithmetic_program(n_limbs, limb_bits, emit_lookups=False)
    assert own.check_trace(t) == -1
    L = et.arithmetic_layout(n_limbs, limb_bits)
    rows = {op: next(r for r in range(t.shape[1]) if t[L["FLAG"] + op, r]) for op in range(5)}
    for op, col in ((et.OP_ADD, L["C"]), (et.OP_SUB, L["C"] + 1), (et.OP_LT, L["CY"]), (et.OP_GT, L["CY"]), (et.OP_MUL, L["C"] + n_limbs - 1)):
        bad = t.copy()
        bad[col, rows[op]] = np.uint64(int(bad[col, rows[op]]) ^ 1)  # a wrong result limb / comparison bit
        assert own.check_trace(bad) // 1000 == rows[op]
    prog = et.arithmetic_program(n_limbs, limb_bits)
    tid = oracle.register_table_ex(prog, prog.aux_spec)
    V.verify(oracle.stark_prove(tid, t), program=prog, max_queries=2)
    # a limb outside the range: the row's arithmetic can be made consistent, the logUp range check cannot
    bad = t.copy()
    r = rows[et.OP_ADD]
    W = 1 << limb_bits
    bad[L["C"], r] = np.uint64(int(bad[L["C"], r]) + W)       # C_0 + 2^w ...
    bad[L["C"] + 1, r] = np.uint64((int(bad[L["C"] + 1, r]) - 1) % P)  # ... and C_1 - 1: the same integer, limbs out of range
    with pytest.raises((V.VerifyError, RuntimeError)):
        V.verify(oracle.stark_prove(tid, bad), program=prog, max_queries=2)


def test_arithmetic_table_as_the_looked_side_of_a_ctl():
    """A looking table lists (opcode, A, B, C, CY) tuples with a filter; the arithmetic table opens the same tuple on its active
    rows: both proofs verify on one transcript and the CTL sums match; a looking row with a wrong result breaks the sum."""
    from test_ctl_oracle import verify_all

    n_limbs, limb_bits, log_n = 4, 5, 6
    L = et.arithmetic_layout(n_limbs, limb_bits)
    t, _ = et.arithmetic_trace(log_n, n_limbs, limb_bits)
    width = 2 + 3 * n_limbs + 1  # filter, opcode, A, B, C, CY

    def looking_table(corrupt=False):
        n = t.shape[1]
        lt = np.zeros((width, n), dtype=np.uint64)
        active = sum(t[L["FLAG"] + k] for k in range(5)).astype(np.uint64)
        lt[0] = active
        lt[1] = sum(t[L["FLAG"] + k] * np.uint64(k + 1) for k in range(5))
        for i in range(n_limbs):
            lt[2 + i], lt[2 + n_limbs + i], lt[2 + 2 * n_limbs + i] = t[L["A"] + i], t[L["B"] + i], t[L["C"] + i]
        lt[2 + 3 * n_limbs] = t[L["CY"]]
        lt = lt[:, ::-1].copy()  # another row order: a lookup is a multiset relation
        if corrupt:
            r = int(np.nonzero(lt[0])[0][0])
            lt[2 + 2 * n_limbs, r] = np.uint64(int(lt[2 + 2 * n_limbs, r]) ^ 1)
        b = cprog.ProgramBuilder(width, 0, 3)
        b.constraint(b.lv(0) * (b.lv(0) - 1))
        for k in range(cprog.NUM_CHALLENGES):
            b.add_ctl_z(k, [(list(range(1, width)), cprog.Filter(constants=[cprog.Column.single(0)]))])
        b.emit_lookup_constraints()
        b.emit_ctl_constraints()
        return b.build(), lt

    def prove(tables):
        tids = [oracle.register_table_ex(p, p.aux_spec) for _, p, _ in tables]
        batches = [oracle.Batch.from_values(tr, 1, 4) for _, _, tr in tables]
        ch = oracle.HostChallenger()
        for bb in batches:
            ch.observe(bb.cap)
        ctl_ch = ch.get_n(4)
        proofs = []
        for tid, (_, _, tr), bb in zip(tids, tables, batches):
            ch.compact()
            proofs.append(oracle.prove_with_commitment(tid, tr, bb, ch, ctl_ch))
        return proofs, [bb.cap for bb in batches]

    arith = et.arithmetic_program(n_limbs, limb_bits, with_ctl=True)
    ctls = [([0], 1)]
    lp, lt = looking_table()
    tables = [("cpu_ops", lp, lt), ("arithmetic", arith, t)]
    proofs, caps = prove(tables)
    verify_all(tables, ctls, proofs, caps, max_queries=2)
    lp, lt = looking_table(corrupt=True)
    tables = [("cpu_ops", lp, lt), ("arithmetic", arith, t)]
    proofs, caps = prove(tables)
    with pytest.raises(V.VerifyError, match="Cross-table lookup"):
        verify_all(tables, ctls, proofs, caps, max_queries=1)


def _sha3_256(msg: bytes) -> bytes:
    rate = 136
    m = bytearray(msg) + b"\x06"
    m += b"\x00" * (-len(m) % rate)
    m[-1] |= 0x80
    st = [0] * 25
    for off in range(0, len(m), rate):
        for i in range(rate // 8):
            st[i] ^= int.from_bytes(m[off + 8 * i:off + 8 * i + 8], "little")
        st = et.keccak_f(st)
    return b"".join(x.to_bytes(8, "little") for x in st[:4])


def test_keccak_table_against_sha3_constraints_and_proof():
    """keccak_f (the definition the trace generator follows) reproduces hashlib's SHA3-256; the 5530-column trace — one round per
    row — satisfies the 7157 constraints of the table's program, a flipped state bit does not; the oracle proves the table and the
    verifier accepts; the input / output limbs the CTL ports open are the permutation's."""
    for msg in (b"", b"abc", bytes(range(200))):
        assert _sha3_256(msg) == hashlib.sha3_256(msg).digest()
    assert et.KECCAK_RC[0] == 1 and et.KECCAK_RC[23] == 0x8000000080008008 and et.KECCAK_ROT[1][0] == 1 and et.KECCAK_ROT[3][2] == 25 and et.KECCAK_ROT[2][3] == 15
    t, io = et.keccak_trace(6)
    L = et.keccak_layout()
    assert t.shape == (L["cols"], 64) and len(io) == 2
    for pid, lanes_in, lanes_out in io:
        assert et.keccak_f(lanes_in) == lanes_out
        first, last = (pid - 1) * 24, (pid - 1) * 24 + 23
        limb = lambda base, row, k: sum(int(t[base + 32 * k + j, row]) << j for j in range(32))
        assert [limb(L["A"], first, k) for k in range(50)] == [(lanes_in[k // 2] >> (32 * (k % 2))) & 0xFFFFFFFF for k in range(50)]
        assert [limb(L["OUT"], last, k) for k in range(50)] == [(lanes_out[k // 2] >> (32 * (k % 2))) & 0xFFFFFFFF for k in range(50)]
    prog = et.keccak_program()
    assert (prog.n_trace, prog.n_constraints, prog.degree) == (5530, 7157, 3)
    assert prog.check_trace(t) == -1
    bad = t.copy()
    bad[L["AP"] + 700, 9] ^= np.uint64(1)
    assert prog.check_trace(bad) != -1
    tid = oracle.register_table_ex(prog, prog.aux_spec)
    V.verify(oracle.stark_prove(tid, t), program=prog, max_queries=2)
    with pytest.raises(V.VerifyError):
        V.verify(oracle.stark_prove(tid, bad), program=prog, max_queries=2)


def test_keccak256_of_messages_proven_by_two_tables_and_two_ctls():
    """Keccak-256 (Ethereum's hash) of short messages: the message table pads the block (pad10*1, byte range checks by logUp) and
    looks (id, 50 input limbs) and (id, 50 output limbs) up in the Keccak-f table — upstream's keccak_sponge -> keccak CTL pair.
    The digest of the empty message is the well-known c5d246...a470; both tables prove on one transcript, the verifier accepts
    and the CTL sums match; a wrong digest limb keeps both table proofs valid and breaks the cross-table lookup; a wrong padding
    byte violates the message table's own constraints."""
    from test_ctl_oracle import verify_all

    msgs = [b"", b"abc", bytes(range(135)), b"The quick brown fox jumps over the lazy dog"]
    tables, ctls, digests = et.keccak256_system(msgs)
    assert digests[0].hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert digests == [et.keccak256(m) for m in msgs]
    L = et.keccak256_layout()
    own = et.keccak256_program(with_ctl=False, emit_lookups=False)
    assert own.check_trace(tables[0][2]) == -1

    def prove(tabs):
        tids = [oracle.register_table_ex(p, p.aux_spec) for _, p, _ in tabs]
        batches = [oracle.Batch.from_values(tr, 1, 4) for _, _, tr in tabs]
        ch = oracle.HostChallenger()
        for bb in batches:
            ch.observe(bb.cap)
        ctl_ch = ch.get_n(4)
        proofs = []
        for tid, (_, _, tr), bb in zip(tids, tabs, batches):
            ch.compact()
            proofs.append(oracle.prove_with_commitment(tid, tr, bb, ch, ctl_ch))
        return proofs, [bb.cap for bb in batches]

    proofs, caps = prove(tables)
    zs = verify_all(tables, ctls, proofs, caps, max_queries=2)
    assert [len(z) for z in zs] == [4, 4]
    bad = tables[0][2].copy()
    bad[L["OUT"] + 3, 1] ^= np.uint64(1)  # a digest limb of message 1
    bad_tables = [(tables[0][0], tables[0][1], bad), tables[1]]
    proofs, caps = prove(bad_tables)
    with pytest.raises(V.VerifyError, match="Cross-table lookup"):
        verify_all(bad_tables, ctls, proofs, caps, max_queries=1)
    bad = tables[0][2].copy()
    bad[L["BYTE"] + 3, 1] = 2  # message 1 = b"abc": byte 3 must be the 0x01 domain byte
    assert own.check_trace(bad) // 1000 == 1
    bad = tables[0][2].copy()
    bad[L["BYTE"] + 135, 0] = 0  # the final 0x80 bit of pad10*1
    assert own.check_trace(bad) // 1000 == 0


def test_real_tables_through_the_recursion_layers():
    """cpu_ops -> arithmetic (real semantics, one CTL) proven on one transcript, then wrapped, shrunk and rooted
    (stark_circuit.transaction_recursion_plan with the oracle as the circuit prover): the root circuit's proof is accepted by the
    Python verifier.  tools/real_tables_recursion_cpu.py runs the same with the Keccak-256 tables and the layers above."""
    import types

    import plonk_verifier
    from eth_tx_proof_b200 import stark_circuit as sc
    from test_circuit_cpu import _words_from_oracle_proof

    n_limbs, limb_bits = 4, 5
    L = et.arithmetic_layout(n_limbs, limb_bits)
    at, _ = et.arithmetic_trace(6, n_limbs, limb_bits)
    width = 2 + 3 * n_limbs + 1
    lt = np.zeros((width, at.shape[1]), dtype=np.uint64)
    lt[0] = sum(at[L["FLAG"] + k] for k in range(5)).astype(np.uint64)
    lt[1] = sum(at[L["FLAG"] + k] * np.uint64(k + 1) for k in range(5))
    for i in range(n_limbs):
        lt[2 + i], lt[2 + n_limbs + i], lt[2 + 2 * n_limbs + i] = at[L["A"] + i], at[L["B"] + i], at[L["C"] + i]
    lt[2 + 3 * n_limbs] = at[L["CY"]]
    b = cprog.ProgramBuilder(width, 0, 3)
    b.constraint(b.lv(0) * (b.lv(0) - 1))
    for k in range(cprog.NUM_CHALLENGES):
        b.add_ctl_z(k, [(list(range(1, width)), cprog.Filter(constants=[cprog.Column.single(0)]))])
    b.emit_lookup_constraints()
    b.emit_ctl_constraints()
    tables = [("cpu_ops", b.build(), lt), ("arithmetic", et.arithmetic_program(n_limbs, limb_bits, with_ctl=True), at)]
    ctls = [([0], 1)]
    tids = [oracle.register_table_ex(p, p.aux_spec) for _, p, _ in tables]
    batches = [oracle.Batch.from_values(t, 1, 4) for _, _, t in tables]
    ch = oracle.HostChallenger()
    for bb in batches:
        ch.observe(bb.cap)
    ctl_ch = ch.get_n(4)
    proofs, states = [], []
    for tid, (_, _, t), bb in zip(tids, tables, batches):
        states.append(ch.compact())
        proofs.append(oracle.prove_with_commitment(tid, t, bb, ch, ctl_ch))
    digest = [4, 3, 2, 1]

    def circuit_prove(c, w, p):
        pr = oracle.circuit_prove(c, w, p, digest)
        plonk_verifier.verify(pr, c, pr["constants_sigmas_cap"], digest, max_queries=1)
        return types.SimpleNamespace(c=c, digest=digest, constants_sigmas_cap=pr["constants_sigmas_cap"]), _words_from_oracle_proof(c, pr, p)

    plan = sc.transaction_recursion_plan(tables, ctls, types.SimpleNamespace(stark_proofs=proofs, init_challenger_states=states, ctl_challenges=ctl_ch),
                                         circuit_prove, max_queries=1)
    assert [s["kind"] for s in plan] == ["wrapper", "shrink", "wrapper", "shrink", "root"]
    assert plan[-1]["public_inputs"][-4:] == [int(x) for x in ctl_ch]


def test_byte_packing_table_with_thirty_two_looking_entries():
    """Byte packing (own layout of the upstream design): one operation per row, bytes past the length are zero, every byte
    range-checked; the table is looked by a CPU-side port on (is_read, address, length, timestamp, eight value limbs) and looks
    into a memory-side port with 32 column sets in ONE Z per challenge (chunked helper columns) — address VIRT + LEN - 1 - i as a
    linear-combination Column, filter [i < LEN] as a sum of flags.  Proven on one transcript, CTL sums verified; a memory access
    with another byte breaks the lookup, a non-zero byte past the length breaks the table's own constraints."""
    from test_ctl_oracle import verify_all

    tables, ctls, ops = et.byte_packing_system()
    L = et.byte_packing_layout()
    t = tables[1][2]
    for r, (is_read, ctx, seg, virt, ln, ts, value) in enumerate(ops):
        assert sum(int(t[L["BYTE"] + i, r]) << (8 * i) for i in range(32)) == value < (1 << (8 * ln))
    assert {o[4] for o in ops} >= {1, 32}
    own = et.byte_packing_program(with_ctl=False, emit_lookups=False)
    assert own.check_trace(t) == -1
    bad = t.copy()
    r = next(i for i, o in enumerate(ops) if o[4] < 32)
    bad[L["BYTE"] + ops[r][4], r] = 7
    assert own.check_trace(bad) // 1000 == r
    prog = tables[1][1]
    assert len(prog.ctl_zs) == 4 and [len(sets) for _, sets in prog.ctl_zs] == [1, 1, 32, 32] and prog.n_ctl_helper_cols == 32

    def prove(tabs):
        tids = [oracle.register_table_ex(p, p.aux_spec) for _, p, _ in tabs]
        batches = [oracle.Batch.from_values(tr, 1, 4) for _, _, tr in tabs]
        ch = oracle.HostChallenger()
        for bb in batches:
            ch.observe(bb.cap)
        ctl_ch = ch.get_n(4)
        proofs = []
        for tid, (_, _, tr), bb in zip(tids, tabs, batches):
            ch.compact()
            proofs.append(oracle.prove_with_commitment(tid, tr, bb, ch, ctl_ch))
        return proofs, [bb.cap for bb in batches]

    proofs, caps = prove(tables)
    zs = verify_all(tables, ctls, proofs, caps, max_queries=2)
    assert [len(z) for z in zs] == [2, 4, 2]
    mem = tables[2][2].copy()
    mem[5, 3] = np.uint64(int(mem[5, 3]) ^ 1)  # the byte of one memory access
    bad_tables = [tables[0], tables[1], (tables[2][0], tables[2][1], mem)]
    proofs, caps = prove(bad_tables)
    with pytest.raises(V.VerifyError, match="Cross-table lookup 1"):
        verify_all(bad_tables, ctls, proofs, caps, max_queries=1)


def test_real_tables_compile_for_sm_100a():
    """The product's NVRTC path accepts the real tables' programs (no GPU needed for the compile): full-width arithmetic with its
    CTL port, byte packing with its 32 looking sets, the Keccak-256 message table.  (The Keccak-f table itself — 56 k ops — goes
    through the segmented code generation and takes minutes: tests/test_cprog_cpu.py covers that path on small programs.)"""
    from test_cprog_cpu import _compile_check

    for prog in (et.arithmetic_program(16, 16, with_ctl=True), et.byte_packing_program(), et.keccak256_program()):
        rc, size, err = _compile_check(prog.words)
        assert rc == 0 and size > 100000, err


def test_seven_real_tables_wired_like_all_stark():
    """evm_tables.real_transaction_system: arithmetic, byte packing, cpu (dispatcher), keccak, keccak sponge (Keccak-256 messages),
    logic, memory (port) — upstream's table order and its seven cross-table lookups — proven on ONE transcript
    (prove_with_traces' shape) and verified including every CTL sum.  tools/real_tables_recursion_cpu.py takes the same system
    through the wrapper / shrink / root / aggregation / block circuits (profiles/r02_real_tables_recursion_cpu.log)."""
    from test_ctl_oracle import verify_all

    tables, ctls = et.real_transaction_system()
    assert tuple(n for n, _, _ in tables) == et.REAL_TABLE_ORDER and ctls == et.REAL_CTLS
    tids = [oracle.register_table_ex(p, p.aux_spec) for _, p, _ in tables]
    batches = [oracle.Batch.from_values(tr, 1, 4) for _, _, tr in tables]
    ch = oracle.HostChallenger()
    for bb in batches:
        ch.observe(bb.cap)
    ctl_ch = ch.get_n(4)
    proofs = []
    for tid, (_, _, tr), bb in zip(tids, tables, batches):
        ch.compact()
        proofs.append(oracle.prove_with_commitment(tid, tr, bb, ch, ctl_ch))
    zs = verify_all(tables, ctls, proofs, [bb.cap for bb in batches], max_queries=1)
    assert [len(z) for z in zs] == [2, 4, 8, 4, 6, 2, 2]
    # an operation the cpu table claims but the arithmetic table did not perform: every table proof stands, CTL 0 fails
    bad = tables[2][2].copy()
    r = int(np.nonzero(bad[0])[0][0])
    bad[4 + 1, r] = np.uint64(int(bad[4 + 1, r]) ^ 1)
    bad_tables = list(tables)
    bad_tables[2] = (tables[2][0], tables[2][1], bad)
    batches = [oracle.Batch.from_values(tr, 1, 4) for _, _, tr in bad_tables]
    ch = oracle.HostChallenger()
    for bb in batches:
        ch.observe(bb.cap)
    ctl_ch = ch.get_n(4)
    proofs = []
    for tid, (_, _, tr), bb in zip(tids, bad_tables, batches):
        ch.compact()
        proofs.append(oracle.prove_with_commitment(tid, tr, bb, ch, ctl_ch))
    with pytest.raises(V.VerifyError, match="Cross-table lookup 0"):
        verify_all(bad_tables, ctls, proofs, [bb.cap for bb in batches], max_queries=1)
