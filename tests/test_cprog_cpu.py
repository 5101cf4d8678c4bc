"""CPU tests of the constraint-program path: the Python recorder, the wire format, the oracle's interpreter against
the oracle's built-in tables and the independent verifier, and — without any GPU — that the product library parses a
program and NVRTC-compiles its quotient kernel for sm_100a (etp_cprog_compile_check)."""
import ctypes as C

import numpy as np
import pytest

P = 0xFFFFFFFF00000001


def test_builder_hash_conses_and_encodes():
    from eth_tx_proof_b200 import cprog

    b = cprog.ProgramBuilder(3, 1, 3)
    x, y = b.lv(0), b.nv(1)
    e1 = x * y + 5
    e2 = x * y + 5
    assert e1.id == e2.id                      # shared sub-expressions are evaluated once
    b.constraint(e1 - b.pi(0))
    b.constraint(e1 - b.pi(0))                 # ...but every constraint is emitted
    p = b.build()
    assert p.n_constraints == 2 and p.words[0] == cprog.MAGIC and p.words[1] == len(p.ops)
    assert p.words.size == 8 + 2 * len(p.ops)
    out = p.evaluate([3, 0, 0], [0, 4, 0], pi=[17])
    assert [v for _, v in out] == [0, 0] and [k for k, _ in out] == [cprog.EMIT, cprog.EMIT]


def test_sample_tables_are_satisfied_by_their_traces():
    from eth_tx_proof_b200 import cprog, synthetic as syn

    assert cprog.logic_program(1).check_trace(cprog.logic_trace(4, 1)) == -1
    t = cprog.logic_trace(4, 1)
    t[cprog.logic_layout(1)["RES"], 2] += np.uint64(1)
    assert cprog.logic_program(1).check_trace(t) // 1000 == 2
    tr, pi = syn.fibonacci_trace(5)
    assert cprog.fibonacci_program().check_trace(tr, pi) == -1


def test_oracle_interprets_the_memory_program_like_its_builtin_table():
    import oracle
    import stark_verifier as V
    from eth_tx_proof_b200 import cprog, synthetic as syn

    prog = cprog.memory_program()
    tid = oracle.register_table(prog, prog.lookups)
    t = syn.memory_trace(6)
    a = oracle.stark_prove(oracle.TABLE_MEMORY, t)
    b = oracle.stark_prove(tid, t)
    assert (a[2:] == b[2:]).all() and int(b[1]) == tid
    V.verify(b, program=prog, max_queries=3)
    with pytest.raises(V.VerifyError):
        V.verify(b)                            # a registered table cannot be verified without its program


def test_oracle_multi_column_lookup_and_verifier():
    import oracle
    import stark_verifier as V
    from eth_tx_proof_b200 import cprog

    prog = cprog.rangecheck_program(5)
    tid = oracle.register_table(prog, prog.lookups)
    t = cprog.rangecheck_trace(6, 5)
    assert oracle.lib().orc_table_num_aux_columns(tid, 2) == 8 == prog.n_aux
    proof = oracle.stark_prove(tid, t)
    V.verify(proof, program=prog, max_queries=3)
    aux = oracle.lookup_helper_columns(tid, t, [11, 12])
    # Z telescopes: Z[n-1] + last term == 0 is what the wrap-around constraint enforces; first Z is 0
    assert (aux[3] [0] == 0) and (aux[7][0] == 0)


def _compile_check(words):
    import eth_tx_proof_b200 as etp

    L = etp.load_library()
    n = C.c_size_t()
    err = C.create_string_buffer(512)
    rc = L.etp_cprog_compile_check(words.ctypes.data_as(C.POINTER(C.c_uint64)), words.size, C.byref(n), err, 512)
    return rc, n.value, err.value.decode()


def test_product_compiles_programs_for_sm100a_without_a_gpu():
    from eth_tx_proof_b200 import cprog

    rc, size, err = _compile_check(cprog.memory_program().words)
    assert rc == 0 and size > 10000, err
    rc, size, err = _compile_check(cprog.logic_program(2).words)   # > 600 ops: compact-code path
    assert rc == 0 and size > 10000, err


def test_product_rejects_malformed_programs():
    from eth_tx_proof_b200 import cprog

    w = cprog.fibonacci_program().words.copy()
    bad = w.copy(); bad[0] = 1
    assert _compile_check(bad)[0] != 0
    bad = w[:-1].copy()
    assert _compile_check(bad)[0] != 0
    bad = w.copy(); bad[8] = np.uint64(cprog.LV | (7 << 8))        # column 7 of a 2-column table
    rc, _, err = _compile_check(bad)
    assert rc != 0 and "malformed" in err
    bad = w.copy(); bad[7] = 99                                     # n_constraints does not match
    assert _compile_check(bad)[0] != 0


@pytest.mark.parametrize("n_cols,n_lookup", [(9, 0), (37, 3), (22, 4), (80, 8)])
def test_shape_tables_are_satisfied_and_prove(n_cols, n_lookup):
    """The shape-only stand-ins for the evm_arithmetization tables (cprog.shape_program): the generated trace satisfies the
    program, a tampered one does not, and the oracle's proof is accepted by the independent verifier."""
    import oracle
    import stark_verifier as V
    from eth_tx_proof_b200 import cprog

    prog = cprog.shape_program(n_cols, n_lookup)
    t = cprog.shape_trace(6, n_cols, n_lookup)
    own = cprog.shape_program(n_cols, n_lookup, emit_lookups=False)
    assert own.check_trace(t) == -1
    L = cprog.shape_layout(n_cols, n_lookup)
    bad = t.copy()
    bad[L["GROUP"] + 3, 7] += np.uint64(1)
    assert own.check_trace(bad) // 1000 == 7
    if n_cols <= 40:
        tid = oracle.register_table(prog, prog.lookups)
        V.verify(oracle.stark_prove(tid, t), program=prog, max_queries=2)


def test_evm_table_shapes_compile_for_sm100a():
    """Every table of the synthetic transaction job compiles (keccak: 2400 columns, 600 constraints)."""
    from eth_tx_proof_b200 import cprog

    for name, (cols, lk) in cprog.EVM_TABLE_SHAPES.items():
        prog = cprog.shape_program(cols, lk)
        assert prog.n_trace == cols
        rc, size, err = _compile_check(prog.words)
        assert rc == 0 and size > 10000, (name, err)
    assert set(cprog.TX_TABLE_DEGREE_BITS) == set(cprog.EVM_TABLE_SHAPES) | {"logic", "memory"}


def _generated_source(words):
    import eth_tx_proof_b200 as etp

    L = etp.load_library()
    ptr = words.ctypes.data_as(C.POINTER(C.c_uint64))
    n = L.etp_cprog_generate_cuda(ptr, words.size, None, 0)
    assert n > 0
    buf = C.create_string_buffer(n + 1)
    assert L.etp_cprog_generate_cuda(ptr, words.size, buf, n + 1) == n
    return buf.value.decode()


def _interpret_generated(src, lv, nv, la, na, pi, ch):
    """Re-interprets the CUDA text the library generates (one kernel, or segment functions called in order) with Python integers
    -> [(kind, value)] in the order the consumer would see them."""
    import re

    val = re.compile(r"^\s*const uint64_t v(\d+) = (.+);$")
    emit = re.compile(r"^\s*r\.cs\.(constraint|transition|first_row|last_row)\(v(\d+)\);$")
    kinds = {"constraint": 10, "transition": 11, "first_row": 12, "last_row": 13}
    funcs, cur, order = {}, None, []
    for line in src.splitlines():
        m = re.match(r"^static __device__ __noinline__ void (etp_seg_\d+)\(", line)
        if m:
            cur = funcs.setdefault(m.group(1), [])
            continue
        if "etp_cprog_quotient(" in line:
            cur = funcs.setdefault("kernel", [])
            continue
        if line.strip() == "}":
            cur = None
            continue
        m = re.match(r"^\s*(etp_seg_\d+)\(q, r\);$", line)
        if m and cur is funcs.get("kernel"):
            order.append(m.group(1))
            continue
        if cur is not None and (val.match(line) or emit.match(line)):
            cur.append(line)
    out = []

    def run(lines):
        v = {}
        for line in lines:
            m = val.match(line)
            if m:
                k, rhs = int(m.group(1)), m.group(2)
                g = re.match(r"^gl::(add|sub|mul)\(v(\d+), v(\d+)\)$", rhs)
                ld = re.match(r"^r\.(lv|nv|la|na)\(q, (\d+)\)$", rhs)
                if g:
                    a, b = v[int(g.group(2))], v[int(g.group(3))]  # KeyError = a value used in a segment that does not define it
                    v[k] = {"add": a + b, "sub": a - b, "mul": a * b}[g.group(1)] % P
                elif ld:
                    v[k] = {"lv": lv, "nv": nv, "la": la, "na": na}[ld.group(1)][int(ld.group(2))] % P
                elif rhs.startswith("q.pi["):
                    v[k] = pi[int(rhs[5:-1])] % P
                elif rhs.startswith("q.lookup_ch["):
                    v[k] = ch[int(rhs[12:-1])] % P
                else:
                    v[k] = int(rhs.replace("ULL", ""), 16) % P
            else:
                m = emit.match(line)
                out.append((kinds[m.group(1)], v[int(m.group(2))]))

    run(funcs["kernel"])
    for name in order:
        run(funcs[name])
    return out, len(order)


@pytest.mark.parametrize("segment_ops", [0, 25, 300])
def test_generated_cuda_is_the_program_also_when_segmented(segment_ops, monkeypatch):
    """Large programs are emitted as a chain of segment functions (cprog.h: ptxas cannot digest a 50 k-op table as one
    function): the generated text, re-interpreted with Python integers on random rows, produces exactly the program's
    emissions in the program's order — for the one-function form and for forced segment sizes — and it compiles for sm_100a."""
    from eth_tx_proof_b200 import cprog, evm_tables as et

    if segment_ops:
        monkeypatch.setenv("ETP_CPROG_SEGMENT_OPS", str(segment_ops))
    else:
        monkeypatch.delenv("ETP_CPROG_SEGMENT_OPS", raising=False)
    rng = np.random.default_rng(segment_ops)
    for prog in (cprog.memory_program(), et.arithmetic_program(4, 5, with_ctl=True), cprog.fibonacci_program()):
        rnd = lambda n: [int(x) % P for x in rng.integers(0, 2**64, max(n, 1), dtype=np.uint64)]
        lv, nv, la, na = rnd(prog.n_trace), rnd(prog.n_trace), rnd(prog.n_aux), rnd(prog.n_aux)
        pi, ch = rnd(prog.n_pi), rnd(max(prog.n_ch, 8))
        want = [(k, v % P) for k, v in prog.evaluate(lv, nv, la, na, pi, ch)]
        got, n_seg = _interpret_generated(_generated_source(prog.words), lv, nv, la, na, pi, ch)
        assert got == want
        segmented = bool(segment_ops and len(prog.ops) > segment_ops)
        assert (n_seg >= 1) == segmented and (not segmented or n_seg >= len(prog.ops) // (4 * segment_ops))
        rc, size, err = _compile_check(prog.words)
        assert rc == 0 and size > 1000, err
