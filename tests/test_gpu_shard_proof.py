"""GPU parity tests of the proof of ONE column-split table (VERDICT r1 item 8 / SURVEY.md 8(e)): quotient, openings and
FRI over a trace whose columns live on different GPUs must give, word for word, the proof one GPU makes of the same trace
(starky 0.4.0 src/prover.rs prove -> prove_with_commitment; /root/reference/Cargo.lock:4529, reached from
/root/reference/ops/src/lib.rs:52) — and therefore the oracle's.  Single-GPU runs drive G shards from G threads of one
process (a context per shard, all on device 0, peers wired by pointer); with >= 2 GPUs the one-process-per-GPU protocol
(CUDA IPC + NVLink peer loads inside the quotient / combination kernels) runs under torchrun."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _split_proof(world, make_table, trace, pi, leader=None, challenger=None, ctl_challenges=None, want_cap=None):
    """Runs the protocol with `world` threads (one context + one shard each); returns the leader's proof.  `challenger`
    (advanced in place) and `ctl_challenges`: the prove_with_commitment form."""
    import eth_tx_proof_b200 as etp
    from eth_tx_proof_b200 import parallel

    cols, log_n = trace.shape[0], int(np.log2(trace.shape[1]))
    ctxs = [etp.Context(0) for _ in range(world)]
    tables = [make_table(c) for c in ctxs]
    shards = [etp.BatchShard(ctxs[r], cols, log_n, 1, 4, r, world) for r in range(world)]
    for r, s in enumerate(shards):
        c0, c1 = parallel.column_split_plan(cols, 2 << log_n, 4, r, world)["cols"]
        s.transform_values(np.ascontiguousarray(trace[c0:c1]))
    for s in shards:
        for r, t in enumerate(shards):
            if r != s.rank and t.num_local_cols:
                s.set_peer(r, t.lde_ptr)
    cap = parallel.assemble_cap([s.commit_rows() for s in shards])
    if want_cap is not None:
        assert (cap == want_cap).all()
    lead = world - 1 if leader is None else leader
    comm = parallel.ThreadComm(world)
    out, err = [None] * world, [None] * world

    def work(r):
        try:
            out[r] = parallel.prove_column_split(shards[r], tables[r], cap, pi, leader=leader, comm=comm.rank(r),
                                                 challenger=challenger if r == lead else None, ctl_challenges=ctl_challenges)
        except BaseException as e:  # noqa: BLE001
            err[r] = e
            comm._barrier.abort()

    ths = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    first = next((e for e in err if e is not None and not isinstance(e, threading.BrokenBarrierError)), None)
    if first is not None:
        raise first
    assert all(out[r] is None for r in range(world) if r != lead)
    proof = out[lead]
    del shards
    for c in ctxs:
        c.close()
    return proof


@pytest.fixture(scope="module")
def ctx():
    import eth_tx_proof_b200 as etp

    c = etp.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("log_n", [6, 9])
def test_split_fibonacci_proof_equals_single_gpu_and_oracle(ctx, world, log_n):
    import eth_tx_proof_b200 as etp
    import oracle
    from eth_tx_proof_b200 import synthetic as syn

    trace, pi = syn.fibonacci_trace(log_n, seed=2)
    want = ctx.stark_prove(etp.TABLE_FIBONACCI, trace, pi)
    got = _split_proof(world, lambda c: etp.TABLE_FIBONACCI, trace, list(pi))
    assert got.shape == want.shape and (got == want).all()
    assert (got == oracle.stark_prove(oracle.TABLE_FIBONACCI, trace, pi)).all()


@pytest.mark.parametrize("world,leader", [(2, None), (2, 0), (4, 1), (8, None)])
@pytest.mark.parametrize("shape", ["shape21", "logic68", "shape100"])
def test_split_registered_table_proof_equals_single_gpu(ctx, world, leader, shape):
    """Program-defined tables (NVRTC): the column-split variant of the compiled quotient kernel reads every column through
    the pointer table.  21 columns over 8 ranks: ranks 3..7 own none; 100 columns: ragged last shard."""
    from eth_tx_proof_b200 import cprog

    log_n = 8
    if shape == "logic68":
        prog, trace = cprog.logic_program(1), cprog.logic_trace(log_n, 1, seed=6)
    else:
        cols = int(shape[5:])
        prog, trace = cprog.shape_program(cols, 0), cprog.shape_trace(log_n, cols, 0, seed=5)
    want = ctx.stark_prove(ctx.register_table(prog), trace)
    got = _split_proof(world, lambda c: c.register_table(prog), trace, [], leader=leader)
    # the table id in the header is context-local (both are the first registered table of their context or not): compare past it
    assert got.shape == want.shape
    assert (got[2:] == want[2:]).all() and got[0] == want[0]


def test_split_proof_is_accepted_by_the_verifier(ctx):
    from eth_tx_proof_b200 import cprog

    import stark_verifier

    prog, trace = cprog.shape_program(21, 0), cprog.shape_trace(10, 21, 0, seed=9)
    got = _split_proof(4, lambda c: c.register_table(prog), trace, [])
    stark_verifier.verify(got, program=prog)


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("log_n", [7, 10])
def test_split_memory_table_with_lookups_equals_single_gpu_and_oracle(ctx, world, log_n):
    """The memory-shaped table (21 columns, range-check logUp: 2 x (helper, Z) auxiliary polynomials) — the table that
    outgrows one GPU upstream.  The leader recovers the three trace columns the lookup reads from the mapped LDE."""
    import eth_tx_proof_b200 as etp
    import oracle
    from eth_tx_proof_b200 import synthetic as syn

    trace = syn.memory_trace(log_n, seed=4)
    want = ctx.stark_prove(etp.TABLE_MEMORY, trace)
    got = _split_proof(world, lambda c: etp.TABLE_MEMORY, trace, [])
    assert got.shape == want.shape and (got == want).all()
    if log_n <= 7:
        assert (got == oracle.stark_prove(oracle.TABLE_MEMORY, trace)).all()


@pytest.mark.parametrize("world", [2, 8])
def test_split_registered_tables_with_general_lookups(ctx, world):
    """Program-defined tables with lookups: the memory program (plain columns) and a shape table with 8 range-checked limbs
    (chunked helpers); the compacted spec renumbers the columns the leader recovered."""
    from eth_tx_proof_b200 import cprog, synthetic as syn

    for prog, trace in ((cprog.memory_program(), syn.memory_trace(8, seed=5)),
                        (cprog.shape_program(40, 8), cprog.shape_trace(8, 40, 8, seed=6))):
        want = ctx.stark_prove(ctx.register_table(prog), trace)
        got = _split_proof(world, lambda c: c.register_table(prog), trace, [])
        assert got.shape == want.shape and (got[2:] == want[2:]).all()


@pytest.mark.parametrize("world", [2, 4])
def test_split_ctl_system_equals_single_gpu_table_by_table(ctx, world):
    """prove_with_traces on the three-table CTL system (filtered lookups, linear-combination and next-row Columns, CTL
    helper columns + Z, ctl_zs_first and the third FRI batch): every table column-split, one challenger threaded through
    all of them — proofs and challenger states equal the single-GPU prover's, and the verifier accepts the CTL sums."""
    import torch

    import eth_tx_proof_b200 as etp
    from eth_tx_proof_b200 import cprog, prover
    from test_ctl_oracle import verify_all

    tables, ctls = cprog.ctl_demo_tables(7, 6, 5)
    tids = [ctx.register_table(p) for _, p, _ in tables]
    devs = [torch.from_numpy(np.ascontiguousarray(t).view(np.int64)).cuda() for _, _, t in tables]
    torch.cuda.synchronize()
    traces_dev = [(d.data_ptr(), t.shape[1], t.shape[0], int(t.shape[1]).bit_length() - 1) for d, (_, _, t) in zip(devs, tables)]
    want = prover.prove_with_traces(ctx, tids, traces_dev)
    ch = etp.Challenger()
    for cap in want.trace_caps:
        ch.observe_cap(cap)
    ctl_ch = ch.get_n_challenges(4)
    assert (ctl_ch == want.ctl_challenges).all()
    proofs = []
    for k, (_, prog, trace) in enumerate(tables):
        assert (ch.compact() == want.init_challenger_states[k]).all()
        got = _split_proof(world, lambda c: c.register_table(prog), trace, [], challenger=ch, ctl_challenges=ctl_ch,
                           want_cap=want.trace_caps[k])
        assert got.shape == want.stark_proofs[k].shape and (got[2:] == want.stark_proofs[k][2:]).all(), f"table {k}"
        proofs.append(got)
    zs = verify_all(tables, ctls, proofs, want.trace_caps, max_queries=2)
    assert [len(z) for z in zs] == [2, 2, 2]


def test_split_ctl_table_needs_ctl_challenges(ctx):
    import eth_tx_proof_b200 as etp
    from eth_tx_proof_b200 import cprog

    tables, _ = cprog.ctl_demo_tables(5, 4, 4)
    _, prog, trace = tables[1]
    with pytest.raises(etp.EtpError, match="CTL"):
        _split_proof(2, lambda c: c.register_table(prog), trace, [])


def test_split_lookup_with_vanishing_denominator_fails_loudly(ctx):
    """challenge + table value = 0 on some row: upstream panics in batch_multiplicative_inverse; here ETP_ERR_PROOF, on the
    split path too (the zero flag is checked after the auxiliary columns)."""
    import eth_tx_proof_b200 as etp
    from eth_tx_proof_b200 import cprog

    tables, _ = cprog.ctl_demo_tables(5, 4, 4)
    _, prog, trace = tables[1]
    # rom: combine = KEY + beta * VAL + gamma; gamma = -(KEY + beta * VAL) at row 3 makes that row's denominator vanish
    P = 0xFFFFFFFF00000001
    beta = 5
    gamma = (-(int(trace[0, 3]) + beta * int(trace[1, 3]))) % P
    ch = etp.Challenger()
    with pytest.raises(etp.EtpError, match="denominator"):
        _split_proof(2, lambda c: c.register_table(prog), trace, [], challenger=ch, ctl_challenges=[beta, gamma, beta, gamma])


def test_split_proof_of_an_invalid_trace_fails_like_the_single_gpu_prover(ctx):
    """A trace that violates the constraints: the quotient is not a polynomial of the expected degree — both provers
    either fail or produce a proof the verifier rejects; they must behave identically."""
    import eth_tx_proof_b200 as etp
    from eth_tx_proof_b200 import synthetic as syn

    trace, pi = syn.fibonacci_trace(7, seed=2)
    trace = trace.copy()
    trace[1, 5] ^= np.uint64(1)
    try:
        want = ctx.stark_prove(etp.TABLE_FIBONACCI, trace, pi)
    except etp.EtpError as e:
        with pytest.raises(etp.EtpError, match=str(e)[:20]):
            _split_proof(2, lambda c: etp.TABLE_FIBONACCI, trace, list(pi))
        return
    got = _split_proof(2, lambda c: etp.TABLE_FIBONACCI, trace, list(pi))
    assert (got == want).all()


def _n_gpus():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_column_split_proof_across_gpus(world):
    """One process per GPU: the leader's quotient and FRI-combination kernels read the peers' LDE columns over NVLink."""
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29520 + world),
                        os.path.join(ROOT, "tests", "shard_worker.py"), "--prove", "--check", "--log-n", "12", "--cols", "100"],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("shard proof ok") == world, r.stdout[-3000:]
