"""World-size-2 test (gloo, CPU) of the multi-GPU host logic: independent jobs are sharded over ranks with
no data-path collective, per-rank times reduce with MAX, results are gathered to every rank in job order.
The per-job work here is the oracle's commit (CPU) — the same sharding code drives the CUDA path in
bench.py under torchrun (one process per GPU, NCCL)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import oracle
    from eth_tx_proof_b200 import parallel, synthetic as syn

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    jobs = parallel.shard_jobs(5, rank, world)
    local = []
    for j in jobs:
        vals = syn.random_columns(9, 6, seed=100 + j)
        local.append((j, oracle.Batch.from_values(vals, 1, 2).cap.tolist()))
    merged = parallel.gather_results(local)
    t = parallel.max_over_ranks(10.0 + rank)
    # column-split commit, host side: each rank contributes the cap entries of the rows it owns (here cut out of
    # the oracle's cap of the whole table) and every rank assembles the same whole cap
    vals = syn.random_columns(20, 6, seed=9)
    whole = oracle.Batch.from_values(vals, 1, 3).cap
    plan = parallel.column_split_plan(20, 2 << 6, 3, rank, world)
    lo, hi = plan["cap_entries"]
    parts = [None] * world
    dist.all_gather_object(parts, whole[lo:hi].tolist())
    assert (parallel.assemble_cap(parts) == whole).all()
    dist.barrier()
    q.put((rank, jobs, merged, t))
    dist.destroy_process_group()


def test_two_ranks_shard_reduce_gather():
    import torch.multiprocessing as mp

    sys.path.insert(0, ROOT)
    import oracle
    from eth_tx_proof_b200 import parallel, synthetic as syn

    assert parallel.shard_jobs(5, 0, 2) == [0, 2, 4] and parallel.shard_jobs(5, 1, 2) == [1, 3]
    assert sorted(parallel.shard_jobs(8, 0, 4) + parallel.shard_jobs(8, 1, 4) + parallel.shard_jobs(8, 2, 4)
                  + parallel.shard_jobs(8, 3, 4)) == list(range(8))
    with pytest.raises(ValueError):
        parallel.shard_jobs(3, 2, 2)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [(j, oracle.Batch.from_values(syn.random_columns(9, 6, seed=100 + j), 1, 2).cap.tolist()) for j in range(5)]
    for rank, jobs, merged, t in res:
        assert jobs == parallel.shard_jobs(5, rank, 2)
        assert merged == want          # every rank sees all results, in job order
        assert t == 11.0               # MAX over ranks


def test_single_process_fallbacks():
    sys.path.insert(0, ROOT)
    from eth_tx_proof_b200 import parallel

    assert parallel.max_over_ranks(3.5) == 3.5
    assert parallel.gather_results([(1, "b"), (0, "a")]) == [(0, "a"), (1, "b")]


def test_column_split_plan_covers_columns_rows_and_cap():
    sys.path.insert(0, ROOT)
    from eth_tx_proof_b200 import parallel

    for n_cols, world, cap in [(128, 8, 4), (21, 2, 4), (7, 4, 2), (100, 8, 3), (1, 2, 1)]:
        lde = 1 << 12
        cols, rows, caps = [], [], []
        for r in range(world):
            p = parallel.column_split_plan(n_cols, lde, cap, r, world)
            assert p["cols"][0] % 8 == 0 or p["cols"][0] == n_cols   # sponge chunks never straddle two GPUs
            cols += list(range(*p["cols"]))
            rows.append(p["rows"])
            caps += list(range(*p["cap_entries"]))
        assert cols == list(range(n_cols))
        assert rows[0][0] == 0 and rows[-1][1] == lde and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
        assert caps == list(range(1 << cap))
    assert parallel.cols_per_rank(128, 8) == 16 and parallel.cols_per_rank(100, 8) == 16 and parallel.cols_per_rank(21, 2) == 16
    with pytest.raises(ValueError):
        parallel.column_split_plan(16, 64, 1, 0, 4)   # fewer cap subtrees than ranks
    with pytest.raises(ValueError):
        parallel.column_split_plan(16, 64, 4, 0, 3)


def test_prover_pool_scheduling_and_error_propagation():
    """ProverPool.map without a GPU (stand-in contexts): job i runs on context i % workers, in submission order per
    context, results come back in job order, and a failure on a worker thread is raised on the caller."""
    import threading

    from eth_tx_proof_b200 import parallel

    class Fake:
        def __init__(self, name):
            self.name, self.seen, self.threads = name, [], set()

        def close(self):
            self.closed = True

    ctxs = [Fake("a"), Fake("b"), Fake("c")]
    pool = parallel.ProverPool(0, contexts=ctxs)

    def fn(c, job):
        c.seen.append(job)
        c.threads.add(threading.get_ident())
        return (c.name, job * job)

    out = pool.map(fn, list(range(10)))
    assert out == [("abc"[i % 3], i * i) for i in range(10)]
    assert ctxs[0].seen == [0, 3, 6, 9] and ctxs[1].seen == [1, 4, 7] and ctxs[2].seen == [2, 5, 8]
    assert all(len(c.threads) == 1 for c in ctxs) and threading.get_ident() not in ctxs[0].threads
    assert pool.map(fn, []) == []

    def boom(c, job):
        if job == 4:
            raise RuntimeError("job 4 failed")
        return job

    with pytest.raises(RuntimeError, match="job 4 failed"):
        pool.map(boom, list(range(6)))
    pool.close()
    assert all(getattr(c, "closed", False) for c in ctxs)
    with pytest.raises(ValueError):
        parallel.ProverPool(0, contexts=[])


def test_thread_comm_all_gather_and_collective_errors():
    """Host logic of the column-split proof (parallel.prove_column_split): ThreadComm gathers in rank order, repeatedly;
    a failure on one rank is raised on every rank — nobody stays blocked in the next collective."""
    import threading

    sys.path.insert(0, ROOT)
    from eth_tx_proof_b200 import EtpError, parallel

    world = 4
    comm = parallel.ThreadComm(world)
    got, errs = [None] * world, [None] * world

    def work(r):
        c = comm.rank(r)
        a = parallel._all_gather({"r": r}, c)
        b = parallel._collective(lambda: r * r, c)
        try:
            def step():
                if r == 2:
                    raise EtpError(-4, "boom")
                return r
            parallel._collective(step, c)
        except EtpError as e:
            errs[r] = e
        got[r] = (a, b, parallel._collective(lambda: "after", c))  # the communicator is still usable

    ths = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in ths:
        t.start()
    for t in ths:
        t.join(timeout=30)
    assert all(not t.is_alive() for t in ths)
    for r in range(world):
        a, b, c = got[r]
        assert a == [{"r": k} for k in range(world)] and b == [0, 1, 4, 9] and c == ["after"] * world
        assert errs[r] is not None and errs[r].code == -4 and "rank 2" in str(errs[r]) and "boom" in str(errs[r])


def _collective_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from eth_tx_proof_b200 import EtpError, parallel

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = parallel._collective(lambda: rank + 10, None)
    try:
        def step():
            if rank == 1:
                raise ValueError("bad shard")
            return rank
        parallel._collective(step, None)
        err = None
    except EtpError as e:
        err = (e.code, str(e))
    q.put((rank, ok, err))
    dist.barrier()
    dist.destroy_process_group()


def test_collective_errors_across_processes():
    import torch.multiprocessing as mp

    world, port = 2, _free_port()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    ps = [ctxm.Process(target=_collective_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, err in res:
        assert ok == [10, 11]
        assert err is not None and err[0] == -3 and "rank 1" in err[1] and "bad shard" in err[1]
