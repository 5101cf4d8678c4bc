"""GPU parity tests of (1) the streamed host commit — column groups transformed and absorbed by the leaf
sponges while the next group crosses PCIe — and (2) the column-split commit of one table across GPUs
(SURVEY.md 8(e)): the assembled cap, Merkle paths, rows and coefficients must equal those of the unsplit
PolynomialBatch::from_values (plonky2/src/fri/oracle.rs) and of the oracle.  Single-GPU runs wire G shards
inside one process (all on device 0, peers set by pointer); with >= 2 GPUs the one-process-per-GPU protocol
(CUDA IPC handles over torch.distributed, NVLink peer loads inside the hashing kernel) runs under torchrun."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 0xFFFFFFFF00000001


@pytest.fixture(scope="module")
def ctx():
    import eth_tx_proof_b200 as etp

    c = etp.Context(0)
    yield c
    c.close()


def np_rand(seed, shape):
    from eth_tx_proof_b200 import synthetic as syn

    return syn._rand(seed, 0, int(np.prod(shape))).reshape(shape)  # any u64, non-canonical included


@pytest.mark.parametrize("cols,log_n", [(128, 15), (100, 16), (9, 19), (23, 18), (129, 15), (68, 16)])
def test_streamed_host_commit_matches_device_commit_and_oracle(ctx, cols, log_n):
    """Sizes above the 32 MiB pipelining threshold: several column groups, ragged last group / last chunk."""
    import torch

    import eth_tx_proof_b200 as etp
    import oracle

    vals = np_rand(5 + cols, (cols, 1 << log_n))
    host = etp.PolynomialBatch.from_values(ctx, vals, 1, False, 4)
    d = torch.from_numpy(vals.view(np.int64)).cuda()
    dev = etp.PolynomialBatch.from_values_dev(ctx, d.data_ptr(), 1 << log_n, cols, log_n, 1, False, 4)
    assert (host.cap == dev.cap).all()
    idx = [0, 1, (1 << (log_n + 1)) - 1, 12345]
    assert (host.leaves_at(idx) == dev.leaves_at(idx)).all()
    for i in idx:
        assert (host.prove(i) == dev.prove(i)).all()
    if log_n <= 16:
        o = oracle.Batch.from_values(vals, 1, 4)
        assert (host.cap == o.cap).all()
        assert (host.polynomials == o.coeffs).all()
    # from_coeffs through the same streamed path
    co = dev.polynomials
    hc = etp.PolynomialBatch.from_coeffs(ctx, co, 1, False, 4)
    assert (hc.cap == dev.cap).all()


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("cols,log_n,cap", [(128, 10, 4), (21, 9, 3), (7, 8, 3), (67, 11, 4)])
def test_column_split_commit_in_one_process(ctx, world, cols, log_n, cap):
    import eth_tx_proof_b200 as etp
    import oracle
    from eth_tx_proof_b200 import parallel

    if (1 << cap) < world:
        with pytest.raises(etp.EtpError):
            etp.BatchShard(ctx, cols, log_n, 1, cap, 0, world)
        return
    vals = np_rand(77 + cols, (cols, 1 << log_n))
    whole = etp.PolynomialBatch.from_values(ctx, vals, 1, False, cap)
    o = oracle.Batch.from_values(vals, 1, cap)
    assert (whole.cap == o.cap).all()
    shards = [etp.BatchShard(ctx, cols, log_n, 1, cap, r, world) for r in range(world)]
    for r, s in enumerate(shards):
        plan = parallel.column_split_plan(cols, 2 << log_n, cap, r, world)
        assert (s.first_col, s.first_col + s.num_local_cols) == plan["cols"] or s.num_local_cols == 0
        assert (s.first_row, s.first_row + s.num_rows) == plan["rows"]
        s.transform_values(vals[plan["cols"][0]:plan["cols"][1]])
    for s in shards:
        for r, t in enumerate(shards):
            if r != s.rank and t.num_local_cols:
                s.set_peer(r, t.lde_ptr)
    parts = [s.commit_rows() for s in shards]
    assert (parallel.assemble_cap(parts) == whole.cap).all()
    lde_n = 2 << log_n
    idx = [0, 1, lde_n // 2 - 1, lde_n // 2, lde_n - 1, 37 % lde_n]
    for s in shards:
        assert (s.leaves_at(idx) == whole.leaves_at(idx)).all()
        c0 = s.first_col
        assert (s.polynomials == o.coeffs[c0:c0 + s.num_local_cols]).all()
        for i in idx:
            if s.first_row <= i < s.first_row + s.num_rows:
                assert (s.prove(i) == whole.prove(i)).all()
            else:
                with pytest.raises(etp.EtpError):
                    s.prove(i)


def test_column_split_misuse(ctx):
    import eth_tx_proof_b200 as etp

    with pytest.raises(etp.EtpError):
        etp.BatchShard(ctx, 16, 8, 1, 4, 0, 3)       # world not a power of two
    with pytest.raises(etp.EtpError):
        etp.BatchShard(ctx, 16, 8, 1, 4, 2, 2)       # rank out of range
    s = etp.BatchShard(ctx, 16, 8, 1, 4, 0, 2)
    s.transform_values(np.zeros((8, 256), dtype=np.uint64))
    with pytest.raises(etp.EtpError):                 # peer 1 not mapped
        s.commit_rows()


def _n_gpus():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_column_split_commit_across_gpus(world):
    """One process per GPU, CUDA IPC + NVLink peer loads; every rank checks the assembled cap against the
    unsplit commit on its own GPU and against the oracle."""
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29500 + world),
                        os.path.join(ROOT, "tests", "shard_worker.py"), "--check"], env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("shard ok") == world, r.stdout[-3000:]


def test_context_on_another_device_from_a_fresh_thread():
    """Callers are arbitrary host threads (tokio workers upstream, parallel.ProverPool here): every entry point binds the
    calling thread to its context's device.  A thread that never touched CUDA drives a context on device 1."""
    import threading

    import eth_tx_proof_b200 as etp
    from eth_tx_proof_b200 import synthetic as syn

    if etp.load_library().etp_device_count() < 2:
        pytest.skip("needs two GPUs")
    vals = syn.random_columns(11, 10, seed=9)
    c0 = etp.Context(0)
    want = etp.PolynomialBatch.from_values(c0, vals, 1, False, 4).cap.copy()
    out = {}

    def work():
        try:
            c1 = etp.Context(1)
            out["cap"] = etp.PolynomialBatch.from_values(c1, vals, 1, False, 4).cap.copy()
            out["proof"] = c1.stark_prove(etp.TABLE_MEMORY, syn.memory_trace(8, seed=2))
            c1.close()
        except BaseException as e:  # noqa: BLE001
            out["err"] = e

    th = threading.Thread(target=work)
    th.start()
    th.join()
    assert "err" not in out, out.get("err")
    assert (out["cap"] == want).all()
    assert (out["proof"] == c0.stark_prove(etp.TABLE_MEMORY, syn.memory_trace(8, seed=2))).all()
    c0.close()


def test_host_pin_makes_no_difference_to_results(ctx):
    """etp_host_pin / etp_host_unpin (cudaHostRegister on caller-owned memory): same commitment from pageable and from
    pinned columns."""
    import eth_tx_proof_b200 as etp
    from eth_tx_proof_b200 import synthetic as syn

    vals = syn.random_columns(24, 16, seed=77)  # 12 MiB: below the streaming threshold is fine, the call path is the same
    cap0 = etp.PolynomialBatch.from_values(ctx, vals, 1, False, 4).cap.copy()
    ctx.pin(vals)
    try:
        cap1 = etp.PolynomialBatch.from_values(ctx, vals, 1, False, 4).cap.copy()
    finally:
        ctx.unpin(vals)
    assert (cap0 == cap1).all()
    with pytest.raises(etp.EtpError):
        ctx.unpin(vals)  # not registered any more
