"""One process per GPU: job sharding and the two small reductions the multi-GPU runs need.

The path shards with no data-path collective: per-table proofs and a block's
transaction segments are independent, exactly like the reference's job graph
`IndexedStream::from(txs).map(&TxProof).fold(&AggProof)` (/root/reference/leader/src/prover.rs:26-30)
that Paladin spreads over workers, one worker pinned to one GPU
(/root/reference/deploy/paladin-worker@.service:10).  `torch.distributed` (NCCL on GPUs, gloo in the
CPU tests) is only plumbing: a barrier, a MAX over per-rank device times, and the gather of the
KB-sized results (caps / proofs) to the leader rank — the analogue of results travelling back
through Paladin.
"""
from __future__ import annotations

from typing import Any, List, Sequence, Tuple


def shard_jobs(n_jobs: int, rank: int, world: int) -> List[int]:
    """Round-robin assignment of independent jobs (segments / tables) to ranks."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_jobs, world))


def max_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_results(local: Sequence[Any]) -> List[Any]:
    """All ranks' (job_id, result) pairs, merged and ordered by job id (every rank gets the list)."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return sorted(local, key=lambda kv: kv[0])
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, list(local))
    merged = [kv for part in out for kv in part]
    return sorted(merged, key=lambda kv: kv[0])


# ---- column-split commit of one oversized table ---------------------------------------
def cols_per_rank(n_cols: int, world: int) -> int:
    """Columns per shard: ceil(n_cols / world) rounded up to the sponge rate (8), as etp_shard_cols_per_rank."""
    cps = -(-n_cols // world)
    return -(-cps // 8) * 8


def column_split_plan(n_cols: int, lde_rows: int, cap_height: int, rank: int, world: int) -> dict:
    """What rank `rank` owns: a column range (a multiple of 8 wide, so that no 8-column sponge chunk straddles
    two GPUs), a leaf-row range and the cap entries of those rows (whole cap subtrees)."""
    if world < 1 or world & (world - 1) or world > 8:
        raise ValueError("world must be a power of two <= 8")
    if (1 << cap_height) < world:
        raise ValueError("2^cap_height must be >= world")
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    cps = cols_per_rank(n_cols, world)
    c0 = min(cps * rank, n_cols)
    c1 = min(c0 + cps, n_cols)
    rows = lde_rows // world
    caps = (1 << cap_height) // world
    return {"cols": (c0, c1), "rows": (rows * rank, rows * (rank + 1)), "cap_entries": (caps * rank, caps * (rank + 1))}


def assemble_cap(parts: Sequence[Any]):
    """Cap of the whole table = the ranks' parts in rank order."""
    import numpy as np

    return np.concatenate([np.asarray(p, dtype=np.uint64).reshape(-1, 4) for p in parts], axis=0)


def commit_column_split(shard, local_values=None, values_dev: Tuple[int, int] = None):
    """Runs the column-split commit protocol on this rank (call on every rank of the default process group):
    local transforms -> IPC handle exchange -> barrier -> leaf hashing over NVLink peer loads + own subtrees ->
    all-gather of the cap parts.  Returns the whole cap (2^cap_height x 4).  The peers' mappings stay open
    (needed for `leaves_at`); call `finish_column_split(shard)` before freeing the shard."""
    import torch.distributed as dist

    if values_dev is not None:
        shard.transform_values_dev(*values_dev)
    else:
        shard.transform_values(local_values)
    handles = [None] * shard.world
    dist.all_gather_object(handles, shard.export_handle())
    for r, hnd in enumerate(handles):
        if r != shard.rank and column_split_plan(shard.n_cols_total, 8, shard.cap_height, r, shard.world)["cols"][0] < shard.n_cols_total:
            shard.open_peer(r, hnd)
    dist.barrier()  # every LDE matrix complete (the transforms synchronise their stream) and mapped
    part = shard.commit_rows()
    parts = [None] * shard.world
    dist.all_gather_object(parts, part.tolist())
    return assemble_cap(parts)


def recommit_column_split(shard, values_dev: Tuple[int, int]):
    """Second and later commits into a shard whose peers are already mapped (the LDE buffers persist, so the IPC
    handles stay valid): local transforms -> barrier -> leaf hashing over NVLink + own subtrees -> NCCL all-gather
    of the cap parts.  Returns the whole cap."""
    import numpy as np
    import torch
    import torch.distributed as dist

    dist.barrier()  # peers have finished reading the previous LDE
    shard.transform_values_dev(*values_dev)
    dist.barrier()
    part = torch.from_numpy(shard.commit_rows().view(np.int64)).cuda()
    whole = torch.empty((shard.world * part.shape[0], 4), dtype=torch.int64, device="cuda")
    dist.all_gather_into_tensor(whole, part)
    return whole.cpu().numpy().view(np.uint64)


def finish_column_split(shard):
    import torch.distributed as dist

    dist.barrier()  # nobody is still reading this rank's LDE
    shard.close_peers()


# ---- several prover contexts on one GPU ----------------------------------------------------------------
class ProverPool:
    """`workers` independent contexts (own stream, own scratch) on one device, each driven by a host thread.

    The reference runs several Paladin workers per machine (/root/reference/README.md:100-106); on a GPU box the
    analogue is more than one prover context per GPU: the latency-bound tail of one proof (FRI tails, tree tops,
    proof-of-work, query gathers, transcript round trips) overlaps the commits of another.  Measured on B200
    (tools/prove_concurrent.py): 2 contexts give +8 % proofs/min at 2^22 rows and +45 % at 2^16; a third one
    gains nothing.  The library calls release the GIL (ctypes), so plain threads suffice."""

    def __init__(self, device: int, workers: int = 2, contexts: Sequence[Any] = None):
        """`contexts`: ready-made contexts to drive instead of creating `workers` new ones on `device`."""
        if contexts is not None:
            self.contexts = list(contexts)
            if not self.contexts:
                raise ValueError("contexts must not be empty")
            return
        import eth_tx_proof_b200 as etp

        if workers < 1:
            raise ValueError("workers must be >= 1")
        self.contexts = [etp.Context(device) for _ in range(workers)]

    def close(self):
        for c in self.contexts:
            c.close()
        self.contexts = []

    def map(self, fn, jobs: Sequence[Any]) -> List[Any]:
        """results[i] = fn(context, jobs[i]); job i runs on context i % workers, in submission order per context."""
        import threading

        results: List[Any] = [None] * len(jobs)
        errors: List[BaseException] = []

        def run(w):
            try:
                for i in range(w, len(jobs), len(self.contexts)):
                    results[i] = fn(self.contexts[w], jobs[i])
            except BaseException as e:  # surfaced on the calling thread
                errors.append(e)

        threads = [threading.Thread(target=run, args=(w,)) for w in range(len(self.contexts))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return results

    def stark_prove_dev(self, table: int, log_n: int, traces_dev: Sequence[Tuple[int, int]]) -> List[Any]:
        """One proof per (device pointer, column stride) trace, spread over the pool's contexts."""
        return self.map(lambda c, t: c.stark_prove_dev(table, log_n, t[0], t[1]), list(traces_dev))
