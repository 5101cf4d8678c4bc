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

from typing import Any, List, Sequence


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
