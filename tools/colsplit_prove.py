"""Proof of ONE oversized table column-split over all ranks (torchrun, one process per GPU): a shape_program(cols) table
(cprog.py: counter column, groups c = a*b + d / a*b*d, boolean flags; --lookups K adds K limbs range-checked against the
counter with a logUp lookup: chunked helper columns + Z per challenge) whose trace columns are generated on
the GPUs that own them, committed with the column-split commit and proved with parallel.prove_column_split.
  --verify          the leader runs the Python verifier (tests/stark_verifier.py) on the proof
  --compare-single  the leader also proves the whole table alone (must fit one GPU) and compares word for word
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/colsplit_prove.py --log-n 26 --cols 21 --verify"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import eth_tx_proof_b200 as etp
from eth_tx_proof_b200 import cprog, parallel, synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--log-n", type=int, default=20)
ap.add_argument("--cols", type=int, default=21)
ap.add_argument("--lookups", type=int, default=0)
ap.add_argument("--verify", action="store_true")
ap.add_argument("--compare-single", action="store_true")
ap.add_argument("--reps", type=int, default=1, help="prove this many times (same committed shard); the last run is reported")
ap.add_argument("--out", default=None)
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = etp.Context(local)
log_n, cols = args.log_n, args.cols
n = 1 << log_n
prog = cprog.shape_program(cols, args.lookups)
table = ctx.register_table(prog)


def columns(c0, c1):
    return syn.shape_trace_columns_dev(log_n, cols, args.lookups, c0, c1)


c0, c1 = parallel.column_split_plan(cols, 2 * n, 4, rank, world)["cols"]
xs = columns(c0, c1)
torch.cuda.synchronize()
shard = etp.BatchShard(ctx, cols, log_n, 1, 4, rank, world)
dist.barrier()
t0 = time.perf_counter()
cap = parallel.commit_column_split(shard, values_dev=(xs.data_ptr(), n))
ctx.synchronize()
dist.barrier()
t_commit = time.perf_counter() - t0
del xs
torch.cuda.empty_cache()
for _ in range(max(args.reps, 1)):
    timings = {}
    dist.barrier()
    t0 = time.perf_counter()
    proof = parallel.prove_column_split(shard, table, cap, timings=timings)
    dist.barrier()
    t_prove = time.perf_counter() - t0
leader = world - 1
res = None
if rank == leader:
    res = {"workload": f"shape_program({cols}, {args.lookups} range-checked limbs) 2^{log_n} rows column-split over {world} GPUs", "log_n": log_n, "cols": cols, "world": world,
           "trace_commit_ms": round(t_commit * 1e3, 1), "prove_after_commit_ms": round(t_prove * 1e3, 1),
           "total_ms": round((t_commit + t_prove) * 1e3, 1), "proof_words": int(proof.size), "reps": args.reps,
           "phases_ms": {k: round(v, 1) for k, v in timings.items()},
           "hbm_in_use_gb_leader": round(torch.cuda.mem_get_info()[1] / 1e9 - torch.cuda.mem_get_info()[0] / 1e9, 1)}
    if args.verify:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import stark_verifier

        t0 = time.perf_counter()
        stark_verifier.verify(proof, program=prog)
        res["verifier"] = "accepted (tests/stark_verifier.py, all %d queries)" % parallel.STARK_NUM_QUERIES
        res["verify_s"] = round(time.perf_counter() - t0, 1)
parallel.finish_column_split(shard)
del shard
ctx.trim()
if args.compare_single and rank == leader:
    whole = columns(0, cols)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    want = ctx.stark_prove_dev(table, log_n, whole.data_ptr(), n)
    res["single_gpu_prove_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
    assert want.shape == proof.shape and (want == proof).all(), "column-split proof differs from the single-GPU proof"
    res["parity"] = "proof words == the single-GPU proof of the same trace"
if rank == leader:
    line = json.dumps(res)
    print(line, flush=True)
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "a") as f:
            f.write(line + "\n")
dist.barrier()
ctx.close()
dist.destroy_process_group()
