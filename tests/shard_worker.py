"""torchrun worker of the multi-GPU column-split commit (tests/test_gpu_shard.py, bench.py --column-split)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--log-n", type=int, default=12)
    ap.add_argument("--cols", type=int, default=100)
    ap.add_argument("--prove", action="store_true", help="prove a column-split table and compare with the single-GPU proof")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import eth_tx_proof_b200 as etp
    from eth_tx_proof_b200 import parallel, synthetic as syn

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = etp.Context(local)
    if args.prove:
        prove(ctx, rank, world, args)
        ctx.close()
        dist.destroy_process_group()
        return
    cols, log_n, cap = args.cols, args.log_n, 4
    vals = syn._rand(4242, 0, cols << log_n).reshape(cols, 1 << log_n)
    shard = etp.BatchShard(ctx, cols, log_n, 1, cap, rank, world)
    c0, c1 = parallel.column_split_plan(cols, 2 << log_n, cap, rank, world)["cols"]
    whole_cap = parallel.commit_column_split(shard, vals[c0:c1])
    if args.check:
        import oracle

        ref = etp.PolynomialBatch.from_values(ctx, vals, 1, False, cap)
        assert (whole_cap == ref.cap).all(), "assembled cap differs from the unsplit commit"
        assert (whole_cap == oracle.Batch.from_values(vals, 1, cap).cap).all(), "cap differs from the oracle"
        idx = [0, shard.first_row, shard.first_row + shard.num_rows - 1, (2 << log_n) - 1]
        assert (shard.leaves_at(idx) == ref.leaves_at(idx)).all(), "rows gathered over NVLink differ"
        for i in idx[1:3]:
            assert (shard.prove(i) == ref.prove(i)).all()
        print(f"shard ok rank {rank}/{world}", flush=True)
    parallel.finish_column_split(shard)
    del shard
    ctx.close()
    dist.destroy_process_group()


def prove(ctx, rank, world, args):
    """Column-split proofs == the proofs one GPU makes of the same traces (and the oracle's at small sizes)."""
    import eth_tx_proof_b200 as etp
    from eth_tx_proof_b200 import cprog, parallel, synthetic as syn

    log_n = args.log_n
    fib, fib_pi = syn.fibonacci_trace(log_n, seed=3)
    cases = [("fibonacci", etp.TABLE_FIBONACCI, fib, list(fib_pi)),
             ("shape21", ctx.register_table(cprog.shape_program(21, 0)), cprog.shape_trace(log_n, 21, 0, seed=5), []),
             ("logic68", ctx.register_table(cprog.logic_program(1)), cprog.logic_trace(log_n, 1, seed=6), []),
             ("shape%d" % args.cols, ctx.register_table(cprog.shape_program(args.cols, 0)), cprog.shape_trace(log_n, args.cols, 0, seed=7), []),
             ("memory", etp.TABLE_MEMORY, syn.memory_trace(log_n, seed=8), []),
             ("shape40+8 lookups", ctx.register_table(cprog.shape_program(40, 8)), cprog.shape_trace(log_n, 40, 8, seed=9), [])]
    for name, table, trace, pi in cases:
        cols = trace.shape[0]
        shard = etp.BatchShard(ctx, cols, log_n, 1, 4, rank, world)
        c0, c1 = parallel.column_split_plan(cols, 2 << log_n, 4, rank, world)["cols"]
        cap = parallel.commit_column_split(shard, np.ascontiguousarray(trace[c0:c1]))
        proof = parallel.prove_column_split(shard, table, cap, pi)
        if rank == world - 1:
            want = ctx.stark_prove(table, trace, pi)
            assert proof.shape == want.shape and (proof == want).all(), f"{name}: column-split proof differs from the single-GPU proof"
            if args.check and log_n <= 12 and table == etp.TABLE_FIBONACCI:
                import oracle

                assert (proof == oracle.stark_prove(oracle.TABLE_FIBONACCI, trace, pi)).all(), "proof differs from the oracle"
        else:
            assert proof is None
        parallel.finish_column_split(shard)
        del shard
    print(f"shard proof ok rank {rank}/{world}", flush=True)


if __name__ == "__main__":
    main()
