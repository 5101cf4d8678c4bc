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
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import eth_tx_proof_b200 as etp
    from eth_tx_proof_b200 import parallel, synthetic as syn

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = etp.Context(local)
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


if __name__ == "__main__":
    main()
