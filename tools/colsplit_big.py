"""ONE oversized table column-split over all ranks (torchrun, one process per GPU): commit, timing, and size-independent checks
— sampled Merkle paths verify against the assembled cap (hashing with the library's own host Poseidon), rows gathered across
GPUs equal the LDE of the same column committed alone on one GPU.
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/colsplit_big.py 26 128"""
import ctypes as C
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import eth_tx_proof_b200 as etp
from eth_tx_proof_b200 import parallel

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 128
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = etp.Context(local)
n, cap_h = 1 << log_n, 4
L = etp.load_library()


def column(c):  # any rank can rebuild any column
    g = torch.Generator(device="cuda").manual_seed(1000 + c)
    return torch.randint(0, 2**62, (n,), dtype=torch.int64, device="cuda", generator=g)


c0, c1 = parallel.column_split_plan(cols, 2 * n, cap_h, rank, world)["cols"]
xs = torch.stack([column(c) for c in range(c0, c1)]) if c1 > c0 else torch.zeros((0, n), dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
shard = etp.BatchShard(ctx, cols, log_n, 1, cap_h, rank, world)
cap = parallel.commit_column_split(shard, values_dev=(xs.data_ptr(), n))
dist.barrier()
t0 = time.perf_counter()
reps = 2
for _ in range(reps):
    cap2 = parallel.recommit_column_split(shard, (xs.data_ptr(), n))
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / reps
t = torch.tensor([dt], dtype=torch.float64, device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert (cap == cap2).all()


def permute(state):
    a = (C.c_uint64 * 12)(*[int(x) for x in state])
    L.etp_host_poseidon_permute(a)
    return list(a)


def hash_no_pad(row):
    st = [0] * 12
    for off in range(0, len(row), 8):
        chunk = row[off:off + 8]
        st[:len(chunk)] = [int(x) for x in chunk]
        st = permute(st)
    return st[:4]


# sampled leaves of the rows THIS rank owns: path -> cap entry
lde_n = 2 * n
rng = np.random.default_rng(5 + rank)
idx = [shard.first_row, shard.first_row + shard.num_rows - 1] + [int(shard.first_row + x) for x in rng.integers(0, shard.num_rows, 3)]
rows = shard.leaves_at(idx)
for i, row in zip(idx, rows):
    cur = hash_no_pad(list(row))
    j = i
    for sib in shard.prove(i):
        s = [int(x) for x in sib]
        cur = permute((s + cur if j & 1 else cur + s) + [0] * 4)[:4]
        j >>= 1
    assert cur == [int(x) for x in cap[j]], f"rank {rank}: Merkle path of leaf {i} does not reach the cap"
# the first local column committed alone on this GPU: same LDE values in the gathered rows
if c1 > c0:
    single = etp.PolynomialBatch.from_values_dev(ctx, xs.data_ptr(), n, 1, log_n, 1, False, 0)
    assert (single.leaves_at(idx)[:, 0] == rows[:, c0]).all(), "column-split LDE differs from the single-GPU LDE"
    del single
parallel.finish_column_split(shard)
if rank == 0:
    nbytes = 8 * cols * n * 4 + 32 * (2 * (2 * n - 16) + 16)
    print(f"column-split commit 2^{log_n} x {cols} over {world} GPUs: {float(t.item()) * 1e3:.1f} ms per commit = {nbytes / float(t.item()) / 1e9:.1f} GB/s algorithmic; "
          f"paths of {len(idx)} sampled leaves per rank verify against the cap; gathered rows == single-GPU LDE", flush=True)
dist.destroy_process_group()
