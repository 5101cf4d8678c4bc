"""Column-split commit of ONE table over the GPUs of a box from a SINGLE process (one context per device, peers wired by
pointer after enabling peer access) — the shape ncu can profile (it must not wrap a multi-rank launcher): the leaf-hash kernel
of every device reads the other devices' LDE columns over NVLink.  python tools/colsplit_inprocess.py [log_n] [cols] [gpus]
Checks the assembled cap against the unsplit commit on device 0 when the table fits."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import eth_tx_proof_b200 as etp
from eth_tx_proof_b200 import parallel

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 128
world = int(sys.argv[3]) if len(sys.argv) > 3 else torch.cuda.device_count()
n, cap_h = 1 << log_n, 4
# peer access both ways for every pair (torch enables it on the first cross-device copy)
for a in range(world):
    for b in range(world):
        if a != b:
            assert torch.cuda.can_device_access_peer(a, b), f"no peer access {a}->{b}"
            torch.zeros(8, device=f"cuda:{a}").to(f"cuda:{b}")
torch.cuda.synchronize()
ctxs = [etp.Context(d) for d in range(world)]
g = torch.Generator(device="cuda:0").manual_seed(7)
whole = torch.randint(0, 2**62, (cols, n), dtype=torch.int64, device="cuda:0", generator=g)
shards, parts = [], []
for r in range(world):
    c0, c1 = parallel.column_split_plan(cols, 2 * n, cap_h, r, world)["cols"]
    parts.append(whole[c0:c1].to(f"cuda:{r}").contiguous())
for d in range(world):
    torch.cuda.synchronize(d)
for r in range(world):
    s = etp.BatchShard(ctxs[r], cols, log_n, 1, cap_h, r, world)
    s.transform_values_dev(parts[r].data_ptr(), n)
    shards.append(s)
for s in shards:
    for r, t in enumerate(shards):
        if r != s.rank and t.num_local_cols:
            s.set_peer(r, t.lde_ptr)
for rep in range(2):
    t0 = time.perf_counter()
    caps = [s.commit_rows() for s in shards]  # one after the other: each kernel has the NVLink to itself
    dt = (time.perf_counter() - t0) * 1e3
cap = parallel.assemble_cap(caps)
print(f"column-split commit rows 2^{log_n} x {cols} over {world} GPUs (sequential per device): {dt:.2f} ms for the hashing + subtrees")
if 8 * cols * n * 4 < 60e9:
    ref = etp.PolynomialBatch.from_values_dev(ctxs[0], whole.data_ptr(), n, cols, log_n, 1, False, cap_h)
    assert (ref.cap == cap).all(), "assembled cap differs from the unsplit commit"
    idx = [0, n, 2 * n - 1]
    assert (shards[0].leaves_at(idx) == ref.leaves_at(idx)).all()
    print("assembled cap and probe rows == unsplit commit on device 0")
