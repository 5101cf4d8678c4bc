"""Scratch timing of the commit path (not the judged bench): python tools/quick_bench.py [log_n] [cols]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import eth_tx_proof_b200 as etp

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 128
n = 1 << log_n
ctx = etp.Context(0)
st = torch.cuda.ExternalStream(ctx.stream)
x = torch.randint(0, 2**62, (cols, n), dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
b = etp.PolynomialBatch.from_values_dev(ctx, x.data_ptr(), n, cols, log_n, 1, False, 4)
ctx.synchronize()
ts = []
for it in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record()
        b.recommit_values_dev(x.data_ptr(), n)
        e1.record()
    ctx.synchronize(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
bytes_ = 8 * cols * n * 4 + 32 * (2 * (2 * n - 16) + 16)
print("phases ms:", {k: round(v, 3) for k, v in b.last_commit_timings().items()})
print(f"commit 2^{log_n} x {cols}: ms {ts}  best {min(ts):.3f} ms  -> {bytes_ / min(ts) / 1e6:.1f} GB/s algorithmic")
if "--e2e" in sys.argv:
    import time
    import numpy as np
    del b
    host = torch.empty((cols, n), dtype=torch.int64).pin_memory()
    host.copy_(x)
    harr = host.numpy().view(np.uint64)
    b2 = etp.PolynomialBatch.from_values(ctx, harr, 1, False, 4); del b2
    ts = []
    for it in range(4):
        t0 = time.perf_counter()
        b2 = etp.PolynomialBatch.from_values(ctx, harr, 1, False, 4)
        cap = b2.cap
        ts.append((time.perf_counter() - t0) * 1e3)
        del b2
    print(f"e2e host->commit->cap 2^{log_n} x {cols}: ms {[round(t, 2) for t in ts]}  H2D alone would be {8 * cols * n / 55e6:.1f} ms at 55 GB/s")
