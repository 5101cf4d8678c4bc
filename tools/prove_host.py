"""etp_stark_prove_host end to end (trace in host memory -> proof bytes on the host), pageable vs pinned trace.
   python tools/prove_host.py [log_n]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import eth_tx_proof_b200 as etp
from eth_tx_proof_b200 import synthetic as syn

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
ctx = etp.Context(0)
t = syn.memory_trace(log_n)
d = torch.from_numpy(t.view(np.int64)).cuda()
torch.cuda.synchronize()
want = ctx.stark_prove_dev(etp.TABLE_MEMORY, log_n, d.data_ptr(), 1 << log_n)
for label in ("pageable", "pinned"):
    if label == "pinned":
        ctx.pin(t)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        proof = ctx.stark_prove(etp.TABLE_MEMORY, t)
        ts.append((time.perf_counter() - t0) * 1e3)
    assert (proof == want).all()
    print(f"prove_host 2^{log_n} x 21 ({t.nbytes >> 20} MiB trace, {label}): ms {[round(x, 2) for x in ts]}")
ctx.unpin(t)
ts = []
for _ in range(5):
    t0 = time.perf_counter()
    ctx.stark_prove_dev(etp.TABLE_MEMORY, log_n, d.data_ptr(), 1 << log_n)
    ts.append((time.perf_counter() - t0) * 1e3)
print(f"prove_dev (trace resident): ms {[round(x, 2) for x in ts]}")
