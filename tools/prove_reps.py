import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import eth_tx_proof_b200 as etp
from eth_tx_proof_b200 import synthetic as syn
ctx = etp.Context(0)
log_n = 22
t = torch.from_numpy(syn.memory_trace(log_n).view(np.int64)).cuda()
torch.cuda.synchronize()
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
    t0 = time.perf_counter()
    ctx.stark_prove_dev(etp.TABLE_MEMORY, log_n, t.data_ptr(), 1 << log_n)
    dt = time.perf_counter() - t0
    ph = ctx.last_prove_timings()
    print(rep, round(dt * 1e3, 2), {k[:14]: round(v, 2) for k, v in ph.items() if v > 5})
