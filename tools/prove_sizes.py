"""Scratch timing: single-table STARK prove time vs table size (memory-shaped table), trace resident in HBM."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import eth_tx_proof_b200 as etp
from eth_tx_proof_b200 import synthetic as syn

ctx = etp.Context(0)
for log_n in [int(a) for a in sys.argv[1:]] or [8, 10, 12, 14, 16, 18, 20, 22]:
    t = torch.from_numpy(syn.memory_trace(log_n).view(np.int64)).cuda()
    torch.cuda.synchronize()
    ctx.stark_prove_dev(etp.TABLE_MEMORY, log_n, t.data_ptr(), 1 << log_n)
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        ctx.stark_prove_dev(etp.TABLE_MEMORY, log_n, t.data_ptr(), 1 << log_n)
    dt = (time.perf_counter() - t0) / reps
    ph = ctx.last_prove_timings()
    print(f"2^{log_n}: {dt * 1e3:8.3f} ms  " + "  ".join(f"{k.split(':')[-1].strip()[:18]}={v:.2f}" for k, v in ph.items()))
