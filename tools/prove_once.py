"""One memory-table proof at 2^log_n after a warm-up (target of ncu launch lists): python tools/prove_once.py [log_n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import eth_tx_proof_b200 as etp
from eth_tx_proof_b200 import synthetic as syn
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
ctx = etp.Context(0)
t = torch.from_numpy(syn.memory_trace(log_n).view(np.int64)).cuda()
torch.cuda.synchronize()
for _ in range(2):
    ctx.stark_prove_dev(etp.TABLE_MEMORY, log_n, t.data_ptr(), 1 << log_n)
print({k: round(v, 3) for k, v in ctx.last_prove_timings().items()})
