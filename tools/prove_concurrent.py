"""Do several prover contexts on ONE GPU raise proofs/min?  (the latency-bound tail of one proof — FRI tails, tree
tops, PoW, query gathers — can overlap another proof's commits)   python tools/prove_concurrent.py [log_n] [jobs]"""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import eth_tx_proof_b200 as etp
from eth_tx_proof_b200 import synthetic as syn

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
jobs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
t = torch.from_numpy(syn.memory_trace(log_n).view(np.int64)).cuda()
torch.cuda.synchronize()
for workers in (1, 2, 3):
    ctxs = [etp.Context(0) for _ in range(workers)]
    for c in ctxs:
        c.stark_prove_dev(etp.TABLE_MEMORY, log_n, t.data_ptr(), 1 << log_n)
    def work(c, k):
        for _ in range(k):
            c.stark_prove_dev(etp.TABLE_MEMORY, log_n, t.data_ptr(), 1 << log_n)
    per = [jobs // workers + (1 if i < jobs % workers else 0) for i in range(workers)]
    th = [threading.Thread(target=work, args=(c, k)) for c, k in zip(ctxs, per)]
    t0 = time.perf_counter()
    for x in th: x.start()
    for x in th: x.join()
    dt = time.perf_counter() - t0
    print(f"2^{log_n}: {workers} context(s): {jobs} proofs in {dt*1e3:.1f} ms -> {jobs*60/dt:.0f} proofs/min")
    for c in ctxs: c.close()
