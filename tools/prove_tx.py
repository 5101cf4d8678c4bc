"""Synthetic transaction job (seven table proofs of the evm_arithmetization shapes) on one GPU: per-table and total times.
   python tools/prove_tx.py [scale_bits] [contexts]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import eth_tx_proof_b200 as etp
from eth_tx_proof_b200 import parallel, synthetic as syn

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 0
workers = int(sys.argv[2]) if len(sys.argv) > 2 else 2
t0 = time.perf_counter()
tables = syn.tx_job_tables(scale)
print(f"traces built in {time.perf_counter() - t0:.1f} s")
dev = [torch.from_numpy(t.view(np.int64)).cuda() for _, _, _, t in tables]
torch.cuda.synchronize()
pool = parallel.ProverPool(0, workers)
t0 = time.perf_counter()
ids = [[(c.register_table(p, p.lookups) if p is not None else etp.TABLE_MEMORY) for _, p, _, _ in tables] for c in pool.contexts]
print(f"tables registered (NVRTC) on {workers} context(s) in {time.perf_counter() - t0:.1f} s")
c0 = pool.contexts[0]
for rep in range(2):
    line = []
    tt = time.perf_counter()
    for k, (name, _, bits, _) in enumerate(tables):
        t1 = time.perf_counter()
        proof = c0.stark_prove_dev(ids[0][k], bits, dev[k].data_ptr(), 1 << bits)
        line.append(f"{name} 2^{bits}x{tables[k][3].shape[0]}: {(time.perf_counter() - t1) * 1e3:.2f} ms ({proof.size * 8 >> 10} KiB)")
        if rep == 1 and name in ("keccak", "logic"):
            print("   ", name, {kk.split(":")[-1].strip()[:22]: round(v, 2) for kk, v in c0.last_prove_timings().items()})
    print(f"one tx, one context: {(time.perf_counter() - tt) * 1e3:.2f} ms | " + " | ".join(line))
n_tx = 8
def run(c, job):
    k = job
    w = pool.contexts.index(c)
    return c.stark_prove_dev(ids[w][k], tables[k][2], dev[k].data_ptr(), 1 << tables[k][2])
flat = [k for _ in range(n_tx) for k in range(len(tables))]
pool.map(run, flat[: 2 * len(tables)])
t0 = time.perf_counter()
pool.map(run, flat)
dt = time.perf_counter() - t0
print(f"{n_tx} txs ({len(flat)} table proofs) through {workers} context(s): {dt * 1e3:.1f} ms -> {n_tx * 60 / dt:.0f} tx/min")
pool.close()
