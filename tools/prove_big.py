"""A memory-table proof at a large size, checked by the independent verifier (a few query rounds): python tools/prove_big.py [log_n]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import eth_tx_proof_b200 as etp
from eth_tx_proof_b200 import synthetic as syn
import stark_verifier as V

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
t0 = time.perf_counter(); t = syn.memory_trace(log_n); print(f"trace 2^{log_n} x {t.shape[0]} built in {time.perf_counter() - t0:.1f} s")
ctx = etp.Context(0)
ctx.pin(t)
for _ in range(2):
    t0 = time.perf_counter(); proof = ctx.stark_prove(etp.TABLE_MEMORY, t); dt = time.perf_counter() - t0
print(f"prove_host 2^{log_n}: {dt * 1e3:.1f} ms, proof {proof.size * 8 >> 10} KiB", {k.split(':')[-1].strip()[:20]: round(v, 1) for k, v in ctx.last_prove_timings().items()})
ctx.unpin(t)
t0 = time.perf_counter(); V.verify(proof, max_queries=3); print(f"verifier accepted (3 query rounds) in {time.perf_counter() - t0:.1f} s")
