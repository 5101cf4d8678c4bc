"""One synthetic CTL transaction (cprog.evm_shaped_system through prover.prove_with_traces) with per-table phase timings:
python tools/prove_tx_ctl.py [scale_bits]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import eth_tx_proof_b200 as etp
from eth_tx_proof_b200 import cprog
from eth_tx_proof_b200.api import Challenger, PolynomialBatch

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 0
tables, ctls = cprog.evm_shaped_system(scale_bits=scale)
ctx = etp.Context(0)
ids = [ctx.register_table(p) for _, p, _ in tables]
dev = [torch.from_numpy(t.view(np.int64)).cuda() for _, _, t in tables]
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    coms = [PolynomialBatch.from_values_dev(ctx, d.data_ptr(), t.shape[1], t.shape[0], int(t.shape[1]).bit_length() - 1, 1, False, 4)
            for d, (_, _, t) in zip(dev, tables)]
    t_commit = (time.perf_counter() - t0) * 1e3
    ch = Challenger()
    for c in coms:
        ch.observe_cap(c.cap)
    ctl = ch.get_n_challenges(4)
    rows = []
    for tid, d, (name, p, t), com in zip(ids, dev, tables, coms):
        ch.compact()
        t1 = time.perf_counter()
        ctx.prove_with_commitment(tid, com, d.data_ptr(), t.shape[1], ch, ctl)
        wall = (time.perf_counter() - t1) * 1e3
        ph = ctx.last_prove_timings()
        rows.append((name, t.shape, wall, sum(ph.values()), ph))
    total = (time.perf_counter() - t0) * 1e3
print(f"trace commits {t_commit:.2f} ms, transaction {total:.2f} ms")
for name, shape, wall, devsum, ph in rows:
    print(f"{name:14s} {str(shape):16s} wall {wall:7.2f} ms  device phases {devsum:7.2f} ms  | " + "  ".join(f"{k.split(':')[-1].strip()[:22]}={v:.2f}" for k, v in ph.items()))
