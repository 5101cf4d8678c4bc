"""The `tx_with_recursion` job of bench.py alone, for a sweep over prover contexts per GPU (1 GPU):
    python tools/bench_tx_recursive.py 1 2 4 8"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from eth_tx_proof_b200 import circuit as cc, cprog, parallel, prover

tables, ctls = cprog.evm_shaped_system()
chain_bits, root_bits = (13, 13, 12), 13
circuits = {db: cc.hash_chain_circuit(db, seed=db) for db in sorted(set(chain_bits) | {root_bits})}
dev = [torch.from_numpy(t.view(np.int64)).cuda() for _, _, t in tables]
traces_dev = [(d.data_ptr(), t.shape[1], t.shape[0], int(t.shape[1]).bit_length() - 1) for d, (_, _, t) in zip(dev, tables)]
torch.cuda.synchronize()
for n_ctx in [int(a) for a in sys.argv[1:]] or [1, 2, 4]:
    pool = parallel.ProverPool(0, n_ctx)
    ids = [[c.register_table(p) for _, p, _ in tables] for c in pool.contexts]
    cprovers = [{db: cc.CircuitProver(c, circuits[db][0]) for db in circuits} for c in pool.contexts]

    def job(c, _):
        k = pool.contexts.index(c)
        out = [prover.prove_with_traces(c, ids[k], traces_dev)]
        for _table in range(len(tables)):
            for db in chain_bits:
                out.append(cprovers[k][db].prove_words(circuits[db][1], circuits[db][2]))
        out.append(cprovers[k][root_bits].prove_words(circuits[root_bits][1], circuits[root_bits][2]))
        return out

    pool.map(job, list(range(n_ctx)))
    n_jobs = 16
    t0 = time.perf_counter()
    pool.map(job, list(range(n_jobs)))
    dt = time.perf_counter() - t0
    print(f"contexts {n_ctx}: {dt / n_jobs * 1e3:.1f} ms per job, {n_jobs * 60 / dt:.0f} jobs/min", flush=True)
    pool.close()
    del cprovers
