"""The `tx_with_recursion` job of bench.py alone, for a sweep over prover contexts per GPU (1 GPU):
    python tools/bench_tx_recursive.py 1 2 4 8"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import eth_tx_proof_b200 as etp
from eth_tx_proof_b200 import circuit as cc, cprog, fri_circuit as fc, parallel, prover

tables, ctls = cprog.evm_shaped_system()
ctx0 = etp.Context(0)
base_c, base_w, base_pi = cc.hash_chain_circuit(12, seed=12)
p_base = cc.CircuitProver(ctx0, base_c)
w_base = p_base.prove_words(base_w, base_pi)
layer = fc.recursive_verifier_circuit([(p_base, w_base, base_pi)])
p_layer = cc.CircuitProver(ctx0, layer[0])
w_layer = p_layer.prove_words(layer[1], layer[2])
root = fc.recursive_verifier_circuit([(p_base, w_base, base_pi), (p_layer, w_layer, layer[2])])
print(f"layer circuit 2^{layer[0].degree_bits} rows, root circuit 2^{root[0].degree_bits} rows", flush=True)
del p_base, p_layer
circuits = {"layer": layer, "root": root}
chain_bits, root_bits = ("layer",) * 3, "root"
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
