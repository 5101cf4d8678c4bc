"""One synthetic transaction through the reference's whole proof structure on the device (eth_tx_proof_b200/stark_circuit.py):
seven table STARKs on one transcript with CTLs -> per table the wrapper circuit of that proof + a shrinking step -> the root
circuit -> aggregation levels -> the block circuit; every circuit proof checked by the Python verifier when --verify is given.

    python tools/tx_recursion.py [--small] [--verify] [--levels 3] [--reps 3]

--small uses the 2^5..2^7-row variant of the seven tables (seconds); the default sizes are the bench's (2^10..2^18 rows)."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import eth_tx_proof_b200 as etp
from eth_tx_proof_b200 import circuit as cc, cprog, prover, stark_circuit as sc

ap = argparse.ArgumentParser()
ap.add_argument("--small", action="store_true")
ap.add_argument("--verify", action="store_true")
ap.add_argument("--levels", type=int, default=3)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()

ctx = etp.Context(0)
small_bits = {"arithmetic": 6, "byte_packing": 5, "cpu": 6, "keccak": 5, "keccak_sponge": 5, "logic": 5, "memory": 7}
tables, ctls = cprog.evm_shaped_system(degree_bits=small_bits) if args.small else cprog.evm_shaped_system()
ids = [ctx.register_table(p) for _, p, _ in tables]
dev = [torch.from_numpy(t.view(np.int64)).cuda() for _, _, t in tables]
traces_dev = [(d.data_ptr(), t.shape[1], t.shape[0], int(t.shape[1]).bit_length() - 1) for d, (_, _, t) in zip(dev, tables)]
torch.cuda.synchronize()
pv = [0xB200, 1, 2, 3] * 2
t0 = time.perf_counter()
first = prover.prove_with_traces(ctx, ids, traces_dev, pv)
provers = []


def circuit_prove(circuit, wires, pis):
    cp = cc.CircuitProver(ctx, circuit)
    provers.append(cp)
    words = cp.prove_words(wires, pis)
    if args.verify:
        import plonk_verifier

        plonk_verifier.verify(cp.prove(wires, pis), circuit, cp.constants_sigmas_cap, cp.digest, max_queries=2)
    return cp, words


log = lambda s: print(f"  [{time.perf_counter() - t0:6.1f} s] {s}")
plan = sc.transaction_recursion_plan(tables, ctls, first, circuit_prove, public_values=pv, log=log)
plan += sc.block_recursion_plan(plan[-1], len(tables), circuit_prove, levels=args.levels, log=log)
print(f"built and proved once in {time.perf_counter() - t0:.1f} s; root public inputs: {len(plan[-2 - args.levels]['public_inputs'])}, "
      f"block public inputs: {plan[-1]['public_inputs']}")
for _ in range(args.reps):
    t1 = time.perf_counter()
    prover.prove_with_traces(ctx, ids, traces_dev, pv)
    t_stark = (time.perf_counter() - t1) * 1e3
    ms = {}
    for cp, s in zip(provers, plan):
        t2 = time.perf_counter()
        cp.prove_words(s["wires"], s["public_inputs"])
        ms[s["kind"]] = ms.get(s["kind"], 0.0) + (time.perf_counter() - t2) * 1e3
    print(f"table STARKs {t_stark:.1f} ms | " + " | ".join(f"{k} {v:.1f} ms" for k, v in ms.items()))
