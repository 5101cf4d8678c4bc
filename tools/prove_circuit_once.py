"""One circuit proof (eth_tx_proof_b200/circuit.py) after a warm-up, for profilers:
    python tools/prove_circuit_once.py [degree_bits] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import eth_tx_proof_b200 as etp
from eth_tx_proof_b200 import circuit as cc

db = int(sys.argv[1]) if len(sys.argv) > 1 else 12
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = etp.Context(0)
circ, wires, pis = cc.hash_chain_circuit(db, seed=db)
prover = cc.CircuitProver(ctx, circ)
prover.prove(wires, pis)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("circuit_proof")
for _ in range(reps):
    out = prover.prove(wires, pis)
torch.cuda.nvtx.range_pop()
print({k: round(v, 3) for k, v in out["ms"].items()})
