"""One proof of a shape-only table after a warm-up (target of ncu launch lists): python tools/prove_shape.py name [bits]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import eth_tx_proof_b200 as etp
from eth_tx_proof_b200 import cprog
name = sys.argv[1] if len(sys.argv) > 1 else "keccak"
bits = int(sys.argv[2]) if len(sys.argv) > 2 else cprog.TX_TABLE_DEGREE_BITS[name]
cols, lk = cprog.EVM_TABLE_SHAPES[name]
ctx = etp.Context(0)
prog = cprog.shape_program(cols, lk)
tid = ctx.register_table(prog, prog.lookups)
t = torch.from_numpy(cprog.shape_trace(bits, cols, lk).view(np.int64)).cuda()
torch.cuda.synchronize()
import time
for _ in range(3):
    t0 = time.perf_counter()
    ctx.stark_prove_dev(tid, bits, t.data_ptr(), 1 << bits)
    dt = time.perf_counter() - t0
print(f"{name} 2^{bits} x {cols}: {dt * 1e3:.2f} ms", {k.split(':')[-1].strip()[:24]: round(v, 3) for k, v in ctx.last_prove_timings().items()})
