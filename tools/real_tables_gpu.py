"""The seven-table transaction with real semantics (eth_tx_proof_b200/evm_tables.py real_transaction_system) on the DEVICE:
prove_with_traces through the C ABI, every table proof compared word for word with the oracle's, the CTL sums verified.

    ETP_CUBIN_CACHE=/tmp/etp_cubins python tools/real_tables_gpu.py [--skip-keccak]

NOT run on a GPU in round 2 (the round's GPU budget was spent before this script existed); the pieces it uses were: the arithmetic
table and the segmented kernels in tests/test_gpu_evm_tables.py, prove_with_traces in tests/test_gpu_ctl.py.  The Keccak-f table's
56 k-op program takes ~9 minutes of NVRTC on first registration (segmented code generation, csrc/cprog.h; cached on disk with
ETP_CUBIN_CACHE); --skip-keccak only registers (compiles) the six other tables and exits."""
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
import oracle
from eth_tx_proof_b200 import evm_tables as et, prover
from test_ctl_oracle import verify_all

ap = argparse.ArgumentParser()
ap.add_argument("--skip-keccak", action="store_true")
args = ap.parse_args()
t0 = time.perf_counter()
log = lambda s: print(f"[{time.perf_counter() - t0:7.1f} s] {s}", flush=True)
tables, ctls = et.real_transaction_system()
ctx = etp.Context(0)
if args.skip_keccak:
    log("registering six tables (the Keccak-f table is skipped: nothing can be proven as a system without it)")
    for name, prog, _ in tables:
        if name != "keccak":
            ctx.register_table(prog)
            log(f"  {name}: {len(prog.ops)} ops compiled")
    sys.exit(0)
tids = []
for name, prog, _ in tables:
    tids.append(ctx.register_table(prog))
    log(f"registered {name}: {len(prog.ops)} ops")
dev = [torch.from_numpy(np.ascontiguousarray(t).view(np.int64)).cuda() for _, _, t in tables]
torch.cuda.synchronize()
traces_dev = [(d.data_ptr(), t.shape[1], t.shape[0], int(t.shape[1]).bit_length() - 1) for d, (_, _, t) in zip(dev, tables)]
got = prover.prove_with_traces(ctx, tids, traces_dev)
log("seven table proofs on the device")
oids = [oracle.register_table_ex(p, p.aux_spec) for _, p, _ in tables]
batches = [oracle.Batch.from_values(t, 1, 4) for _, _, t in tables]
och = oracle.HostChallenger()
for b in batches:
    och.observe(b.cap)
octl = och.get_n(4)
for k, (oid, (name, _, t), b) in enumerate(zip(oids, tables, batches)):
    och.compact()
    want = oracle.prove_with_commitment(oid, t, b, och, octl)
    assert (np.delete(got.stark_proofs[k], 1) == np.delete(want, 1)).all(), f"{name}: device proof differs from the oracle"
    log(f"  {name}: device proof == oracle")
verify_all(tables, ctls, got.stark_proofs, got.trace_caps, max_queries=2)
log("verifier accepts all seven proofs and the seven cross-table lookups")
