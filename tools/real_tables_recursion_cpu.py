"""The whole proof structure over tables with REAL semantics, on the CPU (oracle as the prover; no GPU needed):

    cpu_ops (looking)  --CTL-->  arithmetic (ADD / SUB / LT / GT / MUL, range checks)
    keccak256 (messages, pad10*1)  --2 CTLs-->  keccak (Keccak-f[1600], one round per row)

four table STARKs on one transcript -> per table the wrapper circuit (STARK verifier, the table's own constraint program
evaluated in-circuit at zeta) + a shrinking step -> the root circuit (CTL challenges, challenger chain, cross-table lookup sums)
-> aggregation -> block; every circuit proof is checked by the independent Python verifier.

    python tools/real_tables_recursion_cpu.py [--queries 1]

--queries: FRI queries verified in-circuit per proof (all 84 / 28 make the pure-Python builders and the CPU prover slow)."""
import argparse
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import oracle
import plonk_verifier
from eth_tx_proof_b200 import cprog, evm_tables as et, stark_circuit as sc
from test_circuit_cpu import _words_from_oracle_proof

ap = argparse.ArgumentParser()
ap.add_argument("--queries", type=int, default=1)
args = ap.parse_args()
t0 = time.perf_counter()
log = lambda s: print(f"[{time.perf_counter() - t0:7.1f} s] {s}", flush=True)

# ---- the four tables
n_limbs, limb_bits = 4, 5
L = et.arithmetic_layout(n_limbs, limb_bits)
at, ops = et.arithmetic_trace(6, n_limbs, limb_bits)
width = 2 + 3 * n_limbs + 1
lt = np.zeros((width, at.shape[1]), dtype=np.uint64)
lt[0] = sum(at[L["FLAG"] + k] for k in range(5)).astype(np.uint64)
lt[1] = sum(at[L["FLAG"] + k] * np.uint64(k + 1) for k in range(5))
for i in range(n_limbs):
    lt[2 + i], lt[2 + n_limbs + i], lt[2 + 2 * n_limbs + i] = at[L["A"] + i], at[L["B"] + i], at[L["C"] + i]
lt[2 + 3 * n_limbs] = at[L["CY"]]
lt = lt[:, ::-1].copy()
b = cprog.ProgramBuilder(width, 0, 3)
b.constraint(b.lv(0) * (b.lv(0) - 1))
for k in range(cprog.NUM_CHALLENGES):
    b.add_ctl_z(k, [(list(range(1, width)), cprog.Filter(constants=[cprog.Column.single(0)]))])
b.emit_lookup_constraints()
b.emit_ctl_constraints()
msgs = [b"", b"abc", b"eth-tx-proof on B200"]
ktables, _, digests = et.keccak256_system(msgs)
tables = [("cpu_ops", b.build(), lt), ("arithmetic", et.arithmetic_program(n_limbs, limb_bits, with_ctl=True), at)] + ktables
ctls = [([0], 1), ([2], 3), ([2], 3)]
log("tables: " + ", ".join(f"{n} {t.shape[0]} x 2^{int(t.shape[1]).bit_length() - 1} ({len(p.ops)} ops)" for n, p, t in tables))
log(f"{len(ops)} arithmetic operations, keccak256 digests: " + ", ".join(d.hex()[:16] + "..." for d in digests))

# ---- table proofs on one transcript (prove_with_traces' shape)
pv = [1, 2, 3, 4, 1, 2, 3, 4]
tids = [oracle.register_table_ex(p, p.aux_spec) for _, p, _ in tables]
batches = [oracle.Batch.from_values(t, 1, 4) for _, _, t in tables]
ch = oracle.HostChallenger()
for bb in batches:
    ch.observe(bb.cap)
ch.observe(pv)
ctl_ch = ch.get_n(4)
proofs, states = [], []
for tid, (_, _, t), bb in zip(tids, tables, batches):
    states.append(ch.compact())
    proofs.append(oracle.prove_with_commitment(tid, t, bb, ch, ctl_ch))
log("four table STARKs proven (oracle)")
all_proof = types.SimpleNamespace(stark_proofs=proofs, init_challenger_states=states, ctl_challenges=ctl_ch)
DIGEST = [4, 3, 2, 1]


def circuit_prove(c, w, p):
    pr = oracle.circuit_prove(c, w, p, DIGEST)
    plonk_verifier.verify(pr, c, pr["constants_sigmas_cap"], DIGEST, max_queries=1)
    return types.SimpleNamespace(c=c, digest=DIGEST, constants_sigmas_cap=pr["constants_sigmas_cap"]), _words_from_oracle_proof(c, pr, p)


plan = sc.transaction_recursion_plan(tables, ctls, all_proof, circuit_prove, max_queries=args.queries, public_values=pv, log=log)
plan += sc.block_recursion_plan(plan[-1], len(tables), circuit_prove, levels=1, max_queries=args.queries, log=log)
log(f"block public inputs (state before ++ after ++ number): {plan[-1]['public_inputs']}")
log("every circuit proof accepted by the Python verifier; " + ", ".join(f"{s['name']}/{s['kind']} 2^{s['circuit'].degree_bits}" for s in plan))
