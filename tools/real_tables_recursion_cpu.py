"""The whole proof structure over tables with REAL semantics, on the CPU (oracle as the prover; no GPU needed):

    arithmetic, byte_packing, cpu (dispatcher), keccak (Keccak-f[1600]), keccak_sponge (Keccak-256 messages), logic, memory (port)
    wired by upstream's seven cross-table lookups (evm_tables.real_transaction_system)

seven table STARKs on one transcript -> per table the wrapper circuit (STARK verifier, the table's own constraint program
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
import oracle
import plonk_verifier
from eth_tx_proof_b200 import cprog, evm_tables as et, stark_circuit as sc
from test_circuit_cpu import _words_from_oracle_proof

ap = argparse.ArgumentParser()
ap.add_argument("--queries", type=int, default=1)
args = ap.parse_args()
t0 = time.perf_counter()
log = lambda s: print(f"[{time.perf_counter() - t0:7.1f} s] {s}", flush=True)

# ---- the seven tables
tables, ctls = et.real_transaction_system()
log("tables: " + ", ".join(f"{n} {t.shape[0]} x 2^{int(t.shape[1]).bit_length() - 1} ({len(p.ops)} ops, {len(p.ctl_zs)} CTL Zs)" for n, p, t in tables))

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
log("seven table STARKs proven (oracle)")
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
