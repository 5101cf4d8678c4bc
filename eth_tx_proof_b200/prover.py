"""Host-side mirror of the multi-table prover the reference's worker runs (`evm_arithmetization::prover::prove_with_traces`,
reached from /root/reference/ops/src/lib.rs:52 through `proof_gen::generate_txn_proof`; crate pinned at
/root/reference/Cargo.lock:1675, not on disk).  Same shape as upstream:

    trace_commitments = [PolynomialBatch::from_values(trace_t, rate_bits, false, cap_height)  for every table]
    challenger = Challenger::new();  for cap in trace_caps: challenger.observe_cap(cap)
    observe_public_values(&mut challenger, &public_values)
    ctl_challenges = get_grand_product_challenge_set(&mut challenger, num_challenges)       # (beta, gamma) x 2
    for every table, IN ORDER and on the SAME challenger:
        init_challenger_state = challenger.compact()
        proof_t = starky::prover::prove_with_commitment(stark_t, config, trace_t, commitment_t, ctl_data_t, ctl_challenges,
                                                        &mut challenger, &[], timing)

Everything heavy is a C-ABI call (commits, auxiliary/CTL columns, quotient, openings, FRI); only the transcript glue
lives here.  The tables are registered constraint programs carrying their CTL Z descriptors (cprog.ProgramBuilder.add_ctl_z).
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from .api import Challenger, Context, PolynomialBatch

RATE_BITS, CAP_HEIGHT, NUM_CHALLENGES = 1, 4, 2  # StarkConfig::standard_fast_config()


class AllProof:
    """evm_arithmetization::proof::AllProof, reduced to what this path produces: per-table StarkProofWithMetadata
    (flat proof words + init_challenger_state) and the CTL challenges."""

    def __init__(self, stark_proofs: List[np.ndarray], init_challenger_states: List[np.ndarray], ctl_challenges: np.ndarray,
                 trace_caps: List[np.ndarray]):
        self.stark_proofs = stark_proofs
        self.init_challenger_states = init_challenger_states
        self.ctl_challenges = ctl_challenges
        self.trace_caps = trace_caps


def prove_with_traces(ctx: Context, table_ids: Sequence[int], traces_dev: Sequence[tuple], public_values: Sequence[int] = ()) -> AllProof:
    """traces_dev: per table (device pointer, column stride, n_cols, log_n) of the trace values (column-major).
    Returns the per-table proofs; the challenger threading is upstream's (one transcript through all tables)."""
    commitments = [PolynomialBatch.from_values_dev(ctx, ptr, stride, n_cols, log_n, RATE_BITS, False, CAP_HEIGHT)
                   for ptr, stride, n_cols, log_n in traces_dev]
    challenger = Challenger()
    caps = [c.cap for c in commitments]
    for cap in caps:
        challenger.observe_cap(cap)
    if len(public_values):
        challenger.observe(public_values)  # observe_public_values: a flat list of field elements here
    ctl_challenges = challenger.get_n_challenges(2 * NUM_CHALLENGES)  # get_grand_product_challenge_set: beta, gamma per challenge
    proofs, states = [], []
    for tid, (ptr, stride, _, _), com in zip(table_ids, traces_dev, commitments):
        states.append(challenger.compact())
        proofs.append(ctx.prove_with_commitment(tid, com, ptr, stride, challenger, ctl_challenges))
    return AllProof(proofs, states, ctl_challenges, caps)
