"""First slice of the recursion layers: the commitment / opening / FRI skeleton of plonky2's circuit prover
(`plonky2::plonk::prover::prove`, plonky2 0.2.2 — /root/reference/Cargo.lock:3441), which is what every shrink, root,
aggregation and block proof of the reference runs (/root/reference/ops/src/lib.rs:52 -> 7 shrink chains + root circuit,
:72 AggProof, :95 BlockProof; /root/reference/leader/src/prover.rs:26-36).

What is on the device here, under `CircuitConfig::standard_recursion_config()` (rate_bits 3, cap_height 4, 28 query rounds,
16 PoW bits, ConstantArityBits(4, 5), 135 wires of which 80 routed, 2 challenges, quotient degree factor 8):

    wires_commitment          = PolynomialBatch::from_values(wires, rate_bits, blinding=false, cap_height)        135 polys
    all_wires_permutation_partial_products(witness, betas, gammas, ..)       Z and partial products of the permutation argument
    partial_products_and_zs   = PolynomialBatch::from_values(...)                          2 x (1 Z + 9 partial products) = 20 polys
    quotient_polys_commitment = PolynomialBatch::from_coeffs(quotient chunks)                                  2 x 8 = 16 polys
    openings at zeta (all four oracles) and g*zeta (the Zs), observe, PolynomialBatch::prove_openings over the
    four-oracle FriInstanceInfo (constants_sigmas is the circuit's pre-committed batch: 4 constants + 80 sigmas)

through the same C-ABI calls the starky path uses (`etp_batch_from_*`, `etp_batch_eval_at_ext_point`, `etp_prove_openings`)
and `etp_plonk_partial_products_and_zs_dev`, with a caller-side `Challenger`.  What is NOT: witness generation and the
gate-constraint quotient — they stay on the CPU side of the fork for now (gate evaluation is the next slice); this module feeds seeded
stand-in polynomials of the right SHAPES, so the numbers it produces time the device skeleton of a recursion proof, not a
recursion proof.  The transcript order follows plonk/prover.rs: circuit digest, public-input hash, wires cap -> betas,
gammas; Z cap -> alphas; quotient cap -> zeta; openings -> FRI.
"""
from __future__ import annotations

import time
from typing import Dict

import numpy as np

from .api import Challenger, Context, FriParams, PolynomialBatch
from . import synthetic as syn

P = 0xFFFFFFFF00000001
NUM_WIRES, NUM_ROUTED, NUM_CONSTANTS, NUM_CHALLENGES, QUOTIENT_DEGREE_FACTOR = 135, 80, 4, 2, 8
NUM_PARTIAL_PRODUCTS = -(-NUM_ROUTED // QUOTIENT_DEGREE_FACTOR) - 1  # 9
RATE_BITS, CAP_HEIGHT, POW_BITS, NUM_QUERIES = 3, 4, 16, 28
ORACLE_SHAPES = {"constants_sigmas": NUM_CONSTANTS + NUM_ROUTED, "wires": NUM_WIRES,
                 "zs_partial_products": NUM_CHALLENGES * (1 + NUM_PARTIAL_PRODUCTS), "quotient": NUM_CHALLENGES * QUOTIENT_DEGREE_FACTOR}


def root_of_unity(n_log: int) -> int:
    return pow(1753635133440165772, 1 << (32 - n_log), P)


def fri_instance(zeta, degree_bits: int):
    """plonky2::plonk::circuit_data::CommonCircuitData::get_fri_instance: the zeta batch opens every polynomial of the four
    oracles, the g*zeta batch the Z polynomials (oracle 2, the first num_challenges columns)."""
    g = root_of_unity(degree_bits)
    zeta_next = [int(zeta[0]) * g % P, int(zeta[1]) * g % P]
    all_polys = [(o, c) for o, n in enumerate(ORACLE_SHAPES.values()) for c in range(n)]
    zs = [(2, c) for c in range(NUM_CHALLENGES)]
    return [([int(zeta[0]), int(zeta[1])], all_polys), (zeta_next, zs)]


def coset_shifts(num_shifts: int):
    """plonky2::field::cosets::get_unique_coset_shifts: k_j = g^j for the multiplicative generator g = 7 (k_0 = 1)."""
    return [pow(7, j, P) for j in range(num_shifts)]


def permutation_witness(degree_bits: int, num_routed: int = NUM_ROUTED, seed: int = 1):
    """A wire assignment that satisfies a random copy-constraint permutation, with the sigma values that encode it.
    Positions (row i, routed column j) are paired at random; both wires of a pair hold the same value and sigma maps each
    to the other's identity value k_j * g^i (plonky2::plonk::permutation_argument / CircuitBuilder::sigma_vecs).
    -> (wires (num_routed, n), sigmas (num_routed, n), k_is)"""
    n = 1 << degree_bits
    rng = np.random.default_rng(seed)
    g = root_of_unity(degree_bits)
    k_is = coset_shifts(num_routed)
    xs = [1]
    for _ in range(n - 1):
        xs.append(xs[-1] * g % P)
    k_obj, x_obj = np.array(k_is, dtype=object), np.array(xs, dtype=object)
    total = n * num_routed  # position = column * n + row
    order = rng.permutation(total)
    a, b = order[0:total - total % 2:2], order[1::2]
    ident = lambda pos: np.array(k_obj[pos // n] * x_obj[pos % n] % P, dtype=np.uint64)
    vals = rng.integers(0, 2**63, size=a.size, dtype=np.uint64)
    wires, sig = np.zeros(total, dtype=np.uint64), np.zeros(total, dtype=np.uint64)
    wires[a] = wires[b] = vals
    sig[a], sig[b] = ident(b), ident(a)
    if total % 2:
        last = order[-1:]
        wires[last] = 7
        sig[last] = ident(last)
    return wires.reshape(num_routed, n), sig.reshape(num_routed, n), np.array(k_is, dtype=np.uint64)


def stand_in_polys(degree_bits: int, seed: int = 0xC1C) -> Dict[str, np.ndarray]:
    """Seeded inputs of one skeleton proof: the routed wires satisfy a random permutation argument (so Z closes), the
    advice wires, constants and quotient chunks are random stand-ins; `sigmas` are the sigma VALUES on the subgroup (the
    constants_sigmas oracle commits to them), `k_is` the coset shifts."""
    wires_routed, sigmas, k_is = permutation_witness(degree_bits, NUM_ROUTED, seed)
    polys = {name: syn.random_columns(n, degree_bits, seed=seed + 1000 * i) for i, (name, n) in enumerate(ORACLE_SHAPES.items())}
    polys["wires"][:NUM_ROUTED] = wires_routed
    polys["constants_sigmas"][NUM_CONSTANTS:] = sigmas
    polys["k_is"] = k_is
    return polys


def prove_skeleton(ctx: Context, degree_bits: int, polys: Dict[str, np.ndarray] = None, constants_sigmas: PolynomialBatch = None) -> dict:
    """One recursion-proof skeleton on the device.  Returns the caps, the openings, the flat FriProof and per-step wall times
    (ms).  `constants_sigmas`: the circuit's pre-committed batch (built once per circuit, outside the proof)."""
    polys = polys or stand_in_polys(degree_bits)
    if constants_sigmas is None:
        constants_sigmas = PolynomialBatch.from_values(ctx, polys["constants_sigmas"], RATE_BITS, False, CAP_HEIGHT)
    t = {}
    ch = Challenger()
    ch.observe(constants_sigmas.cap)  # stands for the circuit digest (which commits to this cap) + public-input hash
    t0 = time.perf_counter()
    wires = PolynomialBatch.from_values(ctx, polys["wires"], RATE_BITS, False, CAP_HEIGHT)
    t["wires commit"] = (time.perf_counter() - t0) * 1e3
    ch.observe_cap(wires.cap)
    betas, gammas = ch.get_n_challenges(NUM_CHALLENGES), ch.get_n_challenges(NUM_CHALLENGES)
    # all_wires_permutation_partial_products on the device: routed wires + sigma values -> Z and partial products
    import torch

    n = 1 << degree_bits
    t0 = time.perf_counter()
    d_w = torch.from_numpy(np.ascontiguousarray(polys["wires"][:NUM_ROUTED]).view(np.int64)).cuda()
    d_s = torch.from_numpy(np.ascontiguousarray(polys["constants_sigmas"][NUM_CONSTANTS:]).view(np.int64)).cuda()
    d_z = torch.empty((ORACLE_SHAPES["zs_partial_products"], n), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    t["upload routed wires + sigmas"] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    ctx.plonk_partial_products_and_zs_dev(d_w.data_ptr(), n, d_s.data_ptr(), n, polys["k_is"], degree_bits, QUOTIENT_DEGREE_FACTOR,
                                          betas, gammas, d_z.data_ptr())
    t["partial products and Zs"] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    zs = PolynomialBatch.from_values_dev(ctx, d_z.data_ptr(), n, ORACLE_SHAPES["zs_partial_products"], degree_bits, RATE_BITS, False, CAP_HEIGHT)
    t["partial products and Zs commit"] = (time.perf_counter() - t0) * 1e3
    zs_values = d_z.cpu().numpy().view(np.uint64)
    ch.observe_cap(zs.cap)
    alphas = ch.get_n_challenges(NUM_CHALLENGES)
    t0 = time.perf_counter()
    quotient = PolynomialBatch.from_coeffs(ctx, polys["quotient"], RATE_BITS, False, CAP_HEIGHT)
    t["quotient commit"] = (time.perf_counter() - t0) * 1e3
    ch.observe_cap(quotient.cap)
    zeta = ch.get_extension_challenge()
    oracles = [constants_sigmas, wires, zs, quotient]
    g = root_of_unity(degree_bits)
    zeta_next = [int(zeta[0]) * g % P, int(zeta[1]) * g % P]
    t0 = time.perf_counter()
    openings = [o.eval_at_ext_point(zeta) for o in oracles]
    openings_next = zs.eval_at_ext_point(zeta_next)[:NUM_CHALLENGES]
    t["openings"] = (time.perf_counter() - t0) * 1e3
    for o in openings:
        ch.observe(o)
    ch.observe(openings_next)
    fp = FriParams.make(degree_bits, RATE_BITS, CAP_HEIGHT, POW_BITS, NUM_QUERIES)
    t0 = time.perf_counter()
    fri = ctx.prove_openings(fri_instance(zeta, degree_bits), oracles, ch, fp)
    t["prove_openings (FRI)"] = (time.perf_counter() - t0) * 1e3
    t["total"] = sum(t.values())
    return {"zs_partial_products": zs_values, "caps": [o.cap for o in oracles], "betas": betas, "gammas": gammas, "alphas": alphas, "zeta": zeta, "openings": openings,
            "openings_next": openings_next, "fri_proof": fri, "challenger": ch, "ms": t, "fri_params": fp}
