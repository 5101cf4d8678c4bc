"""STARK verifier for the flat proof wire format (test infrastructure).

An independent Python restatement of the VERIFIER side of the path, used to check that proofs
produced by the oracle prover and by the CUDA prover are accepted:

  starky 0.4.0   src/verifier.rs        verify_stark_proof_with_challenges, eval_l_0_and_l_last
                 src/get_challenges.rs  get_challenges
  plonky2 0.2.2  src/fri/verifier.rs    verify_fri_proof, fri_verifier_query_round, fri_combine_initial,
                                        compute_evaluation
                 src/fri/challenges.rs  fri_challenges
                 src/hash/merkle_proofs.rs verify_merkle_proof_to_cap

(third-party crates, not on disk; pins at /root/reference/Cargo.lock:3441,4529; the reference calls the
prover at /root/reference/ops/src/lib.rs:52 and the verifier never — acceptance by this verifier is the
repo's stand-in for "every proof must pass the reference verifier").

Constraints are evaluated over the extension field with Python integers, independently of the C
oracle's and the CUDA kernels' base-field evaluators.  Hashing uses oracle.pyref (pure Python) or the
C oracle's permutation when ``fast=True``.
"""
from __future__ import annotations

import numpy as np

from oracle import pyref as R

P = R.P
MAGIC = 0x4232303053544B32  # "B200STK2"
HEADER_WORDS = 24
TABLE_FIBONACCI, TABLE_MEMORY = 0, 1
MEM_TRIE_DATA_SEGMENT = 13


class VerifyError(Exception):
    pass


def _hashers(fast):
    if not fast:
        return R.hash_or_noop, R.two_to_one, R.poseidon
    import oracle

    return (lambda x: [int(v) for v in oracle.hash_or_noop(np.array(x, dtype=np.uint64))],
            lambda l, r: [int(v) for v in oracle.two_to_one(np.array(l, dtype=np.uint64), np.array(r, dtype=np.uint64))],
            lambda s: [int(v) for v in oracle.poseidon_permute(np.array(s, dtype=np.uint64))])


class _Challenger(R.Challenger):
    def __init__(self, perm):
        super().__init__()
        self._perm = perm

    def _duplex(self):
        for i, v in enumerate(self.inb):
            self.state[i] = v
        self.inb = []
        self.state = self._perm(self.state)
        self.out = list(self.state[:8])


def parse_proof(words):
    w = [int(x) for x in words]
    if w[0] != MAGIC:
        raise VerifyError("bad magic")
    h = dict(table=w[1], degree_bits=w[2], n_trace=w[3], n_aux=w[4], n_quot=w[5], cap_height=w[6], n_layers=w[7],
             arity_bits=w[8], final_len=w[9], num_queries=w[10], n_pi=w[11], rate_bits=w[12], pow_bits=w[13],
             num_challenges=w[14], total=w[15], n_ctl_zs=w[16], n_lookup_cols=w[17], n_ctl_helpers=w[18])
    if h["total"] != len(w):
        raise VerifyError("length mismatch")
    if h["n_aux"] != h["n_lookup_cols"] + h["n_ctl_helpers"] + h["n_ctl_zs"]:
        raise VerifyError("auxiliary column counts are inconsistent")
    pos = HEADER_WORDS
    capw = 4 << h["cap_height"]

    def take(n):
        nonlocal pos
        out = w[pos:pos + n]
        if len(out) != n:
            raise VerifyError("truncated proof")
        pos += n
        return out

    def take_cap():
        c = take(capw)
        return [c[4 * i:4 * i + 4] for i in range(capw // 4)]

    def take_ext(n):
        c = take(2 * n)
        return [(c[2 * i], c[2 * i + 1]) for i in range(n)]

    pr = dict(h=h)
    pr["trace_cap"] = take_cap()
    pr["aux_cap"] = take_cap() if h["n_aux"] else None
    pr["quot_cap"] = take_cap()
    pr["local"] = take_ext(h["n_trace"])
    pr["next"] = take_ext(h["n_trace"])
    pr["aux"] = take_ext(h["n_aux"])
    pr["aux_next"] = take_ext(h["n_aux"])
    pr["ctl_zs_first"] = take(h["n_ctl_zs"])
    pr["quot"] = take_ext(h["n_quot"])
    pr["fri_caps"] = [take_cap() for _ in range(h["n_layers"])]
    log_lde = h["degree_bits"] + h["rate_bits"]
    queries = []
    for _ in range(h["num_queries"]):
        q = dict(initial=[], steps=[])
        for ncols in (h["n_trace"], h["n_aux"], h["n_quot"]):
            if ncols == 0:
                continue
            leaf = take(ncols)
            sib = take(4 * (log_lde - h["cap_height"]))
            q["initial"].append((leaf, [sib[4 * i:4 * i + 4] for i in range(len(sib) // 4)]))
        bits = log_lde
        for _l in range(h["n_layers"]):
            bits -= h["arity_bits"]
            ev = take_ext(1 << h["arity_bits"])
            sib = take(4 * (bits - h["cap_height"]))
            q["steps"].append((ev, [sib[4 * i:4 * i + 4] for i in range(len(sib) // 4)]))
        queries.append(q)
    pr["queries"] = queries
    pr["final_poly"] = take_ext(h["final_len"])
    pr["pow_witness"] = take(1)[0]
    pr["public_inputs"] = take(h["n_pi"])
    if pos != len(w):
        raise VerifyError("trailing words")
    return pr


# ---------------------------------------------------------------- constraint evaluation over F_{p^2}
class _Consumer:
    def __init__(self, alphas, z_last, l_first, l_last):
        self.alphas = [R.e_from(a) for a in alphas]
        self.acc = [(0, 0) for _ in alphas]
        self.z_last, self.l_first, self.l_last = z_last, l_first, l_last

    def constraint(self, c):
        self.acc = [R.e_add(R.e_mul(a, al), c) for a, al in zip(self.acc, self.alphas)]

    def transition(self, c):
        self.constraint(R.e_mul(c, self.z_last))

    def first_row(self, c):
        self.constraint(R.e_mul(c, self.l_first))

    def last_row(self, c):
        self.constraint(R.e_mul(c, self.l_last))


def _eval_fibonacci(lv, nv, pi, c):
    c.first_row(R.e_sub(lv[0], R.e_from(pi[0])))
    c.first_row(R.e_sub(lv[1], R.e_from(pi[1])))
    c.last_row(R.e_sub(lv[1], R.e_from(pi[2])))
    c.transition(R.e_sub(nv[0], lv[1]))
    c.transition(R.e_sub(R.e_sub(nv[1], lv[0]), lv[1]))


def _eval_memory(lv, nv, pi, c):
    one = (1, 0)
    mul, sub, add = R.e_mul, R.e_sub, R.e_add
    FILTER, TS, IS_READ, CTX, SEG, VIRT, V0 = 0, 1, 2, 3, 4, 5, 6
    CFC, SFC, VFC, INIT_AUX, RC, COUNTER = 14, 15, 16, 17, 18, 19
    f = lv[FILTER]
    c.constraint(mul(f, sub(f, one)))
    c.constraint(mul(sub(one, f), sub(one, lv[IS_READ])))
    cfc, sfc, vfc = lv[CFC], lv[SFC], lv[VFC]
    unch = sub(sub(sub(one, cfc), sfc), vfc)
    for x in (cfc, sfc, vfc, unch):
        c.constraint(mul(x, sub(one, x)))
    d_ctx, d_seg = sub(nv[CTX], lv[CTX]), sub(nv[SEG], lv[SEG])
    d_virt, d_ts = sub(nv[VIRT], lv[VIRT]), sub(nv[TS], lv[TS])
    c.transition(mul(sfc, d_ctx))
    c.transition(mul(vfc, d_ctx))
    c.transition(mul(vfc, d_seg))
    c.transition(mul(unch, d_ctx))
    c.transition(mul(unch, d_seg))
    c.transition(mul(unch, d_virt))
    computed = add(add(mul(cfc, sub(d_ctx, one)), mul(sfc, sub(d_seg, one))),
                   add(mul(vfc, sub(d_virt, one)), mul(unch, d_ts)))
    c.transition(sub(lv[RC], computed))
    ia = lv[INIT_AUX]
    c.transition(sub(ia, mul(mul(nv[SEG], sub(one, unch)), nv[IS_READ])))
    for i in range(8):
        v, nvv = lv[V0 + i], nv[V0 + i]
        c.transition(mul(mul(nv[IS_READ], unch), sub(nvv, v)))
        c.transition(mul(mul(nv[CTX], ia), nvv))
        c.transition(mul(mul(sub(nv[SEG], (MEM_TRIE_DATA_SEGMENT, 0)), ia), nvv))
    c.first_row(lv[COUNTER])
    c.transition(sub(sub(nv[COUNTER], lv[COUNTER]), one))


def _eval_memory_lookups(lv, aux, aux_next, challenges, c):
    RC, COUNTER, FREQ = 18, 19, 20
    start = 0
    for ch in challenges:
        che = R.e_from(ch)
        h, z, nz = aux[start], aux[start + 1], aux_next[start + 1]
        c.constraint(R.e_sub(R.e_mul(R.e_add(lv[RC], che), h), (1, 0)))
        twc = R.e_add(lv[COUNTER], che)
        y = R.e_sub(R.e_mul(h, twc), lv[FREQ])
        c.first_row(z)
        c.constraint(R.e_sub(R.e_mul(R.e_sub(nz, z), twc), y))
        start += 2


def _merkle_verify(leaf, index, siblings, cap, hash_or_noop, two_to_one):
    cur = hash_or_noop(leaf)
    for sib in siblings:
        cur = two_to_one(sib, cur) if index & 1 else two_to_one(cur, sib)
        index >>= 1
    if cur != cap[index]:
        raise VerifyError("invalid Merkle proof")


def _reduce(alpha, vals):
    acc = (0, 0)
    for v in reversed(vals):
        acc = R.e_add(R.e_mul(acc, alpha), v)
    return acc


def _interpolate(points, x):
    total = (0, 0)
    for i, (xi, yi) in enumerate(points):
        num, den = (1, 0), (1, 0)
        for j, (xj, _) in enumerate(points):
            if i != j:
                num = R.e_mul(num, R.e_sub(x, xj))
                den = R.e_mul(den, R.e_sub(xi, xj))
        total = R.e_add(total, R.e_mul(yi, R.e_mul(num, R.e_inv(den))))
    return total


def _eval_program(program, lv, nv, aux, aux_next, pi, challenges, c):
    """Program-defined table (eth_tx_proof_b200/cprog.py): the program is interpreted over F_{p^2}."""
    out = program.evaluate(lv, nv, aux, aux_next, pi, challenges or (), add=R.e_add, sub=R.e_sub, mul=R.e_mul, lift=R.e_from)
    for kind, val in out:
        (c.constraint, c.transition, c.first_row, c.last_row)[kind - 10](val)


def new_challenger(fast=True):
    return _Challenger(_hashers(fast)[2])


def verify_cross_table_lookups(ctls, ctl_zs_first, num_challenges=2):
    """starky::cross_table_lookup::verify_cross_table_lookups.  ctls: [(looking table indices (with repeats), looked table
    index)]; ctl_zs_first: per table, the list of openings.  For every CTL and challenge the looking tables' Z(1) must sum
    to the looked table's."""
    its = [iter(v) for v in ctl_zs_first]
    for index, (looking, looked) in enumerate(ctls):
        uniq = []
        for t in looking:
            if t not in uniq:
                uniq.append(t)
        for _ in range(num_challenges):
            total = sum(next(its[t]) for t in uniq) % P
            if total != next(its[looked]):
                raise VerifyError(f"Cross-table lookup {index} verification failed.")
    for it in its:
        if next(it, None) is not None:
            raise VerifyError("unused ctl_zs_first openings")


def verify(words, fast=True, max_queries=None, program=None, challenger=None, ctl_challenges=None):
    """Raises VerifyError unless the proof is valid. Returns the parsed proof.  `program`: the cprog.Program of a
    registered table (table id >= 16 in the header); built-in tables are evaluated by the code above.
    Multi-table mode (evm_arithmetization verify_proof -> verify_stark_proof_with_challenges): pass the shared
    `challenger` (it has observed every trace cap and produced `ctl_challenges` = [(beta, gamma)] * num_challenges) —
    the trace cap and public inputs are then NOT observed again and the lookup challenges are the CTL betas."""
    hash_or_noop, two_to_one, perm = _hashers(fast)
    pr = parse_proof(words)
    h = pr["h"]
    table, db, rate_bits = h["table"], h["degree_bits"], h["rate_bits"]
    n_ch = h["num_challenges"]
    degree = 1 << db
    if table >= 16:
        if program is None:
            raise VerifyError("registered table: pass its program")
        factor = max(1, program.degree - 1)
        chunk = max(1, program.degree - 1)
        exp_cols = (program.n_trace, program.n_lookup_cols + program.n_ctl_helper_cols + len(program.ctl_zs))
        if (h["n_lookup_cols"], h["n_ctl_helpers"], h["n_ctl_zs"]) != (program.n_lookup_cols, program.n_ctl_helper_cols, len(program.ctl_zs)):
            raise VerifyError("shape (auxiliary columns)")
    else:
        factor = {TABLE_FIBONACCI: 1, TABLE_MEMORY: 2}[table]
        exp_cols = {TABLE_FIBONACCI: (2, 0), TABLE_MEMORY: (21, 2 * n_ch)}[table]
    if (h["n_trace"], h["n_aux"]) != exp_cols or h["n_quot"] != factor * n_ch:
        raise VerifyError("shape")
    # ---- transcript (get_challenges)
    if challenger is None:
        ch = _Challenger(perm)
        ch.observe(pr["public_inputs"])
        for d in pr["trace_cap"]:
            ch.observe(d)
    else:
        ch = challenger
    if h["n_ctl_zs"] and ctl_challenges is None:
        raise VerifyError("the proof carries CTL openings but no CTL challenges were given")
    lookup_ch = None
    if pr["aux_cap"] is not None:
        if ctl_challenges is not None:
            lookup_ch = [int(b) for b, _ in ctl_challenges]
        else:
            raw = ch.get_n(2 * n_ch)
            lookup_ch = raw[0::2]
        for d in pr["aux_cap"]:
            ch.observe(d)
    scalars = list(lookup_ch or [0] * n_ch)
    if ctl_challenges is not None:
        for b, g_ in ctl_challenges:
            scalars += [int(b), int(g_)]
    alphas = ch.get_n(n_ch)
    for d in pr["quot_cap"]:
        ch.observe(d)
    zeta = ch.get_ext()
    zeta_batch = pr["local"] + pr["aux"] + pr["quot"]
    next_batch = pr["next"] + pr["aux_next"]
    ctl_batch = [(v, 0) for v in pr["ctl_zs_first"]]
    for batch in (zeta_batch, next_batch, ctl_batch):
        for e in batch:
            ch.observe(e)
    fri_alpha = ch.get_ext()
    betas = []
    for cap in pr["fri_caps"]:
        for d in cap:
            ch.observe(d)
        betas.append(ch.get_ext())
    for e in pr["final_poly"]:
        ch.observe(e)
    ch.observe([pr["pow_witness"]])
    pow_response = ch.get()
    lde_bits = db + rate_bits
    lde_size = 1 << lde_bits
    query_indices = [ch.get() % lde_size for _ in range(h["num_queries"])]

    # ---- constraint check at zeta
    g = R.root_of_unity(db)
    zeta_pow = R.e_pow(zeta, degree)
    z_x = R.e_sub(zeta_pow, (1, 0))
    l_first = R.e_mul(z_x, R.e_inv(R.e_scalar(R.e_sub(zeta, (1, 0)), degree)))
    l_last = R.e_mul(z_x, R.e_inv(R.e_scalar(R.e_sub(R.e_scalar(zeta, g), (1, 0)), degree)))
    z_last = R.e_sub(zeta, R.e_from(pow(g, P - 2, P)))
    cons = _Consumer(alphas, z_last, l_first, l_last)
    if table >= 16:
        _eval_program(program, pr["local"], pr["next"], pr["aux"], pr["aux_next"], pr["public_inputs"], scalars, cons)
    elif table == TABLE_FIBONACCI:
        _eval_fibonacci(pr["local"], pr["next"], pr["public_inputs"], cons)
    else:
        _eval_memory(pr["local"], pr["next"], pr["public_inputs"], cons)
        _eval_memory_lookups(pr["local"], pr["aux"], pr["aux_next"], lookup_ch, cons)
    for i in range(n_ch):
        chunk = pr["quot"][i * factor:(i + 1) * factor]
        if cons.acc[i] != R.e_mul(z_x, _reduce(zeta_pow, chunk)):
            raise VerifyError("Mismatch between evaluation and opening of quotient polynomial")

    # ---- FRI
    if h["pow_bits"] and (pow_response >> (64 - h["pow_bits"])) != 0:
        raise VerifyError("Invalid proof of work witness")
    zeta_next = R.e_scalar(zeta, g)
    reduced_openings = [_reduce(fri_alpha, zeta_batch), _reduce(fri_alpha, next_batch), _reduce(fri_alpha, ctl_batch)]
    z_first_col = h["n_lookup_cols"] + h["n_ctl_helpers"]
    caps = [pr["trace_cap"]] + ([pr["aux_cap"]] if pr["aux_cap"] is not None else []) + [pr["quot_cap"]]
    n_trace, n_aux = h["n_trace"], h["n_aux"]
    w_lde = R.root_of_unity(lde_bits)
    arity_bits = h["arity_bits"]
    arity = 1 << arity_bits
    nq = h["num_queries"] if max_queries is None else min(max_queries, h["num_queries"])
    for qi in range(nq):
        x_index = query_indices[qi]
        q = pr["queries"][qi]
        for (leaf, sib), cap in zip(q["initial"], caps):
            _merkle_verify(leaf, x_index, sib, cap, hash_or_noop, two_to_one)
        subgroup_x = R.GENERATOR * pow(w_lde, R.bitrev(x_index, lde_bits), P) % P
        # fri_combine_initial
        leaves = [lf for lf, _ in q["initial"]]
        trace_ev = [R.e_from(v) for v in leaves[0]]
        aux_ev = [R.e_from(v) for v in leaves[1]] if n_aux else []
        quot_ev = [R.e_from(v) for v in leaves[-1]]
        sx = R.e_from(subgroup_x)
        total = (0, 0)
        fri_batches = [(trace_ev + aux_ev + quot_ev, zeta, reduced_openings[0]), (trace_ev + aux_ev, zeta_next, reduced_openings[1])]
        if h["n_ctl_zs"]:
            fri_batches.append((aux_ev[z_first_col:], (1, 0), reduced_openings[2]))
        for evals, point, red in fri_batches:
            numerator = R.e_sub(_reduce(fri_alpha, evals), red)
            denominator = R.e_sub(sx, point)
            total = R.e_mul(total, R.e_pow(fri_alpha, len(evals)))
            total = R.e_add(total, R.e_mul(numerator, R.e_inv(denominator)))
        old_eval = total
        for i, (evals, sib) in enumerate(q["steps"]):
            coset_index = x_index >> arity_bits
            within = x_index & (arity - 1)
            if evals[within] != old_eval:
                raise VerifyError(f"FRI consistency check failed (query {qi}, layer {i})")
            # compute_evaluation
            ga = R.root_of_unity(arity_bits)
            ev = [evals[R.bitrev(k, arity_bits)] for k in range(arity)]
            rev_within = R.bitrev(within, arity_bits)
            coset_start = subgroup_x * pow(ga, arity - rev_within, P) % P
            pts = [(R.e_from(coset_start * pow(ga, k, P) % P), ev[k]) for k in range(arity)]
            old_eval = _interpolate(pts, betas[i])
            flat = [c for e in evals for c in e]
            _merkle_verify(flat, coset_index, sib, pr["fri_caps"][i], hash_or_noop, two_to_one)
            subgroup_x = pow(subgroup_x, arity, P)
            x_index = coset_index
        acc = (0, 0)
        for c in reversed(pr["final_poly"]):
            acc = R.e_add(R.e_mul(acc, R.e_from(subgroup_x)), c)
        if acc != old_eval:
            raise VerifyError("Final polynomial evaluation is invalid.")
    return pr
