"""One process per GPU: job sharding and the two small reductions the multi-GPU runs need.

The path shards with no data-path collective: per-table proofs and a block's
transaction segments are independent, exactly like the reference's job graph
`IndexedStream::from(txs).map(&TxProof).fold(&AggProof)` (/root/reference/leader/src/prover.rs:26-30)
that Paladin spreads over workers, one worker pinned to one GPU
(/root/reference/deploy/paladin-worker@.service:10).  `torch.distributed` (NCCL on GPUs, gloo in the
CPU tests) is only plumbing: a barrier, a MAX over per-rank device times, and the gather of the
KB-sized results (caps / proofs) to the leader rank — the analogue of results travelling back
through Paladin.
"""
from __future__ import annotations

from typing import Any, List, Sequence, Tuple


def shard_jobs(n_jobs: int, rank: int, world: int) -> List[int]:
    """Round-robin assignment of independent jobs (segments / tables) to ranks."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_jobs, world))


def max_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_results(local: Sequence[Any]) -> List[Any]:
    """All ranks' (job_id, result) pairs, merged and ordered by job id (every rank gets the list)."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return sorted(local, key=lambda kv: kv[0])
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, list(local))
    merged = [kv for part in out for kv in part]
    return sorted(merged, key=lambda kv: kv[0])


# ---- column-split commit of one oversized table ---------------------------------------
def cols_per_rank(n_cols: int, world: int) -> int:
    """Columns per shard: ceil(n_cols / world) rounded up to the sponge rate (8), as etp_shard_cols_per_rank."""
    cps = -(-n_cols // world)
    return -(-cps // 8) * 8


def column_split_plan(n_cols: int, lde_rows: int, cap_height: int, rank: int, world: int) -> dict:
    """What rank `rank` owns: a column range (a multiple of 8 wide, so that no 8-column sponge chunk straddles
    two GPUs), a leaf-row range and the cap entries of those rows (whole cap subtrees)."""
    if world < 1 or world & (world - 1) or world > 8:
        raise ValueError("world must be a power of two <= 8")
    if (1 << cap_height) < world:
        raise ValueError("2^cap_height must be >= world")
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    cps = cols_per_rank(n_cols, world)
    c0 = min(cps * rank, n_cols)
    c1 = min(c0 + cps, n_cols)
    rows = lde_rows // world
    caps = (1 << cap_height) // world
    return {"cols": (c0, c1), "rows": (rows * rank, rows * (rank + 1)), "cap_entries": (caps * rank, caps * (rank + 1))}


def assemble_cap(parts: Sequence[Any]):
    """Cap of the whole table = the ranks' parts in rank order."""
    import numpy as np

    return np.concatenate([np.asarray(p, dtype=np.uint64).reshape(-1, 4) for p in parts], axis=0)


def commit_column_split(shard, local_values=None, values_dev: Tuple[int, int] = None):
    """Runs the column-split commit protocol on this rank (call on every rank of the default process group):
    local transforms -> IPC handle exchange -> barrier -> leaf hashing over NVLink peer loads + own subtrees ->
    all-gather of the cap parts.  Returns the whole cap (2^cap_height x 4).  The peers' mappings stay open
    (needed for `leaves_at`); call `finish_column_split(shard)` before freeing the shard."""
    import torch.distributed as dist

    if values_dev is not None:
        shard.transform_values_dev(*values_dev)
    else:
        shard.transform_values(local_values)
    handles = [None] * shard.world
    dist.all_gather_object(handles, shard.export_handle())
    for r, hnd in enumerate(handles):
        if r != shard.rank and column_split_plan(shard.n_cols_total, 8, shard.cap_height, r, shard.world)["cols"][0] < shard.n_cols_total:
            shard.open_peer(r, hnd)
    dist.barrier()  # every LDE matrix complete (the transforms synchronise their stream) and mapped
    part = shard.commit_rows()
    parts = [None] * shard.world
    dist.all_gather_object(parts, part.tolist())
    return assemble_cap(parts)


def recommit_column_split(shard, values_dev: Tuple[int, int]):
    """Second and later commits into a shard whose peers are already mapped (the LDE buffers persist, so the IPC
    handles stay valid): local transforms -> barrier -> leaf hashing over NVLink + own subtrees -> NCCL all-gather
    of the cap parts.  Returns the whole cap."""
    import numpy as np
    import torch
    import torch.distributed as dist

    dist.barrier()  # peers have finished reading the previous LDE
    shard.transform_values_dev(*values_dev)
    dist.barrier()
    part = torch.from_numpy(shard.commit_rows().view(np.int64)).cuda()
    whole = torch.empty((shard.world * part.shape[0], 4), dtype=torch.int64, device="cuda")
    dist.all_gather_into_tensor(whole, part)
    return whole.cpu().numpy().view(np.uint64)


def finish_column_split(shard):
    import torch.distributed as dist

    dist.barrier()  # nobody is still reading this rank's LDE
    shard.close_peers()


# ---- proof of ONE column-split table (starky prove() for a table without lookups / CTLs) --------------------------
P = 0xFFFFFFFF00000001
STARK_RATE_BITS, STARK_CAP_HEIGHT, STARK_POW_BITS, STARK_NUM_QUERIES, STARK_NUM_CHALLENGES = 1, 4, 16, 84, 2


def _ext_mul(a, b):
    """GoldilocksField quadratic extension, X^2 = 7."""
    return [(a[0] * b[0] + 7 * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P]


class ThreadComm:
    """all_gather between the threads of ONE process, each driving one shard (a context per GPU, or several contexts on
    one GPU in the tests): `comm = ThreadComm(world)`, rank r calls `comm.rank(r).all_gather(obj)`."""

    def __init__(self, world: int):
        import threading

        self.world = world
        self._slots = [None] * world
        self._barrier = threading.Barrier(world)

    def rank(self, r: int):
        outer = self

        class _Rank:
            def all_gather(self, obj):
                outer._slots[r] = obj
                outer._barrier.wait()
                out = list(outer._slots)
                outer._barrier.wait()
                return out

        return _Rank()


def _all_gather(obj, comm=None):
    if comm is not None:
        return comm.all_gather(obj)
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def _collective(fn, comm):
    """Runs fn() on this rank and all-gathers the results.  A failure on ANY rank is raised on EVERY rank (as EtpError with
    the first failing rank's code and message) instead of leaving the others blocked in the next collective."""
    from .api import EtpError

    try:
        val, err = fn(), None
    except Exception as e:  # noqa: BLE001
        val, err = None, (getattr(e, "code", -3), f"{type(e).__name__}: {e}")
    out = _all_gather((err, val), comm)
    for r, (e, _) in enumerate(out):
        if e:
            raise EtpError(e[0], f"column-split proof, rank {r}: {e[1]}")
    return [v for _, v in out]


def prove_column_split(shard, table: int, cap, public_inputs: Sequence[int] = (), leader: int = None, timings: dict = None, comm=None,
                       challenger=None, ctl_challenges=None):
    """starky::prover::prove (or, with `challenger`, prove_with_commitment) for ONE table whose trace is column-split over
    the ranks (call on every rank, after `commit_column_split(shard, ...)` returned the trace `cap`).  Returns the proof
    words ("B200STK2", the layout of `Context.stark_prove` / `Context.prove_with_commitment`) on the leader rank and None
    on the others; for the same trace it is word for word the proof one GPU produces.

    Who does what.  Every rank keeps only its columns (coefficients + LDE) and its leaf rows' Merkle subtrees.  The
    leader (default: the last rank, which owns the fewest columns) runs the transcript and every step that needs whole
    rows — the quotient, the FRI combination — reading the peers' LDE columns in place over NVLink (the mappings of the
    commit); the openings at zeta and g*zeta are evaluated where the coefficients live and gathered (2 x 16 bytes per
    column); the Merkle paths of the queried trace rows come from the ranks that own those rows.  Nothing of the trace is
    ever copied between GPUs as a whole.  Lookups / CTLs: the auxiliary polynomials are few; the leader computes them from
    the handful of trace columns they read (recovered from the mapped LDE), commits them as a batch of its own.

    `challenger` (leader only; an api.Challenger that has already observed what upstream's caller observed, e.g. every
    table's trace cap in prove_with_traces) is advanced in place; without it the transcript starts as prove() does
    (public inputs, trace cap).  `ctl_challenges`: the 4 words (beta, gamma) x num_challenges of a multi-table proof.
    `comm`: an object with all_gather(obj) -> list (ThreadComm.rank(r)) when the ranks are threads of one process;
    default: the torch.distributed default group."""
    import time

    import numpy as np

    from .api import Challenger, EtpError, FriParams, PolynomialBatch
    from . import wire

    ctx, rank, world = shard.ctx, shard.rank, shard.world
    leader = world - 1 if leader is None else leader
    log_n, n_cols = shard.degree_log, shard.n_cols_total
    if shard.rate_bits != STARK_RATE_BITS or shard.cap_height != STARK_CAP_HEIGHT:
        raise ValueError("a STARK proof needs the shard committed with rate_bits 1 and cap_height 4 (StarkConfig::standard_fast_config)")
    n, lde_n = 1 << log_n, 1 << (log_n + STARK_RATE_BITS)
    L = ctx.L
    if L.etp_table_num_columns(ctx.h, table) != n_cols:
        raise ValueError("the table does not have the shard's number of columns")
    n_pi = int(L.etp_table_num_public_inputs(ctx.h, table))
    if len(public_inputs) < n_pi:
        raise ValueError("too few public inputs")
    pi = [int(x) % P for x in public_inputs[:n_pi]]
    K = STARK_NUM_CHALLENGES
    n_quot = int(L.etp_table_quotient_degree_factor(ctx.h, table)) * K
    n_aux = int(L.etp_table_num_aux_columns(ctx.h, table, K))
    n_lookup = int(L.etp_table_num_lookup_columns(ctx.h, table, K))
    n_helpers = int(L.etp_table_num_ctl_helper_columns(ctx.h, table))
    n_zs = int(L.etp_table_num_ctl_zs(ctx.h, table))
    if n_zs and ctl_challenges is None:
        raise EtpError(-1, "the table requires CTLs: pass the CTL challenges (and the shared challenger)")
    if ctl_challenges is not None:
        ctl_challenges = [int(x) % P for x in np.asarray(ctl_challenges, dtype=np.uint64).ravel()]
        if len(ctl_challenges) != 2 * K:
            raise EtpError(-1, "ctl_challenges must be num_challenges (beta, gamma) pairs = 4 words")
    fp = FriParams.make(log_n, STARK_RATE_BITS, STARK_CAP_HEIGHT, STARK_POW_BITS, STARK_NUM_QUERIES)
    cap = np.asarray(cap, dtype=np.uint64).reshape(-1, 4)
    t_last = [time.perf_counter()]

    def mark(name):
        if timings is not None:
            ctx.synchronize()
            now = time.perf_counter()
            timings[name] = timings.get(name, 0.0) + (now - t_last[0]) * 1e3
            t_last[0] = now

    # ---- leader: transcript up to zeta; auxiliary polynomials; quotient over the split trace; their commitments
    st = {"zs_first": np.zeros(0, dtype=np.uint64), "aux": None}

    def leader_until_zeta():
        if rank != leader:
            return None
        import ctypes as C

        def dev_matrix(words):
            d = C.c_void_p()
            ctx.check(L.etp_dev_alloc(ctx.h, max(words, 1) * 8, C.byref(d)))
            return d

        if challenger is None:
            ch = Challenger()
            ch.observe(pi)
            ch.observe_cap(cap)
        else:
            ch = challenger
        st["ch"] = ch
        # lookup challenges: the CTL betas when CTL challenges are given, else get_grand_product_challenge_set's betas
        scalars = []
        if n_lookup:
            for k in range(K):
                if ctl_challenges is not None:
                    scalars.append(ctl_challenges[2 * k])
                else:
                    scalars.append(ch.get_challenge())
                    ch.get_challenge()
        lookup_ch = list(scalars)
        if ctl_challenges is not None:
            scalars = (scalars + [0] * K)[:K] + ctl_challenges
        if n_aux:
            d = dev_matrix(n_aux * n)
            try:
                st["zs_first"] = shard.aux_columns_dev(table, lookup_ch if n_lookup else [0] * K, ctl_challenges, d.value)
                mark("compute auxiliary columns (their trace columns recovered from the LDE over NVLink)")
                st["aux"] = PolynomialBatch.from_values_dev(ctx, d.value, n, n_aux, log_n, STARK_RATE_BITS, False, STARK_CAP_HEIGHT)
            finally:
                L.etp_dev_free(ctx.h, d)
            mark("auxiliary polys commit")
            ch.observe_cap(st["aux"].cap)
        alphas = ch.get_n_challenges(K)
        d = dev_matrix(n_quot * n)
        try:
            shard.compute_quotient_polys_dev(table, st["aux"], scalars, pi, alphas, d.value)
            mark("compute quotient polys (trace columns over NVLink)")
            st["quot"] = PolynomialBatch.from_coeffs_dev(ctx, d.value, n, n_quot, log_n, STARK_RATE_BITS, False, STARK_CAP_HEIGHT)
        finally:
            L.etp_dev_free(ctx.h, d)
        mark("quotient polys commit")
        ch.observe_cap(st["quot"].cap)
        zeta = [int(x) for x in ch.get_extension_challenge()]
        zp = zeta
        for _ in range(log_n):
            zp = _ext_mul(zp, zp)
        if zp == [1, 0]:
            raise EtpError(-4, "Opening point is in the subgroup.")
        g = pow(1753635133440165772, 1 << (32 - log_n), P)
        return {"zeta": zeta, "zeta_next": [zeta[0] * g % P, zeta[1] * g % P]}

    msg = _collective(leader_until_zeta, comm)[leader]
    zeta, zeta_next = msg["zeta"], msg["zeta_next"]

    # ---- every rank: openings of its own columns, gathered
    def local_openings():
        e0, e1 = shard.eval_at_ext_points(zeta, zeta_next)
        return e0.tolist(), e1.tolist()

    parts = _collective(local_openings, comm)
    mark("evaluate the local columns at zeta, g*zeta")

    def leader_until_queries():
        if rank != leader:
            return None
        ch, aux, quot, zs_first = st["ch"], st["aux"], st["quot"], st["zs_first"]
        ext0 = np.zeros((0, 2), dtype=np.uint64)
        tr0 = np.array([v for part in parts for v in part[0]], dtype=np.uint64).reshape(-1, 2)
        tr1 = np.array([v for part in parts for v in part[1]], dtype=np.uint64).reshape(-1, 2)
        assert tr0.shape[0] == n_cols and tr1.shape[0] == n_cols
        ax0 = aux.eval_at_ext_point(zeta).reshape(-1, 2) if aux is not None else ext0
        ax1 = aux.eval_at_ext_point(zeta_next).reshape(-1, 2) if aux is not None else ext0
        qu0 = quot.eval_at_ext_point(zeta).reshape(-1, 2)
        st["openings"] = (tr0, tr1, ax0, ax1, qu0)
        # observe_openings(to_fri_openings): zeta batch = local ++ aux ++ quotient, next batch = next ++ aux_next, ctl_zs_first
        for v in (tr0, ax0, qu0, tr1, ax1):
            ch.observe(v)
        for z in zs_first:
            ch.observe([int(z), 0])
        # stark.fri_instance + prove_openings: oracle 0 = trace (split), then the auxiliary batch (if any), then the quotient
        alpha = ch.get_extension_challenge()
        extra = [aux, quot] if aux is not None else [quot]
        o_aux, o_quot = 1, len(extra)
        trace_polys = [(0, c) for c in range(n_cols)]
        aux_polys = [(o_aux, c) for c in range(n_aux)]
        batches = [(zeta, trace_polys + aux_polys + [(o_quot, c) for c in range(n_quot)]), (zeta_next, trace_polys + aux_polys)]
        ys = [np.concatenate([tr0, ax0, qu0]), np.concatenate([tr1, ax1])]
        if n_zs:
            batches.append(([1, 0], [(o_aux, n_lookup + n_helpers + k) for k in range(n_zs)]))
            ys.append(np.array([[int(z), 0] for z in zs_first], dtype=np.uint64))
        fri = shard.fri_begin(extra, batches, ys, alpha, fp)
        mark("combine on the LDE domain (trace columns over NVLink)")
        st["fri"], st["extra"] = fri, extra
        st["fri_caps"], st["final_poly"] = fri.commit_phase(ch)
        mark("fold codewords in the commitment phase")
        st["pow_witness"] = ctx.fri_proof_of_work(ch, STARK_POW_BITS)
        mark("find proof-of-work witness")
        return [ch.get_challenge() % lde_n for _ in range(STARK_NUM_QUERIES)]

    idx = _collective(leader_until_queries, comm)[leader]

    # ---- Merkle paths of the queried trace rows, from their owners; whole rows on the leader (peers' columns over NVLink)
    def local_paths():
        mine = {}
        for i in sorted(set(idx)):
            if shard.first_row <= i < shard.first_row + shard.num_rows:
                mine[i] = shard.prove(i).reshape(-1).tolist()
        if rank == leader:
            st["rows"] = shard.leaves_at(idx)
        return mine

    paths = {}
    for part in _collective(local_paths, comm):  # also the barrier after which no rank reads a peer's LDE any more
        paths.update(part)
    if rank != leader:
        return None
    aux, quot, zs_first, fri, extra, rows = st["aux"], st["quot"], st["zs_first"], st["fri"], st["extra"], st["rows"]
    tr0, tr1, ax0, ax1, qu0 = st["openings"]
    fri_caps, final_poly, pow_witness = st["fri_caps"], st["final_poly"], st["pow_witness"]
    rest = fri.query_rounds(extra, idx)  # per query: auxiliary / quotient rows + paths, then the FRI layers
    mark("build FRI query rounds")

    # ---- the flat proof (wire.py / DESIGN.md "B200STK2")
    total = ctx.stark_proof_words(table, log_n)
    hdr = np.zeros(wire.HEADER_WORDS, dtype=np.uint64)
    vals = {"magic": wire.MAGIC, "table": table, "degree_bits": log_n, "n_trace": n_cols, "n_aux": n_aux, "n_quot": n_quot,
            "cap_height": STARK_CAP_HEIGHT, "n_fri_layers": fp.n_reductions, "arity_bits": 4, "final_poly_len": final_poly.shape[0],
            "num_queries": STARK_NUM_QUERIES, "n_public_inputs": n_pi, "rate_bits": STARK_RATE_BITS, "pow_bits": STARK_POW_BITS,
            "num_challenges": K, "total_words": total, "n_ctl_zs": n_zs, "n_lookup_cols": n_lookup, "n_ctl_helper_cols": n_helpers}
    for k, name in enumerate(wire.HEADER_FIELDS):
        hdr[k] = vals[name]
    out = [hdr, cap.reshape(-1)]
    if aux is not None:
        out.append(np.asarray(aux.cap, dtype=np.uint64).reshape(-1))
    out += [np.asarray(quot.cap, dtype=np.uint64).reshape(-1), tr0.reshape(-1), tr1.reshape(-1), ax0.reshape(-1), ax1.reshape(-1),
            np.asarray(zs_first, dtype=np.uint64), qu0.reshape(-1), np.asarray(fri_caps, dtype=np.uint64).reshape(-1)]
    for q, i in enumerate(idx):
        out += [rows[q], np.array(paths[i], dtype=np.uint64), rest[q]]
    out += [np.asarray(final_poly, dtype=np.uint64).reshape(-1), np.array([pow_witness], dtype=np.uint64), np.array(pi, dtype=np.uint64)]
    proof = np.concatenate([np.asarray(x, dtype=np.uint64).reshape(-1) for x in out])
    if proof.size != total:
        raise EtpError(-3, f"internal error: proof size mismatch ({proof.size} != {total})")
    return proof


# ---- several prover contexts on one GPU ----------------------------------------------------------------
class ProverPool:
    """`workers` independent contexts (own stream, own scratch) on one device, each driven by a host thread.

    The reference runs several Paladin workers per machine (/root/reference/README.md:100-106); on a GPU box the
    analogue is more than one prover context per GPU: the latency-bound tail of one proof (FRI tails, tree tops,
    proof-of-work, query gathers, transcript round trips) overlaps the commits of another.  Measured on B200
    (tools/prove_concurrent.py): 2 contexts give +8 % proofs/min at 2^22 rows and +45 % at 2^16; a third one
    gains nothing.  The library calls release the GIL (ctypes), so plain threads suffice."""

    def __init__(self, device: int, workers: int = 2, contexts: Sequence[Any] = None):
        """`contexts`: ready-made contexts to drive instead of creating `workers` new ones on `device`."""
        if contexts is not None:
            self.contexts = list(contexts)
            if not self.contexts:
                raise ValueError("contexts must not be empty")
            return
        import eth_tx_proof_b200 as etp

        if workers < 1:
            raise ValueError("workers must be >= 1")
        self.contexts = [etp.Context(device) for _ in range(workers)]

    def close(self):
        for c in self.contexts:
            c.close()
        self.contexts = []

    def map(self, fn, jobs: Sequence[Any]) -> List[Any]:
        """results[i] = fn(context, jobs[i]); job i runs on context i % workers, in submission order per context."""
        import threading

        results: List[Any] = [None] * len(jobs)
        errors: List[BaseException] = []

        def run(w):
            try:
                for i in range(w, len(jobs), len(self.contexts)):
                    results[i] = fn(self.contexts[w], jobs[i])
            except BaseException as e:  # surfaced on the calling thread
                errors.append(e)

        threads = [threading.Thread(target=run, args=(w,)) for w in range(len(self.contexts))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return results

    def stark_prove_dev(self, table: int, log_n: int, traces_dev: Sequence[Tuple[int, int]]) -> List[Any]:
        """One proof per (device pointer, column stride) trace, spread over the pool's contexts."""
        return self.map(lambda c, t: c.stark_prove_dev(table, log_n, t[0], t[1]), list(traces_dev))
