"""Synthetic workloads of the shapes BASELINE.json names (host side, numpy).

The reference obtains its traces from the EVM kernel interpreter inside ``evm_arithmetization``
(``generate_traces``; reached from /root/reference/ops/src/lib.rs:52), which needs a real witness from
an Ethereum RPC node (/root/reference/leader/src/lib.rs:158-535) and is out of scope.  These
generators stand in for it: counter-based (SplitMix64), so every rank / the CPU baseline can rebuild
the same data from ``(seed, shape)`` without transfers.

* ``random_columns``      — config 2 (commit microbench): uniform Goldilocks columns.
* ``fibonacci_trace``     — starky's FibonacciStark example table (2 columns, 3 public inputs).
* ``memory_trace``        — config 3: a constraint-satisfying trace for the memory-shaped table
                            (21 columns; SURVEY.md Appendix A sketch of evm_arithmetization's MemoryStark).
"""
from __future__ import annotations

import numpy as np

P = 0xFFFFFFFF00000001
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)

# memory table column indices (evm_arithmetization/src/memory/columns.rs, recalled order)
M_FILTER, M_TIMESTAMP, M_IS_READ, M_CTX, M_SEG, M_VIRT, M_VALUE0 = 0, 1, 2, 3, 4, 5, 6
M_CFC, M_SFC, M_VFC, M_INIT_AUX, M_RANGE_CHECK, M_COUNTER, M_FREQ = 14, 15, 16, 17, 18, 19, 20
MEMORY_COLUMNS = 21


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def _rand(seed: int, stream: int, n: int) -> np.ndarray:
    base = splitmix64(np.array([seed * 0x1000003 + stream], dtype=np.uint64))[0]
    with np.errstate(over="ignore"):
        return splitmix64(np.arange(n, dtype=np.uint64) * np.uint64(0x2545F4914F6CDD1D) + base)


def random_columns(n_cols: int, log_n: int, seed: int = 0xB200) -> np.ndarray:
    """(n_cols, 2^log_n) canonical Goldilocks elements, column c keyed by seed + c."""
    n = 1 << log_n
    out = np.empty((n_cols, n), dtype=np.uint64)
    for c in range(n_cols):
        x = _rand(seed + c, 0, n)
        out[c] = np.where(x >= np.uint64(P), x - np.uint64(P), x)
    return out


def fibonacci_trace(log_n: int, seed: int = 1):
    """Returns (trace (2, n), public_inputs [x0, x1, x1_last])."""
    n = 1 << log_n
    x0, x1 = (seed * 3 + 1) % P, (seed * 5 + 2) % P
    pi = [x0, x1]
    t = np.zeros((2, n), dtype=np.uint64)
    for i in range(n):
        t[0, i], t[1, i] = x0, x1
        x0, x1 = x1, (x0 + x1) % P
    pi.append(int(t[1, n - 1]))
    return t, pi


def _segmented_cumsum(incr: np.ndarray, start_mask: np.ndarray, start_val: np.ndarray) -> np.ndarray:
    """out[i] = start_val[s] + sum(incr[s+1..i]) where s = last index <= i with start_mask set (index 0 is a start)."""
    n = incr.size
    idx = np.where(start_mask, np.arange(n), 0)
    idx = np.maximum.accumulate(idx)
    inc = np.where(start_mask, 0, incr).astype(np.int64)
    cs = np.cumsum(inc)
    return start_val[idx].astype(np.int64) + cs - cs[idx]


def memory_trace(log_n: int, seed: int = 7, dummy_rows: int | None = None) -> np.ndarray:
    """A (21, 2^log_n) trace satisfying every constraint of the memory-shaped table, including the
    logUp range check of RANGE_CHECK against COUNTER with multiplicities FREQUENCIES."""
    n = 1 << log_n
    assert n >= 32
    if dummy_rows is None:
        dummy_rows = max(1, n // 16)
    real = n - dummy_rows
    r_kind = (_rand(seed, 1, n) & np.uint64(0xFF)).astype(np.int64)
    r_a = _rand(seed, 2, n)
    r_b = _rand(seed, 3, n)
    # kind[i] describes how row i differs from row i-1: 0 unchanged, 1 virt, 2 seg, 3 ctx change
    kind = np.where(r_kind < 160, 0, np.where(r_kind < 232, 1, np.where(r_kind < 250, 2, 3)))
    kind[0] = 3
    kind[real:] = 0  # dummy rows repeat the last address
    is_first = kind != 0
    d_ctx = np.where(kind == 3, 1 + (r_a % np.uint64(3)).astype(np.int64), 0)
    d_ctx[0] = 0
    ctx = np.cumsum(d_ctx)
    seg = _segmented_cumsum(np.where(kind == 2, 1 + ((r_a >> np.uint64(8)) % np.uint64(2)).astype(np.int64), 0),
                            kind == 3, ((r_b >> np.uint64(4)) % np.uint64(4)).astype(np.int64))
    virt = _segmented_cumsum(np.where(kind == 1, 1 + ((r_a >> np.uint64(16)) % np.uint64(8)).astype(np.int64), 0),
                             kind >= 2, ((r_b >> np.uint64(12)) % np.uint64(64)).astype(np.int64))
    d_ts = 1 + ((r_a >> np.uint64(24)) % np.uint64(4)).astype(np.int64)
    d_ts[real:] = 0
    ts = _segmented_cumsum(np.where(kind == 0, d_ts, 0), is_first, ((r_b >> np.uint64(20)) % np.uint64(16)).astype(np.int64))
    # reads / writes
    coin = (r_b >> np.uint64(32)) & np.uint64(3)
    is_read = np.where(is_first, coin == 0, coin < 2)
    is_read[real:] = True
    # value definition points: writes, and first-op reads (value 0)
    is_def = is_first | ~is_read
    def_idx = np.maximum.accumulate(np.where(is_def, np.arange(n), 0))
    t = np.zeros((MEMORY_COLUMNS, n), dtype=np.uint64)
    for limb in range(8):
        v = _rand(seed, 10 + limb, n) & np.uint64(0xFFFFFFFF)
        v = np.where(is_first & is_read, np.uint64(0), v)
        t[M_VALUE0 + limb] = v[def_idx]
    t[M_FILTER, :real] = 1
    t[M_TIMESTAMP] = ts.astype(np.uint64)
    t[M_IS_READ] = is_read.astype(np.uint64)
    t[M_CTX] = ctx.astype(np.uint64)
    t[M_SEG] = seg.astype(np.uint64)
    t[M_VIRT] = virt.astype(np.uint64)
    # flags on row i describe the step i -> i+1; last row: all zero
    nk = np.concatenate([kind[1:], [0]])
    t[M_CFC] = (nk == 3).astype(np.uint64)
    t[M_SFC] = (nk == 2).astype(np.uint64)
    t[M_VFC] = (nk == 1).astype(np.uint64)
    nxt = lambda a: np.concatenate([a[1:], a[-1:]])
    rc = np.where(nk == 3, nxt(ctx) - ctx - 1, np.where(nk == 2, nxt(seg) - seg - 1,
                  np.where(nk == 1, nxt(virt) - virt - 1, nxt(ts) - ts)))
    rc[-1] = 0
    assert rc.min() >= 0 and rc.max() < n
    t[M_RANGE_CHECK] = rc.astype(np.uint64)
    init_aux = nxt(seg) * (nk != 0) * nxt(is_read.astype(np.int64))
    init_aux[-1] = 0
    t[M_INIT_AUX] = init_aux.astype(np.uint64)
    t[M_COUNTER] = np.arange(n, dtype=np.uint64)
    t[M_FREQ] = np.bincount(rc, minlength=n).astype(np.uint64)
    return t


def tx_job_tables(scale_bits: int = 0):
    """The seven table proofs of one synthetic transaction: [(name, program or None, degree_bits, trace)] in the order the
    reference's AllStark lists its tables.  `program` is a cprog.Program to register (shape-only stand-ins of the
    evm_arithmetization tables, cprog.EVM_TABLE_SHAPES, and the 523-column logic table) or None for the built-in memory
    table; degree bits = cprog.TX_TABLE_DEGREE_BITS + scale_bits.  Shape only: no cross-table lookups, no recursion."""
    from . import cprog

    out = []
    for name in ("arithmetic", "byte_packing", "cpu", "keccak", "keccak_sponge", "logic", "memory"):
        bits = max(5, cprog.TX_TABLE_DEGREE_BITS[name] + scale_bits)
        if name == "memory":
            out.append((name, None, bits, memory_trace(bits, seed=31)))
        elif name == "logic":
            out.append((name, cprog.logic_program(8), bits, cprog.logic_trace(bits, 8, seed=37)))
        else:
            cols, lk = cprog.EVM_TABLE_SHAPES[name]
            out.append((name, cprog.shape_program(cols, lk), bits, cprog.shape_trace(bits, cols, lk, seed=41)))
    return out


def shape_trace_columns_dev(log_n: int, n_cols: int, n_lookup: int, c0: int, c1: int, seed: int = 77):
    """Columns [c0, c1) of a trace that satisfies cprog.shape_program(n_cols, n_lookup), generated on the current CUDA device
    (torch int64 tensor (c1 - c0, n)): every column is a function of (seed, column) only, so each rank of a column-split
    table builds just the columns it owns and any rank can rebuild any column.  Not the numpy trace of cprog.shape_trace
    (different random stream): a valid trace of the same shape for sizes no host generates in reasonable time."""
    import torch

    from . import cprog

    lay = cprog.shape_layout(n_cols, n_lookup)
    n = 1 << log_n

    def rnd(stream, hi, shape):
        gen = torch.Generator(device="cuda").manual_seed(seed * 100003 + stream)
        return torch.randint(0, hi, shape, dtype=torch.int64, device="cuda", generator=gen)

    out = torch.empty((max(c1 - c0, 0), n), dtype=torch.int64, device="cuda")
    group = (None, None)
    for c in range(c0, c1):
        if c == 0:
            out[c - c0] = torch.arange(n, dtype=torch.int64, device="cuda")
        elif n_lookup and c == lay["FREQ"]:  # multiplicities of the counter values among all limbs
            freq = torch.zeros(n, dtype=torch.int64, device="cuda")
            for j in range(n_lookup):
                freq += torch.bincount(rnd(50000 + j, n, (n,)), minlength=n)
            out[c - c0] = freq
        elif n_lookup and c < lay["GROUP"]:
            out[c - c0] = rnd(50000 + c - lay["LIMB"], n, (n,))
        elif c < lay["FLAG"]:
            g, k = divmod(c - lay["GROUP"], 4)
            if group[0] != g:
                group = (g, rnd(g, 1 << 20, (3, n)))
            a, b, d = group[1]
            out[c - c0] = (a, b, d, a * b * d if g % 5 == 4 else a * b + d)[k]
        else:
            out[c - c0] = rnd(90000 + c, 2, (n,))
    return out
