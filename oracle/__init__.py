"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes loader for ``oracle/liboracle.so``, the CPU restatement (C11 + OpenMP) of the plonky2/starky
proving hot path reached from the reference at ``/root/reference/ops/src/lib.rs:52``.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import this module, and only as the checker or the timed CPU baseline.  Nothing under
``eth_tx_proof_b200/`` imports it.

Parity status: Poseidon constants + permutation pinned by upstream known-answer vectors
(``tests/golden/poseidon_kat.json``); everything composed above is "parity unpinned" against real
plonky2 output (no Rust toolchain, no upstream sources; see ``oracle/oracle.h``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

P = 0xFFFFFFFF00000001
TABLE_FIBONACCI = 0
TABLE_MEMORY = 1


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc + OpenMP). Building the checker is not using it."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h")) or f == "Makefile"]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    if force or stale:
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, env=env,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_u64p = C.POINTER(C.c_uint64)


def _ptr(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u64p)


class Challenger(C.Structure):
    _fields_ = [("state", C.c_uint64 * 12), ("inb", C.c_uint64 * 8), ("n_in", C.c_int),
                ("out", C.c_uint64 * 8), ("n_out", C.c_int)]


class FriParams(C.Structure):
    _fields_ = [("rate_bits", C.c_int), ("cap_height", C.c_int), ("proof_of_work_bits", C.c_int), ("num_query_rounds", C.c_int),
                ("degree_bits", C.c_int), ("n_reductions", C.c_int), ("reduction_arity_bits", C.c_int * 16)]


class FriPoly(C.Structure):
    _fields_ = [("oracle_index", C.c_uint32), ("polynomial_index", C.c_uint32)]


class FriBatch(C.Structure):
    _fields_ = [("point", C.c_uint64 * 2), ("polynomials", C.POINTER(FriPoly)), ("n_polynomials", C.c_size_t)]


class _Batch(C.Structure):
    _fields_ = [("n_cols", C.c_size_t), ("log_n", C.c_int), ("rate_bits", C.c_int), ("cap_height", C.c_int),
                ("coeffs", _u64p), ("leaves", _u64p), ("digests", _u64p), ("cap", _u64p)]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    try:
        build()
        L = C.CDLL(_LIB_PATH)
    except (OSError, subprocess.CalledProcessError):
        build(force=True)
        L = C.CDLL(_LIB_PATH)
    sz = C.c_size_t
    L.orc_poseidon_constants.argtypes = [_u64p]
    L.orc_poseidon_permute.argtypes = [_u64p]
    L.orc_poseidon_permute_naive.argtypes = [_u64p]
    L.orc_hash_no_pad.argtypes = [_u64p, sz, _u64p]
    L.orc_hash_or_noop.argtypes = [_u64p, sz, _u64p]
    L.orc_two_to_one.argtypes = [_u64p, _u64p, _u64p]
    L.orc_challenger_init.argtypes = [C.POINTER(Challenger)]
    L.orc_challenger_observe.argtypes = [C.POINTER(Challenger), _u64p, sz]
    L.orc_challenger_get.argtypes = [C.POINTER(Challenger)]
    L.orc_challenger_get.restype = C.c_uint64
    for f in (L.orc_fft, L.orc_ifft):
        f.argtypes = [_u64p, C.c_int]
    for f in (L.orc_coset_fft, L.orc_coset_ifft):
        f.argtypes = [_u64p, C.c_int, C.c_uint64]
    L.orc_lde.argtypes = [_u64p, C.c_int, C.c_int, _u64p]
    L.orc_merkle_new.argtypes = [_u64p, sz, sz, C.c_int, _u64p, _u64p]
    L.orc_merkle_prove.argtypes = [_u64p, sz, C.c_int, sz, _u64p]
    L.orc_merkle_verify.argtypes = [_u64p, sz, sz, _u64p, C.c_int, _u64p, C.c_int]
    L.orc_merkle_verify.restype = C.c_int
    for f in (L.orc_batch_from_values, L.orc_batch_from_coeffs):
        f.argtypes = [_u64p, sz, C.c_int, C.c_int, C.c_int]
        f.restype = C.POINTER(_Batch)
    L.orc_batch_free.argtypes = [C.POINTER(_Batch)]
    L.orc_batch_num_digests.argtypes = [C.POINTER(_Batch)]
    L.orc_batch_num_digests.restype = sz
    for f in (L.orc_table_num_columns, L.orc_table_constraint_degree, L.orc_table_num_public_inputs,
              L.orc_table_uses_lookup):
        f.argtypes = [C.c_int]
        f.restype = C.c_int
    L.orc_table_register.argtypes = [_u64p, sz, C.POINTER(C.c_int32), sz]
    L.orc_table_register.restype = C.c_int
    L.orc_table_register_ex.argtypes = [_u64p, sz, _u64p, sz]
    L.orc_table_register_ex.restype = C.c_int
    for f in (L.orc_table_requires_ctls, L.orc_table_num_ctl_helper_columns, L.orc_table_num_ctl_zs):
        f.argtypes = [C.c_int]
        f.restype = C.c_int
    L.orc_table_num_lookup_columns.argtypes = [C.c_int, C.c_int]
    L.orc_table_num_lookup_columns.restype = C.c_int
    L.orc_fri_params_make.argtypes = [C.c_int] * 5 + [C.POINTER(FriParams)]
    L.orc_fri_params_standard_fast.argtypes = [C.c_int, C.POINTER(FriParams)]
    L.orc_prove_with_commitment.argtypes = [C.c_int, C.c_int, _u64p, C.POINTER(_Batch), _u64p, C.POINTER(Challenger), _u64p, _u64p]
    L.orc_prove_with_commitment.restype = C.c_int
    L.orc_challenger_compact.argtypes = [C.POINTER(Challenger), _u64p]
    L.orc_aux_columns.argtypes = [C.c_int, C.c_int, _u64p, _u64p, C.c_int, _u64p, _u64p]
    L.orc_aux_columns.restype = C.c_int
    L.orc_batch_eval_at_ext_point.argtypes = [C.POINTER(_Batch), _u64p, _u64p]
    L.orc_fri_proof_words.argtypes = [C.POINTER(sz), sz, C.POINTER(FriParams)]
    L.orc_fri_proof_words.restype = sz
    L.orc_prove_openings.argtypes = [C.POINTER(FriBatch), sz, C.POINTER(C.POINTER(_Batch)), sz, C.POINTER(Challenger),
                                     C.POINTER(FriParams), _u64p]
    L.orc_prove_openings.restype = C.c_int
    L.orc_challenger_get_n.argtypes = [C.POINTER(Challenger), sz, _u64p]
    L.orc_table_num_aux_columns.argtypes = [C.c_int, C.c_int]
    L.orc_table_num_aux_columns.restype = C.c_int
    L.orc_table_check_constraints.argtypes = [C.c_int, C.c_int, _u64p, _u64p]
    L.orc_table_check_constraints.restype = C.c_long
    L.orc_stark_proof_words.argtypes = [C.c_int, C.c_int]
    L.orc_stark_proof_words.restype = sz
    L.orc_stark_prove.argtypes = [C.c_int, C.c_int, _u64p, _u64p, _u64p]
    L.orc_stark_prove.restype = C.c_int
    L.orc_lookup_helper_columns.argtypes = [C.c_int, C.c_int, _u64p, _u64p, C.c_int, _u64p]
    L.orc_compute_quotient_polys.argtypes = [C.c_int, C.c_int, C.POINTER(_Batch), C.POINTER(_Batch), _u64p,
                                             _u64p, _u64p, C.c_int, _u64p]
    L.orc_pow_grind.argtypes = [_u64p, C.c_int, C.c_int]
    L.orc_pow_grind.restype = C.c_uint64
    L.orc_fri_fold_coeffs.argtypes = [_u64p, sz, C.c_int, _u64p, _u64p]
    L.orc_plonk_partial_products_and_zs.argtypes = [_u64p, _u64p, _u64p, C.c_int, C.c_int, C.c_int, _u64p, _u64p, C.c_int, _u64p]
    L.orc_plonk_partial_products_and_zs.restype = C.c_int
    L.orc_num_threads.restype = C.c_int
    _lib = L
    return L


def _u64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint64))


# ---------------------------------------------------------------- thin numpy wrappers
def poseidon_constants() -> np.ndarray:
    out = np.zeros(360, dtype=np.uint64)
    lib().orc_poseidon_constants(_ptr(out))
    return out


def poseidon_permute(state) -> np.ndarray:
    s = _u64(state).copy()
    assert s.shape == (12,)
    lib().orc_poseidon_permute(_ptr(s))
    return s


def poseidon_permute_naive(state) -> np.ndarray:
    """The round-by-round definition (orc_poseidon_permute is the fast form of the same function)."""
    s = _u64(state).copy()
    assert s.shape == (12,)
    lib().orc_poseidon_permute_naive(_ptr(s))
    return s


def hash_no_pad(x) -> np.ndarray:
    x = _u64(x).ravel()
    out = np.zeros(4, dtype=np.uint64)
    lib().orc_hash_no_pad(_ptr(x) if x.size else None, x.size, _ptr(out))
    return out


def hash_or_noop(x) -> np.ndarray:
    x = _u64(x).ravel()
    out = np.zeros(4, dtype=np.uint64)
    lib().orc_hash_or_noop(_ptr(x) if x.size else None, x.size, _ptr(out))
    return out


def two_to_one(l, r) -> np.ndarray:
    out = np.zeros(4, dtype=np.uint64)
    lib().orc_two_to_one(_ptr(_u64(l)), _ptr(_u64(r)), _ptr(out))
    return out


def fft(a) -> np.ndarray:
    a = _u64(a).copy()
    lib().orc_fft(_ptr(a), int(a.size).bit_length() - 1)
    return a


def ifft(a) -> np.ndarray:
    a = _u64(a).copy()
    lib().orc_ifft(_ptr(a), int(a.size).bit_length() - 1)
    return a


def coset_fft(a, shift=7) -> np.ndarray:
    a = _u64(a).copy()
    lib().orc_coset_fft(_ptr(a), int(a.size).bit_length() - 1, shift)
    return a


def coset_ifft(a, shift=7) -> np.ndarray:
    a = _u64(a).copy()
    lib().orc_coset_ifft(_ptr(a), int(a.size).bit_length() - 1, shift)
    return a


def lde(coeffs, rate_bits) -> np.ndarray:
    c = _u64(coeffs)
    out = np.zeros(c.size << rate_bits, dtype=np.uint64)
    lib().orc_lde(_ptr(c), int(c.size).bit_length() - 1, rate_bits, _ptr(out))
    return out


def merkle_new(leaves, cap_height):
    """leaves: (n_leaves, leaf_len) -> (digests (nd,4) in plonky2 layout, cap (2^h,4))"""
    lv = _u64(leaves)
    n, ll = lv.shape
    nd = 2 * (n - (1 << cap_height))
    digests = np.zeros((max(nd, 1), 4), dtype=np.uint64)
    cap = np.zeros((1 << cap_height, 4), dtype=np.uint64)
    lib().orc_merkle_new(_ptr(lv), n, ll, cap_height, _ptr(digests), _ptr(cap))
    return digests[:nd], cap


def merkle_prove(digests, n_leaves, cap_height, leaf_index) -> np.ndarray:
    d = _u64(digests)
    ns = (int(n_leaves).bit_length() - 1) - cap_height
    sib = np.zeros((max(ns, 1), 4), dtype=np.uint64)
    lib().orc_merkle_prove(_ptr(d) if d.size else None, n_leaves, cap_height, leaf_index, _ptr(sib))
    return sib[:ns]


def merkle_verify(leaf, leaf_index, siblings, cap) -> bool:
    leaf = _u64(leaf).ravel()
    sib = _u64(siblings).reshape(-1, 4)
    cap = _u64(cap).reshape(-1, 4)
    return bool(lib().orc_merkle_verify(_ptr(leaf) if leaf.size else None, leaf.size, leaf_index,
                                        _ptr(sib) if sib.size else None, sib.shape[0], _ptr(cap),
                                        int(cap.shape[0]).bit_length() - 1))


class Batch:
    """PolynomialBatch on the CPU (oracle)."""

    def __init__(self, handle):
        self._h = handle
        b = handle.contents
        self.n_cols, self.log_n, self.rate_bits, self.cap_height = b.n_cols, b.log_n, b.rate_bits, b.cap_height

    @classmethod
    def from_values(cls, values, rate_bits, cap_height):
        v = _u64(values)
        c, n = v.shape
        return cls(lib().orc_batch_from_values(_ptr(v), c, int(n).bit_length() - 1, rate_bits, cap_height))

    @classmethod
    def from_coeffs(cls, coeffs, rate_bits, cap_height):
        v = _u64(coeffs)
        c, n = v.shape
        return cls(lib().orc_batch_from_coeffs(_ptr(v), c, int(n).bit_length() - 1, rate_bits, cap_height))

    def _arr(self, p, shape):
        n = int(np.prod(shape))
        if n == 0:
            return np.zeros(shape, dtype=np.uint64)
        return np.ctypeslib.as_array(p, shape=(n,)).reshape(shape).copy()

    @property
    def coeffs(self):
        return self._arr(self._h.contents.coeffs, (self.n_cols, 1 << self.log_n))

    @property
    def leaves(self):
        return self._arr(self._h.contents.leaves, (1 << (self.log_n + self.rate_bits), self.n_cols))

    @property
    def digests(self):
        return self._arr(self._h.contents.digests, (lib().orc_batch_num_digests(self._h), 4))

    @property
    def cap(self):
        return self._arr(self._h.contents.cap, (1 << self.cap_height, 4))

    def __del__(self):
        if getattr(self, "_h", None) is not None and _lib is not None:
            _lib.orc_batch_free(self._h)
            self._h = None


def register_table(program, lookups=()) -> int:
    """Program-defined table (words of eth_tx_proof_b200.cprog.Program or a u64 array; lookups as
    [(looking_columns, table_column, frequencies_column)]): interpreted by the oracle.  Returns its table id."""
    words = _u64(getattr(program, "words", program))
    flat = [len(lookups)]
    for looking, table_col, freq_col in lookups:
        flat += [int(table_col), int(freq_col), len(looking)] + [int(c) for c in looking]
    arr = (C.c_int32 * len(flat))(*flat)
    tid = lib().orc_table_register(_ptr(words), words.size, arr, len(flat) if lookups else 0)
    if tid < 0:
        raise RuntimeError("oracle: table registration failed")
    return tid


def register_table_ex(program, aux_spec) -> int:
    """General registration: `aux_spec` = the u64 words of eth_tx_proof_b200.cprog.AuxSpec (lookups with linear-combination
    Columns and Filters + the table's CTL Z descriptors)."""
    words = _u64(getattr(program, "words", program))
    spec = _u64(getattr(aux_spec, "words", aux_spec))
    tid = lib().orc_table_register_ex(_ptr(words), words.size, _ptr(spec), spec.size)
    if tid < 0:
        raise RuntimeError("oracle: table registration failed")
    return tid


class HostChallenger:
    """plonky2::iop::challenger::Challenger<GoldilocksField, PoseidonHash> (the oracle's restatement)."""

    def __init__(self):
        self.c = Challenger()
        lib().orc_challenger_init(C.byref(self.c))

    def observe(self, elems):
        e = _u64(elems).ravel()
        if e.size:
            lib().orc_challenger_observe(C.byref(self.c), _ptr(e), e.size)

    def get_n(self, n) -> np.ndarray:
        out = np.zeros(n, dtype=np.uint64)
        lib().orc_challenger_get_n(C.byref(self.c), n, _ptr(out))
        return out

    def compact(self) -> np.ndarray:
        out = np.zeros(12, dtype=np.uint64)
        lib().orc_challenger_compact(C.byref(self.c), _ptr(out))
        return out

    def words(self) -> np.ndarray:
        """[sponge_state 12 | input_buffer 8 | input_len | output_buffer 8 | output_len] — the product's etp_challenger layout."""
        return np.array(list(self.c.state) + list(self.c.inb) + [self.c.n_in] + list(self.c.out) + [self.c.n_out], dtype=np.uint64)


def aux_columns(table, trace, lookup_challenges, ctl_challenges=None) -> np.ndarray:
    """All auxiliary polynomials of a table (lookup columns ++ CTL helper columns ++ CTL Zs), values on the trace domain."""
    t = _u64(trace)
    lc = _u64(lookup_challenges)
    cc = _u64(list(ctl_challenges if ctl_challenges is not None else []) + [0, 0, 0, 0])
    na = lib().orc_table_num_aux_columns(table, lc.size)
    aux = np.zeros((max(na, 1), t.shape[1]), dtype=np.uint64)
    rc = lib().orc_aux_columns(table, int(t.shape[1]).bit_length() - 1, _ptr(t), _ptr(lc), lc.size, _ptr(cc), _ptr(aux))
    if rc != 0:
        raise RuntimeError("oracle: zero denominator in an auxiliary column")
    return aux[:na]


def prove_with_commitment(table, trace, trace_batch: "Batch", challenger: HostChallenger, ctl_challenges=None, public_inputs=()):
    """starky::prover::prove_with_commitment; the challenger is updated in place."""
    t = _u64(trace)
    log_n = int(t.shape[1]).bit_length() - 1
    pi = _u64(list(public_inputs) + [0])
    out = np.zeros(lib().orc_stark_proof_words(table, log_n), dtype=np.uint64)
    cc = _u64(ctl_challenges) if ctl_challenges is not None else None
    rc = lib().orc_prove_with_commitment(table, log_n, _ptr(t), trace_batch._h, _ptr(cc) if cc is not None else None,
                                         C.byref(challenger.c), _ptr(pi), _ptr(out))
    if rc != 0:
        raise RuntimeError(f"oracle prove_with_commitment failed rc={rc}")
    return out


def fri_params(degree_bits, rate_bits=1, cap_height=4, pow_bits=16, num_queries=84) -> FriParams:
    p = FriParams()
    lib().orc_fri_params_make(degree_bits, rate_bits, cap_height, pow_bits, num_queries, C.byref(p))
    return p


def batch_eval_at_ext_point(batch: "Batch", z) -> np.ndarray:
    out = np.zeros((max(batch.n_cols, 1), 2), dtype=np.uint64)
    lib().orc_batch_eval_at_ext_point(batch._h, _ptr(_u64(z)), _ptr(out))
    return out[:batch.n_cols]


def prove_openings(batches, oracles, challenger: HostChallenger, params: FriParams) -> np.ndarray:
    """PolynomialBatch::prove_openings for a general FriInstanceInfo.  batches: [(point (2,), [(oracle_index, poly_index)])]."""
    keep = []
    arr = (FriBatch * len(batches))()
    for i, (point, polys) in enumerate(batches):
        pa = (FriPoly * max(len(polys), 1))()
        for k, (o, c) in enumerate(polys):
            pa[k].oracle_index, pa[k].polynomial_index = o, c
        keep.append(pa)
        arr[i].point[0], arr[i].point[1] = int(point[0]), int(point[1])
        arr[i].polynomials = C.cast(pa, C.POINTER(FriPoly))
        arr[i].n_polynomials = len(polys)
    oc = (C.c_size_t * len(oracles))(*[o.n_cols for o in oracles])
    out = np.zeros(lib().orc_fri_proof_words(oc, len(oracles), C.byref(params)), dtype=np.uint64)
    hs = (C.POINTER(_Batch) * len(oracles))(*[o._h for o in oracles])
    rc = lib().orc_prove_openings(arr, len(batches), hs, len(oracles), C.byref(challenger.c), C.byref(params), _ptr(out))
    if rc != 0:
        raise RuntimeError(f"oracle prove_openings failed rc={rc}")
    return out


def check_constraints(table, trace, public_inputs=()) -> int:
    t = _u64(trace)
    pi = _u64(list(public_inputs) + [0])
    return int(lib().orc_table_check_constraints(table, int(t.shape[1]).bit_length() - 1, _ptr(t), _ptr(pi)))


def stark_prove(table, trace, public_inputs=()) -> np.ndarray:
    t = _u64(trace)
    log_n = int(t.shape[1]).bit_length() - 1
    pi = _u64(list(public_inputs) + [0])
    out = np.zeros(lib().orc_stark_proof_words(table, log_n), dtype=np.uint64)
    rc = lib().orc_stark_prove(table, log_n, _ptr(t), _ptr(pi), _ptr(out))
    if rc != 0:
        raise RuntimeError(f"oracle stark_prove failed rc={rc}")
    return out


def lookup_helper_columns(table, trace, challenges) -> np.ndarray:
    t = _u64(trace)
    ch = _u64(challenges)
    na = lib().orc_table_num_aux_columns(table, ch.size)
    aux = np.zeros((na, t.shape[1]), dtype=np.uint64)
    lib().orc_lookup_helper_columns(table, int(t.shape[1]).bit_length() - 1, _ptr(t), _ptr(ch), ch.size, _ptr(aux))
    return aux


def compute_quotient_polys(table, trace_batch: Batch, aux_batch, lookup_challenges, public_inputs, alphas):
    L = lib()
    deg = L.orc_table_constraint_degree(table)
    factor = max(1, deg - 1)
    al = _u64(alphas)
    out = np.zeros((factor * al.size, 1 << trace_batch.log_n), dtype=np.uint64)
    lc = _u64(list(lookup_challenges) + [0] * 8)
    pi = _u64(list(public_inputs) + [0])
    L.orc_compute_quotient_polys(table, trace_batch.log_n, trace_batch._h, aux_batch._h if aux_batch else None,
                                 _ptr(lc), _ptr(pi), _ptr(al), al.size, _ptr(out))
    return out


def pow_grind(state, pos, bits) -> int:
    return int(lib().orc_pow_grind(_ptr(_u64(state)), pos, bits))


def fri_fold_coeffs(coeffs, arity_bits, beta) -> np.ndarray:
    c = _u64(coeffs).reshape(-1, 2)
    out = np.zeros((c.shape[0] >> arity_bits, 2), dtype=np.uint64)
    lib().orc_fri_fold_coeffs(_ptr(c), c.shape[0], arity_bits, _ptr(_u64(beta)), _ptr(out))
    return out


def plonk_partial_products_and_zs(wires, sigmas, k_is, quotient_degree_factor, betas, gammas) -> np.ndarray:
    """plonky2 all_wires_permutation_partial_products: [Z per challenge] ++ [partial products per challenge], values on the subgroup."""
    w, s_ = _u64(wires), _u64(sigmas)
    k, b, g = _u64(k_is), _u64(betas), _u64(gammas)
    n_routed, n = w.shape
    n_pp = -(-n_routed // quotient_degree_factor) - 1
    out = np.zeros((b.size * (1 + n_pp), n), dtype=np.uint64)
    rc = lib().orc_plonk_partial_products_and_zs(_ptr(w), _ptr(s_), _ptr(k), n_routed, int(n).bit_length() - 1, quotient_degree_factor,
                                                 _ptr(b), _ptr(g), b.size, _ptr(out))
    if rc != 0:
        raise RuntimeError("oracle: zero denominator in the permutation argument")
    return out


def num_threads() -> int:
    return int(lib().orc_num_threads())


def circuit_prove(circuit, wires, public_inputs, circuit_digest, table_id=None, timings=None) -> dict:
    """plonky2::plonk::prover::prove over the oracle's CPU primitives (test infrastructure; the checker of
    eth_tx_proof_b200/circuit.py CircuitProver.prove): wires commitment, all_wires_permutation_partial_products, the quotient
    of the circuit's vanishing program (compute_quotient_polys over ONE batch that holds every virtual column
    [constants | sigmas | wires | Zs | partial products | X] at rate_bits 3), openings, prove_openings over the four oracles.
    `circuit`: eth_tx_proof_b200.circuit.Circuit.  Same dict layout as the product's."""
    import time

    from eth_tx_proof_b200 import circuit as cc

    t = {} if timings is None else timings
    n, db = circuit.n, circuit.degree_bits
    t0 = time.perf_counter()
    cs_b = Batch.from_values(np.concatenate([circuit.constants, circuit.sigmas]), cc.RATE_BITS, cc.CAP_HEIGHT)
    t["constants_sigmas commit (per circuit)"] = (time.perf_counter() - t0) * 1e3
    pi_hash = [int(x) for x in hash_no_pad(np.array(public_inputs, dtype=np.uint64))]
    ch = HostChallenger()
    ch.observe(circuit_digest)
    ch.observe(pi_hash)
    t0 = time.perf_counter()
    wires_b = Batch.from_values(wires, cc.RATE_BITS, cc.CAP_HEIGHT)
    t["wires commit"] = (time.perf_counter() - t0) * 1e3
    ch.observe(wires_b.cap)
    betas, gammas = ch.get_n(cc.NUM_CHALLENGES), ch.get_n(cc.NUM_CHALLENGES)
    t0 = time.perf_counter()
    zs_pp = plonk_partial_products_and_zs(wires[:cc.NUM_ROUTED], circuit.sigmas, circuit.k_is, cc.QUOTIENT_DEGREE_FACTOR, betas, gammas)
    zs_b = Batch.from_values(zs_pp, cc.RATE_BITS, cc.CAP_HEIGHT)
    t["partial products and Zs + commit"] = (time.perf_counter() - t0) * 1e3
    ch.observe(zs_b.cap)
    alphas = ch.get_n(cc.NUM_CHALLENGES)
    t0 = time.perf_counter()
    if table_id is None:
        table_id = register_table(circuit.program)
    virt = Batch.from_values(circuit.virtual_trace(wires, zs_pp), cc.RATE_BITS, 0)
    quot = compute_quotient_polys(table_id, virt, None, list(betas) + list(gammas), pi_hash, alphas)
    del virt
    t["compute quotient polys"] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    quot_b = Batch.from_coeffs(quot, cc.RATE_BITS, cc.CAP_HEIGHT)
    t["quotient commit"] = (time.perf_counter() - t0) * 1e3
    ch.observe(quot_b.cap)
    zeta = ch.get_n(2)
    g = pow(1753635133440165772, 1 << (32 - db), 0xFFFFFFFF00000001)
    zeta_next = [int(zeta[0]) * g % 0xFFFFFFFF00000001, int(zeta[1]) * g % 0xFFFFFFFF00000001]
    oracles = [cs_b, wires_b, zs_b, quot_b]
    t0 = time.perf_counter()
    openings = [batch_eval_at_ext_point(o, zeta) for o in oracles]
    zs_next = batch_eval_at_ext_point(zs_b, zeta_next)[:cc.NUM_CHALLENGES]
    t["openings"] = (time.perf_counter() - t0) * 1e3
    for o in openings:
        ch.observe(o)
    ch.observe(zs_next)
    shapes = [o.n_cols for o in oracles]
    all_polys = [(o, k) for o, cnt in enumerate(shapes) for k in range(cnt)]
    batches = [([int(zeta[0]), int(zeta[1])], all_polys), (zeta_next, [(2, k) for k in range(cc.NUM_CHALLENGES)])]
    t0 = time.perf_counter()
    fri = prove_openings(batches, oracles, ch, fri_params(db, cc.RATE_BITS, cc.CAP_HEIGHT, cc.POW_BITS, cc.NUM_QUERIES))
    t["prove_openings (FRI)"] = (time.perf_counter() - t0) * 1e3
    return {"degree_bits": db, "public_inputs": [int(x) for x in public_inputs], "constants_sigmas_cap": cs_b.cap,
            "wires_cap": wires_b.cap, "plonk_zs_partial_products_cap": zs_b.cap, "quotient_polys_cap": quot_b.cap,
            "openings": {"constants_sigmas": openings[0], "wires": openings[1], "zs_partial_products": openings[2], "quotient_polys": openings[3],
                         "plonk_zs_next": zs_next},
            "opening_proof": fri, "quotient_coeffs": quot, "ms": t}
