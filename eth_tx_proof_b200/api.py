"""ctypes binding of ``libetp_b200.so`` shaped like plonky2's operator surface.

Mirrors (names, argument meaning, panics -> ``EtpError``):
  plonky2::fri::oracle::PolynomialBatch::{from_values, from_coeffs, get_lde_values}, .polynomials,
      .merkle_tree.{leaves, digests, cap, prove}
  plonky2::hash::merkle_tree::MerkleTree::{new, prove}
  starky::prover::{prove, compute_quotient_polys}, starky::lookup::lookup_helper_columns
(plonky2 0.2.2 / starky 0.4.0, /root/reference/Cargo.lock:3441,4529; reached from the reference's
worker at /root/reference/ops/src/lib.rs:52).  numpy arrays are host memory; ``torch`` CUDA tensors
(int64 storage viewed as u64) can be passed to the ``*_dev`` variants by their ``data_ptr()``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

TABLE_FIBONACCI = 0
TABLE_MEMORY = 1

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_u64p = C.POINTER(C.c_uint64)


class EtpError(RuntimeError):
    """A failed library call (upstream: panic / FatalError at /root/reference/ops/src/lib.rs:52)."""

    def __init__(self, code, msg):
        super().__init__(f"etp_b200 error {code}: {msg}")
        self.code = code


class Challenger(C.Structure):
    """etp_challenger: plonky2::iop::challenger::Challenger<GoldilocksField, PoseidonHash> as plain data — the transcript
    state every proving entry point takes in and hands back."""
    _fields_ = [("sponge_state", C.c_uint64 * 12), ("input_buffer", C.c_uint64 * 8), ("output_buffer", C.c_uint64 * 8),
                ("input_len", C.c_uint32), ("output_len", C.c_uint32)]

    def observe(self, elements):
        e = _u64(elements).ravel()
        if e.size:
            load_library().etp_challenger_observe(C.byref(self), _p(e), e.size)

    def observe_cap(self, cap):
        self.observe(cap)

    def get_challenge(self) -> int:
        return int(load_library().etp_challenger_get_challenge(C.byref(self)))

    def get_n_challenges(self, n) -> np.ndarray:
        out = np.zeros(n, dtype=np.uint64)
        load_library().etp_challenger_get_n_challenges(C.byref(self), n, _p(out))
        return out

    def get_extension_challenge(self) -> np.ndarray:
        return self.get_n_challenges(2)

    def compact(self) -> np.ndarray:
        load_library().etp_challenger_compact(C.byref(self))
        return np.array(list(self.sponge_state), dtype=np.uint64)

    def words(self) -> np.ndarray:
        """[sponge_state 12 | input_buffer 8 | input_len | output_buffer 8 | output_len]"""
        return np.array(list(self.sponge_state) + list(self.input_buffer) + [self.input_len] + list(self.output_buffer) + [self.output_len],
                        dtype=np.uint64)

    def clone(self) -> "Challenger":
        c = Challenger()
        C.memmove(C.byref(c), C.byref(self), C.sizeof(Challenger))
        return c


class FriParams(C.Structure):
    """etp_fri_params: plonky2::fri::FriParams (FriConfig + degree_bits + reduction_arity_bits)."""
    _fields_ = [("rate_bits", C.c_int), ("cap_height", C.c_int), ("proof_of_work_bits", C.c_int), ("num_query_rounds", C.c_int),
                ("degree_bits", C.c_int), ("n_reductions", C.c_int), ("reduction_arity_bits", C.c_int * 16)]

    @classmethod
    def make(cls, degree_bits, rate_bits=1, cap_height=4, proof_of_work_bits=16, num_query_rounds=84) -> "FriParams":
        """FriConfig::fri_params with ConstantArityBits(4, 5); the defaults are StarkConfig::standard_fast_config(),
        (3, 4, 16, 28) is CircuitConfig::standard_recursion_config()."""
        p = cls()
        rc = load_library().etp_fri_params_make(degree_bits, rate_bits, cap_height, proof_of_work_bits, num_query_rounds, C.byref(p))
        if rc != 0:
            raise EtpError(rc, "bad FRI parameters")
        return p


class FriPoly(C.Structure):
    _fields_ = [("oracle_index", C.c_uint32), ("polynomial_index", C.c_uint32)]


class FriBatch(C.Structure):
    _fields_ = [("point", C.c_uint64 * 2), ("polynomials", C.POINTER(FriPoly)), ("n_polynomials", C.c_size_t)]


def lib_path() -> str:
    """In-tree library; ETP_B200_LIB selects another build of the same sources (kernel-variant A/B runs)."""
    return os.environ.get("ETP_B200_LIB") or os.path.join(_HERE, "libetp_b200.so")


def load_library():
    """Loads the CUDA library. Fails loudly: there is no CPU or PyTorch fallback for this path."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise EtpError(-2, f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a). The CUDA extension is the only implementation of this path.")
    L = C.CDLL(path)
    sz, vp, i32 = C.c_size_t, C.c_void_p, C.c_int
    pp = C.POINTER(vp)
    sig = {
        "etp_version": (C.c_char_p, []),
        "etp_device_count": (i32, []),
        "etp_ctx_create": (i32, [i32, pp]),
        "etp_ctx_destroy": (None, [vp]),
        "etp_last_error": (C.c_char_p, [vp]),
        "etp_ctx_synchronize": (i32, [vp]),
        "etp_ctx_stream": (vp, [vp]),
        "etp_ctx_launch_count": (C.c_uint64, [vp]),
        "etp_ctx_trim": (i32, [vp]),
        "etp_ctx_cached_bytes": (C.c_size_t, [vp]),
        "etp_host_poseidon_permute": (None, [C.POINTER(C.c_uint64)]),
        "etp_host_poseidon_gate_wires": (None, [C.POINTER(C.c_uint64), C.c_int, C.POINTER(C.c_uint64)]),
        "etp_bench_pipe_rates": (i32, [vp, C.POINTER(C.c_double)]),
        "etp_host_pin": (i32, [vp, vp, sz]),
        "etp_host_unpin": (i32, [vp, vp]),
        "etp_dev_alloc": (i32, [vp, sz, pp]),
        "etp_dev_free": (i32, [vp, vp]),
        "etp_dev_upload": (i32, [vp, vp, vp, sz]),
        "etp_dev_download": (i32, [vp, vp, vp, sz]),
        "etp_poseidon_permute_host": (i32, [vp, _u64p, sz]),
        "etp_ifft_host": (i32, [vp, _u64p, sz, i32]),
        "etp_fft_host": (i32, [vp, _u64p, sz, i32]),
        "etp_coset_lde_host": (i32, [vp, _u64p, sz, i32, i32, C.c_uint64, _u64p]),
        "etp_coset_ifft_host": (i32, [vp, _u64p, sz, i32, C.c_uint64]),
        "etp_merkle_new_host": (i32, [vp, _u64p, sz, sz, i32, pp]),
        "etp_tree_free": (None, [vp]),
        "etp_tree_num_digests": (sz, [vp]),
        "etp_tree_cap": (i32, [vp, _u64p]),
        "etp_tree_digests": (i32, [vp, _u64p]),
        "etp_tree_prove": (i32, [vp, sz, _u64p]),
        "etp_batch_from_values_host": (i32, [vp, C.POINTER(_u64p), sz, i32, i32, i32, i32, pp]),
        "etp_batch_from_coeffs_host": (i32, [vp, C.POINTER(_u64p), sz, i32, i32, i32, i32, pp]),
        "etp_batch_from_values_dev": (i32, [vp, vp, sz, sz, i32, i32, i32, i32, pp]),
        "etp_batch_from_coeffs_dev": (i32, [vp, vp, sz, sz, i32, i32, i32, i32, pp]),
        "etp_batch_recommit_values_dev": (i32, [vp, vp, sz]),
        "etp_batch_last_commit_timings": (i32, [vp, C.POINTER(C.c_float)]),
        "etp_batch_free": (None, [vp]),
        "etp_batch_num_cols": (sz, [vp]),
        "etp_batch_degree_log": (i32, [vp]),
        "etp_batch_num_digests": (sz, [vp]),
        "etp_batch_cap": (i32, [vp, _u64p]),
        "etp_batch_download_coeffs": (i32, [vp, _u64p]),
        "etp_batch_download_leaves": (i32, [vp, _u64p]),
        "etp_batch_download_digests": (i32, [vp, _u64p]),
        "etp_batch_leaves_at": (i32, [vp, _u64p, sz, _u64p]),
        "etp_batch_get_lde_values": (i32, [vp, sz, sz, _u64p]),
        "etp_batch_prove": (i32, [vp, sz, _u64p]),
        "etp_batch_lde_dev": (vp, [vp, C.POINTER(sz)]),
        "etp_batch_coeffs_dev": (vp, [vp, C.POINTER(sz)]),
        "etp_shard_cols_per_rank": (sz, [sz, i32]),
        "etp_shard_create": (i32, [vp, sz, i32, i32, i32, i32, i32, pp]),
        "etp_shard_free": (None, [vp]),
        "etp_shard_first_col": (sz, [vp]),
        "etp_shard_num_local_cols": (sz, [vp]),
        "etp_shard_first_row": (sz, [vp]),
        "etp_shard_num_rows": (sz, [vp]),
        "etp_shard_lde_dev": (vp, [vp]),
        "etp_shard_transform_values_host": (i32, [vp, C.POINTER(_u64p)]),
        "etp_shard_transform_values_dev": (i32, [vp, vp, sz]),
        "etp_ipc_export": (i32, [vp, vp, C.c_char_p]),
        "etp_ipc_open": (i32, [vp, C.c_char_p, pp]),
        "etp_ipc_close": (i32, [vp, vp]),
        "etp_shard_set_peer": (i32, [vp, i32, vp]),
        "etp_shard_commit_rows": (i32, [vp, _u64p]),
        "etp_shard_prove": (i32, [vp, sz, _u64p]),
        "etp_shard_leaves_at": (i32, [vp, _u64p, sz, _u64p]),
        "etp_shard_download_coeffs": (i32, [vp, _u64p]),
        "etp_table_num_columns": (i32, [vp, i32]),
        "etp_table_constraint_degree": (i32, [vp, i32]),
        "etp_table_num_public_inputs": (i32, [vp, i32]),
        "etp_table_num_aux_columns": (i32, [vp, i32, i32]),
        "etp_table_quotient_degree_factor": (i32, [vp, i32]),
        "etp_table_register": (i32, [vp, _u64p, sz, C.POINTER(C.c_int32), sz, C.POINTER(i32)]),
        "etp_cprog_compile_check": (i32, [_u64p, sz, C.POINTER(sz), C.c_char_p, sz]),
        "etp_cprog_generate_cuda": (C.c_int64, [_u64p, sz, C.c_char_p, sz]),
        "etp_lookup_helper_columns_dev": (i32, [vp, i32, i32, vp, sz, _u64p, i32, vp]),
        "etp_compute_quotient_polys_dev": (i32, [vp, i32, vp, vp, _u64p, i32, _u64p, _u64p, i32, vp]),
        "etp_pow_grind": (i32, [vp, _u64p, i32, i32, _u64p]),
        "etp_stark_proof_words": (sz, [vp, i32, i32]),
        "etp_stark_prove_host": (i32, [vp, i32, i32, _u64p, _u64p, _u64p]),
        "etp_stark_prove_dev": (i32, [vp, i32, i32, vp, sz, _u64p, _u64p]),
        "etp_last_prove_timings": (i32, [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_float), i32]),
        "etp_challenger_init": (None, [C.POINTER(Challenger)]),
        "etp_challenger_observe": (None, [C.POINTER(Challenger), _u64p, sz]),
        "etp_challenger_get_challenge": (C.c_uint64, [C.POINTER(Challenger)]),
        "etp_challenger_get_n_challenges": (None, [C.POINTER(Challenger), sz, _u64p]),
        "etp_challenger_compact": (None, [C.POINTER(Challenger)]),
        "etp_fri_params_make": (i32, [i32, i32, i32, i32, i32, C.POINTER(FriParams)]),
        "etp_batch_eval_at_ext_point": (i32, [vp, _u64p, _u64p]),
        "etp_table_register_ex": (i32, [vp, _u64p, sz, _u64p, sz, C.POINTER(i32)]),
        "etp_table_num_lookup_columns": (i32, [vp, i32, i32]),
        "etp_table_num_ctl_helper_columns": (i32, [vp, i32]),
        "etp_table_num_ctl_zs": (i32, [vp, i32]),
        "etp_aux_columns_dev": (i32, [vp, i32, i32, vp, sz, _u64p, i32, _u64p, vp]),
        "etp_prove_with_commitment": (i32, [vp, i32, vp, vp, sz, _u64p, C.POINTER(Challenger), _u64p, _u64p]),
        "etp_fri_proof_words": (sz, [C.POINTER(sz), sz, C.POINTER(FriParams)]),
        "etp_prove_openings": (i32, [vp, C.POINTER(FriBatch), sz, C.POINTER(vp), sz, C.POINTER(Challenger), C.POINTER(FriParams), _u64p]),
        "etp_fri_begin": (i32, [vp, vp, C.POINTER(FriParams), pp]),
        "etp_fri_commit_layer": (i32, [vp, _u64p]),
        "etp_fri_fold": (i32, [vp, _u64p]),
        "etp_fri_final_poly": (i32, [vp, _u64p]),
        "etp_fri_commit_phase": (i32, [vp, C.POINTER(Challenger), _u64p, _u64p]),
        "etp_fri_query_rounds": (i32, [vp, C.POINTER(vp), sz, _u64p, sz, _u64p]),
        "etp_fri_free": (None, [vp]),
        "etp_fri_proof_of_work": (i32, [vp, C.POINTER(Challenger), i32, _u64p]),
        "etp_poseidon_constants": (None, [_u64p, _u64p, _u64p]),
        "etp_program_register": (i32, [vp, _u64p, sz, C.POINTER(i32)]),
        "etp_circuit_create": (i32, [vp, _u64p, sz, _u64p, i32, _u64p, _u64p, i32, i32, i32, i32, i32, C.POINTER(FriParams), _u64p, pp]),
        "etp_circuit_free": (None, [vp]),
        "etp_circuit_digest": (i32, [vp, _u64p]),
        "etp_circuit_constants_sigmas_cap": (i32, [vp, _u64p]),
        "etp_circuit_proof_words": (sz, [vp]),
        "etp_circuit_prove_host": (i32, [vp, _u64p, _u64p, _u64p]),
        "etp_circuit_prove_dev": (i32, [vp, vp, sz, _u64p, _u64p]),
        "etp_compute_quotient_polys_cols_dev": (i32, [vp, i32, C.POINTER(vp), sz, i32, i32, _u64p, i32, _u64p, _u64p, i32, vp]),
        "etp_shard_aux_columns_dev": (i32, [vp, i32, _u64p, i32, _u64p, vp, _u64p]),
        "etp_shard_compute_quotient_polys_dev": (i32, [vp, i32, vp, _u64p, i32, _u64p, _u64p, i32, vp]),
        "etp_shard_eval_at_ext_points": (i32, [vp, _u64p, _u64p, _u64p, _u64p]),
        "etp_shard_fri_begin": (i32, [vp, C.POINTER(vp), sz, C.POINTER(FriBatch), sz, _u64p, _u64p, C.POINTER(FriParams), pp]),
        "etp_plonk_partial_products_and_zs_dev": (i32, [vp, vp, sz, vp, sz, _u64p, i32, i32, i32, _u64p, _u64p, i32, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _LIB = L
    L._etp_signatures = sig
    return L


def _fri_batches(batches):
    """[(point (2,), [(oracle_index, polynomial_index)])] -> (etp_fri_batch array, objects to keep alive)."""
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
    return arr, keep


def poseidon_constants():
    """(ALL_ROUND_CONSTANTS (30, 12), MDS_MATRIX_CIRC (12,), MDS_MATRIX_DIAG (12,)) as Python ints, from the library."""
    rc, circ, diag = np.zeros(360, dtype=np.uint64), np.zeros(12, dtype=np.uint64), np.zeros(12, dtype=np.uint64)
    load_library().etp_poseidon_constants(_p(rc), _p(circ), _p(diag))
    return [[int(x) for x in rc[12 * r:12 * r + 12]] for r in range(30)], [int(x) for x in circ], [int(x) for x in diag]


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"], "need a C-contiguous uint64 array"
    return a.ctypes.data_as(_u64p)


def _u64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint64))


class Context:
    """One CUDA device + stream + scratch pools (one per worker thread / Paladin worker)."""

    def __init__(self, device: int = 0):
        self.L = load_library()
        h = C.c_void_p()
        rc = self.L.etp_ctx_create(device, C.byref(h))
        if rc != 0:
            raise EtpError(rc, f"etp_ctx_create(device={device}) failed: no usable CUDA device "
                               f"(devices visible: {self.L.etp_device_count()}); there is no CPU fallback")
        self.h = h
        self.device = device

    def check(self, rc):
        if rc != 0:
            raise EtpError(rc, self.L.etp_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.etp_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        self.check(self.L.etp_ctx_synchronize(self.h))

    @property
    def stream(self) -> int:
        return int(self.L.etp_ctx_stream(self.h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self.L.etp_ctx_launch_count(self.h))

    def pin(self, array: np.ndarray):
        """Page-lock a host array in place (etp_host_pin); call unpin(array) before it is freed."""
        self.check(self.L.etp_host_pin(self.h, C.c_void_p(array.ctypes.data), array.nbytes))

    def unpin(self, array: np.ndarray):
        self.check(self.L.etp_host_unpin(self.h, C.c_void_p(array.ctypes.data)))

    def trim(self):
        """Return the context's cached device blocks to the CUDA runtime (etp_ctx_trim)."""
        self.check(self.L.etp_ctx_trim(self.h))

    @property
    def cached_bytes(self) -> int:
        return int(self.L.etp_ctx_cached_bytes(self.h))

    def pipe_rates(self) -> dict:
        """Measured issue rates (thread-instructions/s, whole device) of the integer / FP64 pipes (etp_bench_pipe_rates)."""
        r = (C.c_double * 3)()
        self.check(self.L.etp_bench_pipe_rates(self.h, r))
        return {"imad_wide_u32_zero_addend": r[0], "mad_wide_u32_accumulate": r[1], "dfma": r[2]}

    # ---- primitives (parity tests)
    def poseidon_permute(self, states) -> np.ndarray:
        s = _u64(states).reshape(-1, 12).copy()
        self.check(self.L.etp_poseidon_permute_host(self.h, _p(s), s.shape[0]))
        return s

    def ifft(self, cols) -> np.ndarray:
        a = np.atleast_2d(_u64(cols)).copy()
        self.check(self.L.etp_ifft_host(self.h, _p(a), a.shape[0], int(a.shape[1]).bit_length() - 1))
        return a

    def fft(self, cols) -> np.ndarray:
        a = np.atleast_2d(_u64(cols)).copy()
        self.check(self.L.etp_fft_host(self.h, _p(a), a.shape[0], int(a.shape[1]).bit_length() - 1))
        return a

    def coset_lde(self, coeffs, rate_bits, shift=7) -> np.ndarray:
        a = np.atleast_2d(_u64(coeffs))
        out = np.zeros((a.shape[0], a.shape[1] << rate_bits), dtype=np.uint64)
        self.check(self.L.etp_coset_lde_host(self.h, _p(a), a.shape[0], int(a.shape[1]).bit_length() - 1, rate_bits,
                                             shift, _p(out)))
        return out

    def coset_ifft(self, cols, shift=7) -> np.ndarray:
        a = np.atleast_2d(_u64(cols)).copy()
        self.check(self.L.etp_coset_ifft_host(self.h, _p(a), a.shape[0], int(a.shape[1]).bit_length() - 1, shift))
        return a

    def pow_grind(self, state, pos, bits) -> int:
        out = C.c_uint64()
        self.check(self.L.etp_pow_grind(self.h, _p(_u64(state)), pos, bits, C.byref(out)))
        return int(out.value)

    # ---- starky
    def fri_proof_of_work(self, challenger: Challenger, proof_of_work_bits: int) -> int:
        """fri_proof_of_work: grind, observe the witness, draw (and check) the response; returns the witness."""
        out = np.zeros(1, dtype=np.uint64)
        self.check(self.L.etp_fri_proof_of_work(self.h, C.byref(challenger), proof_of_work_bits, _p(out)))
        return int(out[0])

    def compute_quotient_polys_cols_dev(self, table, lde_cols, log_n, rate_bits, challenge_scalars, public_inputs, alphas, out_ptr: int):
        """compute_quotient_polys over a list of LDE column pointers (etp_compute_quotient_polys_cols_dev)."""
        cols = (C.c_void_p * len(lde_cols))(*[int(c) for c in lde_cols])
        sc = _u64(list(challenge_scalars) + [0])
        pi = _u64(list(public_inputs) + [0])
        a = _u64(alphas)
        self.check(self.L.etp_compute_quotient_polys_cols_dev(self.h, table, cols, len(lde_cols), log_n, rate_bits, _p(sc), len(challenge_scalars),
                                                              _p(pi), _p(a), a.size, C.c_void_p(out_ptr)))

    def stark_proof_words(self, table, log_n) -> int:
        return int(self.L.etp_stark_proof_words(self.h, table, log_n))

    def _proof_buffer(self, table, log_n) -> np.ndarray:
        words = self.stark_proof_words(table, log_n)
        if words == 0:
            raise EtpError(-1, f"no proof shape for table {table} at degree_bits {log_n}: unknown table, or cap_height=4 should be at "
                               f"most log2(leaves.len())={log_n + 1}")
        return np.zeros(words, dtype=np.uint64)

    def register_table(self, program, lookups=None) -> int:
        """Registers a table from its constraint program (``cprog.Program.words`` or a u64 array) and its lookups
        (``[(looking_columns, table_column, frequencies_column), ...]``): NVRTC-compiles the quotient kernel for
        sm_100a.  Returns the table id accepted by stark_prove / compute_quotient_polys / lookup_helper_columns."""
        if hasattr(program, "simple_lookups") and not program.simple_lookups():
            return self.register_table_ex(program, program.aux_spec)
        if lookups is None:  # a cprog.Program carries its own lookups
            lookups = getattr(program, "lookups", ())
        words = _u64(getattr(program, "words", program))
        flat = [len(lookups)]
        for looking, table_col, freq_col in lookups:
            flat += [int(table_col), int(freq_col), len(looking)] + [int(c) for c in looking]
        arr = (C.c_int32 * len(flat))(*flat)
        out = C.c_int(0)
        self.check(self.L.etp_table_register(self.h, _p(words), words.size, arr, len(flat) if lookups else 0, C.byref(out)))
        return int(out.value)

    def register_program(self, program) -> int:
        """A constraint program that is not a starky table (etp_program_register): evaluated by compute_quotient_polys_cols_dev."""
        words = _u64(getattr(program, "words", program))
        out = C.c_int(0)
        self.check(self.L.etp_program_register(self.h, _p(words), words.size, C.byref(out)))
        return int(out.value)

    def register_table_ex(self, program, aux_spec) -> int:
        """General registration (etp_table_register_ex): lookups with linear-combination Columns and Filters and the
        table's CTL Z descriptors, as the u64 words of ``cprog.Program.aux_spec``."""
        words = _u64(getattr(program, "words", program))
        spec = _u64(aux_spec)
        out = C.c_int(0)
        self.check(self.L.etp_table_register_ex(self.h, _p(words), words.size, _p(spec) if spec.size else None, spec.size, C.byref(out)))
        return int(out.value)

    def aux_columns_dev(self, table, log_n, trace_ptr, col_stride, lookup_challenges, ctl_challenges, aux_ptr):
        """All auxiliary polynomials of a table (lookup columns ++ CTL helper columns ++ CTL Zs) on the device."""
        lc = _u64(list(lookup_challenges) + [0])
        cc = _u64(ctl_challenges) if ctl_challenges is not None else None
        self.check(self.L.etp_aux_columns_dev(self.h, table, log_n, C.c_void_p(trace_ptr), col_stride, _p(lc), len(lookup_challenges),
                                              _p(cc) if cc is not None else None, C.c_void_p(aux_ptr)))

    def prove_with_commitment(self, table, trace_batch, trace_ptr, col_stride, challenger: Challenger, ctl_challenges=None,
                              public_inputs=()) -> np.ndarray:
        """starky::prover::prove_with_commitment: pre-committed trace, CTL challenges (or None), challenger state in/out."""
        n_pi = int(self.L.etp_table_num_public_inputs(self.h, table))
        if n_pi < 0 or len(public_inputs) < n_pi:
            raise EtpError(-1, "unknown table or too few public inputs")
        pi = _u64(list(public_inputs) + [0])
        cc = _u64(ctl_challenges) if ctl_challenges is not None else None
        if cc is not None and cc.size != 4:
            raise EtpError(-1, "ctl_challenges must be num_challenges (beta, gamma) pairs = 4 words")
        out = self._proof_buffer(table, trace_batch.degree_log)
        self.check(self.L.etp_prove_with_commitment(self.h, table, trace_batch.h, C.c_void_p(trace_ptr), col_stride,
                                                    _p(cc) if cc is not None else None, C.byref(challenger), _p(pi), _p(out)))
        return out

    def prove_openings(self, batches, oracles, challenger: Challenger, params: FriParams) -> np.ndarray:
        """PolynomialBatch::prove_openings(instance, oracles, challenger, fri_params) -> flat FriProof.
        batches: [(point (2,), [(oracle_index, polynomial_index)])] — the FriInstanceInfo's FriBatchInfo list."""
        arr, keep = _fri_batches(batches)
        oc = (C.c_size_t * len(oracles))(*[o.n_cols for o in oracles])
        out = np.zeros(int(self.L.etp_fri_proof_words(oc, len(oracles), C.byref(params))), dtype=np.uint64)
        hs = (C.c_void_p * len(oracles))(*[o.h for o in oracles])
        self.check(self.L.etp_prove_openings(self.h, arr, len(batches), hs, len(oracles), C.byref(challenger), C.byref(params), _p(out)))
        return out

    def plonk_partial_products_and_zs_dev(self, wires_ptr, wires_stride, sigmas_ptr, sigmas_stride, k_is, degree_bits,
                                          quotient_degree_factor, betas, gammas, out_ptr):
        """plonky2::plonk::prover::all_wires_permutation_partial_products on the device (routed wires + sigma values in,
        [Z per challenge] ++ [partial products per challenge] out, column-major with stride 2^degree_bits)."""
        k, b, g = _u64(k_is), _u64(betas), _u64(gammas)
        self.check(self.L.etp_plonk_partial_products_and_zs_dev(self.h, C.c_void_p(wires_ptr), wires_stride, C.c_void_p(sigmas_ptr),
                                                                sigmas_stride, _p(k), k.size, degree_bits, quotient_degree_factor,
                                                                _p(b), _p(g), b.size, C.c_void_p(out_ptr)))

    def table_num_aux_columns(self, table, num_challenges=2) -> int:
        return int(self.L.etp_table_num_aux_columns(self.h, table, num_challenges))

    def stark_prove(self, table, trace, public_inputs=()) -> np.ndarray:
        """starky::prover::prove(stark, &StarkConfig::standard_fast_config(), trace, public_inputs)."""
        t = _u64(trace)
        log_n = self._check_trace_shape(table, t.shape, public_inputs)
        pi = _u64(list(public_inputs) + [0])
        out = self._proof_buffer(table, log_n)
        self.check(self.L.etp_stark_prove_host(self.h, table, log_n, _p(t), _p(pi), _p(out)))
        return out

    def _check_trace_shape(self, table, shape, public_inputs) -> int:
        """The C side reads n_cols x 2^log_n trace words and n_pi public inputs unconditionally: reject anything else
        here (upstream: assert! panics in prove / PolynomialBatch::from_values)."""
        n_cols = int(self.L.etp_table_num_columns(self.h, table))
        if n_cols < 0:
            raise EtpError(-1, f"unknown table {table}")
        if len(shape) != 2 or shape[0] != n_cols:
            raise EtpError(-1, f"trace must be ({n_cols}, 2^k) column-major for table {table}, got {tuple(shape)}")
        log_n = int(shape[1]).bit_length() - 1
        if shape[1] < 2 or shape[1] != 1 << log_n:
            raise EtpError(-1, f"trace length {shape[1]} is not a power of two >= 2")
        n_pi = int(self.L.etp_table_num_public_inputs(self.h, table))
        if len(public_inputs) < n_pi:
            raise EtpError(-1, f"table {table} takes {n_pi} public inputs, got {len(public_inputs)}")
        return log_n

    def stark_prove_dev(self, table, log_n, trace_ptr, col_stride, public_inputs=()) -> np.ndarray:
        n_pi = int(self.L.etp_table_num_public_inputs(self.h, table))
        if n_pi < 0 or len(public_inputs) < n_pi or col_stride < (1 << log_n) or not trace_ptr:
            raise EtpError(-1, "stark_prove_dev: unknown table, too few public inputs, null trace or stride < 2^log_n")
        pi = _u64(list(public_inputs) + [0])
        out = self._proof_buffer(table, log_n)
        self.check(self.L.etp_stark_prove_dev(self.h, table, log_n, C.c_void_p(trace_ptr), col_stride, _p(pi), _p(out)))
        return out

    def last_prove_timings(self) -> dict:
        names = (C.c_char_p * 32)()
        ms = (C.c_float * 32)()
        n = self.L.etp_last_prove_timings(self.h, names, ms, 32)
        return {names[i].decode(): float(ms[i]) for i in range(n)}

    def lookup_helper_columns_dev(self, table, log_n, trace_ptr, col_stride, challenges, aux_ptr):
        ch = _u64(challenges)
        self.check(self.L.etp_lookup_helper_columns_dev(self.h, table, log_n, C.c_void_p(trace_ptr), col_stride, _p(ch),
                                                        ch.size, C.c_void_p(aux_ptr)))

    def compute_quotient_polys(self, table, trace_batch, aux_batch, lookup_challenges, public_inputs, alphas) -> np.ndarray:
        """starky::prover::compute_quotient_polys -> quotient chunks (factor * n_alphas, n), host copy."""
        al = _u64(alphas)
        lc = _u64(list(lookup_challenges) + [0])
        pi = _u64(list(public_inputs) + [0])
        factor = self.L.etp_table_quotient_degree_factor(self.h, table)
        n = 1 << trace_batch.degree_log
        out = np.zeros((factor * al.size, n), dtype=np.uint64)
        d = C.c_void_p()
        self.check(self.L.etp_dev_alloc(self.h, out.nbytes, C.byref(d)))
        try:
            self.check(self.L.etp_compute_quotient_polys_dev(self.h, table, trace_batch.h, aux_batch.h if aux_batch else None,
                                                             _p(lc), len(lookup_challenges), _p(pi), _p(al), al.size, d))
            self.check(self.L.etp_dev_download(self.h, out.ctypes.data_as(C.c_void_p), d, out.nbytes))
        finally:
            self.L.etp_dev_free(self.h, d)
        return out


class MerkleTree:
    """plonky2::hash::merkle_tree::MerkleTree<GoldilocksField, PoseidonHash>."""

    def __init__(self, ctx: Context, handle, n_leaves, cap_height):
        self.ctx, self.h, self.n_leaves, self.cap_height = ctx, handle, n_leaves, cap_height

    @classmethod
    def new(cls, ctx: Context, leaves, cap_height: int) -> "MerkleTree":
        lv = _u64(leaves)
        assert lv.ndim == 2
        h = C.c_void_p()
        ctx.check(ctx.L.etp_merkle_new_host(ctx.h, _p(lv) if lv.size else None, lv.shape[0], lv.shape[1], cap_height,
                                            C.byref(h)))
        return cls(ctx, h, lv.shape[0], cap_height)

    @property
    def cap(self) -> np.ndarray:
        out = np.zeros((1 << self.cap_height, 4), dtype=np.uint64)
        self.ctx.check(self.ctx.L.etp_tree_cap(self.h, _p(out)))
        return out

    @property
    def digests(self) -> np.ndarray:
        nd = int(self.ctx.L.etp_tree_num_digests(self.h))
        out = np.zeros((max(nd, 1), 4), dtype=np.uint64)
        self.ctx.check(self.ctx.L.etp_tree_digests(self.h, _p(out)))
        return out[:nd]

    def prove(self, leaf_index: int) -> np.ndarray:
        ns = (int(self.n_leaves).bit_length() - 1) - self.cap_height
        out = np.zeros((max(ns, 1), 4), dtype=np.uint64)
        self.ctx.check(self.ctx.L.etp_tree_prove(self.h, leaf_index, _p(out)))
        return out[:ns]

    def __del__(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.L.etp_tree_free(self.h)
            self.h = None


class PolynomialBatch:
    """plonky2::fri::oracle::PolynomialBatch<GoldilocksField, PoseidonGoldilocksConfig, 2>, device resident.

    The public fields of the upstream struct are lazy host views here: ``polynomials`` (coefficients),
    ``leaves`` / ``digests`` / ``cap`` of ``merkle_tree``.
    """

    def __init__(self, ctx: Context, handle, n_cols, log_n, rate_bits, cap_height):
        self.ctx, self.h = ctx, handle
        self.n_cols, self.degree_log, self.rate_bits, self.cap_height = n_cols, log_n, rate_bits, cap_height

    @staticmethod
    def _cols(a):
        a = _u64(a)
        assert a.ndim == 2
        ptrs = (_u64p * max(a.shape[0], 1))()
        for c in range(a.shape[0]):
            ptrs[c] = a[c].ctypes.data_as(_u64p)
        return a, ptrs

    @classmethod
    def from_values(cls, ctx: Context, values, rate_bits: int, blinding: bool, cap_height: int) -> "PolynomialBatch":
        a, ptrs = cls._cols(values)
        log_n = int(a.shape[1]).bit_length() - 1
        assert a.shape[1] == 1 << log_n
        h = C.c_void_p()
        ctx.check(ctx.L.etp_batch_from_values_host(ctx.h, ptrs, a.shape[0], log_n, rate_bits, int(blinding), cap_height,
                                                   C.byref(h)))
        return cls(ctx, h, a.shape[0], log_n, rate_bits, cap_height)

    @classmethod
    def from_coeffs(cls, ctx: Context, coeffs, rate_bits: int, blinding: bool, cap_height: int) -> "PolynomialBatch":
        a, ptrs = cls._cols(coeffs)
        log_n = int(a.shape[1]).bit_length() - 1
        assert a.shape[1] == 1 << log_n
        h = C.c_void_p()
        ctx.check(ctx.L.etp_batch_from_coeffs_host(ctx.h, ptrs, a.shape[0], log_n, rate_bits, int(blinding), cap_height,
                                                   C.byref(h)))
        return cls(ctx, h, a.shape[0], log_n, rate_bits, cap_height)

    @classmethod
    def from_values_dev(cls, ctx: Context, ptr: int, col_stride: int, n_cols: int, log_n: int, rate_bits: int,
                        blinding: bool, cap_height: int) -> "PolynomialBatch":
        h = C.c_void_p()
        ctx.check(ctx.L.etp_batch_from_values_dev(ctx.h, C.c_void_p(ptr), col_stride, n_cols, log_n, rate_bits,
                                                  int(blinding), cap_height, C.byref(h)))
        return cls(ctx, h, n_cols, log_n, rate_bits, cap_height)

    @classmethod
    def from_coeffs_dev(cls, ctx: Context, ptr: int, col_stride: int, n_cols: int, log_n: int, rate_bits: int,
                        blinding: bool, cap_height: int) -> "PolynomialBatch":
        h = C.c_void_p()
        ctx.check(ctx.L.etp_batch_from_coeffs_dev(ctx.h, C.c_void_p(ptr), col_stride, n_cols, log_n, rate_bits,
                                                  int(blinding), cap_height, C.byref(h)))
        return cls(ctx, h, n_cols, log_n, rate_bits, cap_height)

    def recommit_values_dev(self, ptr: int, col_stride: int):
        self.ctx.check(self.ctx.L.etp_batch_recommit_values_dev(self.h, C.c_void_p(ptr), col_stride))

    def last_commit_timings(self) -> dict:
        ms = (C.c_float * 4)()
        self.ctx.check(self.ctx.L.etp_batch_last_commit_timings(self.h, ms))
        return {"IFFT": ms[0], "FFT + blinding": ms[1], "build Merkle tree (leaves)": ms[2],
                "build Merkle tree (levels)": ms[3]}

    @property
    def n(self):
        return 1 << self.degree_log

    @property
    def lde_n(self):
        return 1 << (self.degree_log + self.rate_bits)

    @property
    def cap(self) -> np.ndarray:
        out = np.zeros((1 << self.cap_height, 4), dtype=np.uint64)
        self.ctx.check(self.ctx.L.etp_batch_cap(self.h, _p(out)))
        return out

    @property
    def polynomials(self) -> np.ndarray:
        out = np.zeros((self.n_cols, self.n), dtype=np.uint64)
        self.ctx.check(self.ctx.L.etp_batch_download_coeffs(self.h, _p(out)))
        return out

    @property
    def leaves(self) -> np.ndarray:
        out = np.zeros((self.lde_n, self.n_cols), dtype=np.uint64)
        self.ctx.check(self.ctx.L.etp_batch_download_leaves(self.h, _p(out)))
        return out

    @property
    def digests(self) -> np.ndarray:
        nd = int(self.ctx.L.etp_batch_num_digests(self.h))
        out = np.zeros((max(nd, 1), 4), dtype=np.uint64)
        self.ctx.check(self.ctx.L.etp_batch_download_digests(self.h, _p(out)))
        return out[:nd]

    def leaves_at(self, idx) -> np.ndarray:
        idx = _u64(idx).ravel()
        out = np.zeros((idx.size, max(self.n_cols, 1)), dtype=np.uint64)
        self.ctx.check(self.ctx.L.etp_batch_leaves_at(self.h, _p(idx), idx.size, _p(out)))
        return out[:, :self.n_cols]

    def get_lde_values(self, index: int, step: int) -> np.ndarray:
        out = np.zeros(max(self.n_cols, 1), dtype=np.uint64)
        self.ctx.check(self.ctx.L.etp_batch_get_lde_values(self.h, index, step, _p(out)))
        return out[:self.n_cols]

    def prove(self, leaf_index: int) -> np.ndarray:
        ns = self.degree_log + self.rate_bits - self.cap_height
        out = np.zeros((max(ns, 1), 4), dtype=np.uint64)
        self.ctx.check(self.ctx.L.etp_batch_prove(self.h, leaf_index, _p(out)))
        return out[:ns]

    def eval_at_ext_point(self, z) -> np.ndarray:
        """polynomials[c].to_extension().eval(z) for every polynomial: (n_cols, 2)."""
        out = np.zeros((max(self.n_cols, 1), 2), dtype=np.uint64)
        self.ctx.check(self.ctx.L.etp_batch_eval_at_ext_point(self.h, _p(_u64(z)), _p(out)))
        return out[:self.n_cols]

    def __del__(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.L.etp_batch_free(self.h)
            self.h = None


class FriState:
    """The FRI prover step by step (plonky2::fri::prover::fri_committed_trees / fri_prover_query_rounds): commit a layer, get
    its cap, fold with the beta the caller's challenger produced, ... — or the fused commit phase with the challenger."""

    def __init__(self, ctx: Context, values_ptr: int, params: FriParams, handle=None):
        self.ctx, self.params = ctx, params
        if handle is None:
            handle = C.c_void_p()
            ctx.check(ctx.L.etp_fri_begin(ctx.h, C.c_void_p(values_ptr), C.byref(params), C.byref(handle)))
        self.h = handle

    def commit_layer(self) -> np.ndarray:
        cap = np.zeros((1 << self.params.cap_height, 4), dtype=np.uint64)
        self.ctx.check(self.ctx.L.etp_fri_commit_layer(self.h, _p(cap)))
        return cap

    def fold(self, beta):
        self.ctx.check(self.ctx.L.etp_fri_fold(self.h, _p(_u64(beta))))

    @property
    def final_poly_len(self) -> int:
        p = self.params
        return 1 << (p.degree_bits - sum(p.reduction_arity_bits[i] for i in range(p.n_reductions)))

    def final_poly(self) -> np.ndarray:
        out = np.zeros((self.final_poly_len, 2), dtype=np.uint64)
        self.ctx.check(self.ctx.L.etp_fri_final_poly(self.h, _p(out)))
        return out

    def commit_phase(self, challenger: Challenger):
        """-> (caps (n_reductions, 2^cap, 4), final_poly (len, 2)); the challenger is advanced."""
        caps = np.zeros((max(self.params.n_reductions, 1), 1 << self.params.cap_height, 4), dtype=np.uint64)
        fin = np.zeros((self.final_poly_len, 2), dtype=np.uint64)
        self.ctx.check(self.ctx.L.etp_fri_commit_phase(self.h, C.byref(challenger), _p(caps), _p(fin)))
        return caps[:self.params.n_reductions], fin

    def query_rounds(self, oracles, x_indices) -> np.ndarray:
        p = self.params
        log_lde = p.degree_bits + p.rate_bits
        per = sum(o.n_cols + 4 * (log_lde - p.cap_height) for o in oracles)
        bits = log_lde
        for i in range(p.n_reductions):
            bits -= p.reduction_arity_bits[i]
            per += 2 * (1 << p.reduction_arity_bits[i]) + 4 * (bits - p.cap_height)
        idx = _u64(x_indices).ravel()
        out = np.zeros((max(idx.size, 1), per), dtype=np.uint64)
        hs = (C.c_void_p * max(len(oracles), 1))(*[o.h for o in oracles])
        self.ctx.check(self.ctx.L.etp_fri_query_rounds(self.h, hs, len(oracles), _p(idx), idx.size, _p(out)))
        return out[:idx.size]

    def __del__(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.L.etp_fri_free(self.h)
            self.h = None


IPC_HANDLE_BYTES = 64


class BatchShard:
    """Rank `rank` of `world`'s share of ONE column-split PolynomialBatch: the local columns'
    coefficients + LDE, and the Merkle subtrees of the local leaf rows.  `parallel.commit_column_split` drives
    the protocol across processes; within one process (several contexts) peers are wired with `set_peer`."""

    def __init__(self, ctx: Context, n_cols_total: int, log_n: int, rate_bits: int, cap_height: int, rank: int, world: int):
        self.ctx = ctx
        h = C.c_void_p()
        ctx.check(ctx.L.etp_shard_create(ctx.h, n_cols_total, log_n, rate_bits, cap_height, rank, world, C.byref(h)))
        self.h = h
        self.n_cols_total, self.degree_log, self.rate_bits, self.cap_height = n_cols_total, log_n, rate_bits, cap_height
        self.rank, self.world = rank, world
        self.first_col = int(ctx.L.etp_shard_first_col(h))
        self.num_local_cols = int(ctx.L.etp_shard_num_local_cols(h))
        self.first_row = int(ctx.L.etp_shard_first_row(h))
        self.num_rows = int(ctx.L.etp_shard_num_rows(h))
        self._opened = []

    @property
    def lde_ptr(self) -> int:
        return int(self.ctx.L.etp_shard_lde_dev(self.h) or 0)

    def transform_values(self, local_values):
        """iFFT + coset LDE of the local columns (host array, num_local_cols x n)."""
        a, ptrs = PolynomialBatch._cols(local_values)
        assert a.shape[0] == self.num_local_cols and (a.size == 0 or a.shape[1] == 1 << self.degree_log)
        self.ctx.check(self.ctx.L.etp_shard_transform_values_host(self.h, ptrs))

    def transform_values_dev(self, ptr: int, col_stride: int):
        self.ctx.check(self.ctx.L.etp_shard_transform_values_dev(self.h, C.c_void_p(ptr), col_stride))

    def export_handle(self) -> bytes:
        buf = C.create_string_buffer(IPC_HANDLE_BYTES)
        self.ctx.check(self.ctx.L.etp_ipc_export(self.ctx.h, C.c_void_p(self.lde_ptr), buf))
        return buf.raw

    def open_peer(self, peer_rank: int, handle: bytes):
        p = C.c_void_p()
        self.ctx.check(self.ctx.L.etp_ipc_open(self.ctx.h, handle, C.byref(p)))
        self._opened.append(p)
        self.set_peer(peer_rank, p.value)

    def set_peer(self, peer_rank: int, ptr: int):
        self.ctx.check(self.ctx.L.etp_shard_set_peer(self.h, peer_rank, C.c_void_p(ptr)))

    def commit_rows(self) -> np.ndarray:
        """Leaf hashing of the own rows (peers' columns read over NVLink) + own subtrees -> own cap entries."""
        out = np.zeros(((1 << self.cap_height) // self.world, 4), dtype=np.uint64)
        self.ctx.check(self.ctx.L.etp_shard_commit_rows(self.h, _p(out)))
        return out

    def prove(self, leaf_index: int) -> np.ndarray:
        ns = self.degree_log + self.rate_bits - self.cap_height
        out = np.zeros((max(ns, 1), 4), dtype=np.uint64)
        self.ctx.check(self.ctx.L.etp_shard_prove(self.h, leaf_index, _p(out)))
        return out[:ns]

    def leaves_at(self, idx) -> np.ndarray:
        idx = _u64(idx).ravel()
        out = np.zeros((idx.size, self.n_cols_total), dtype=np.uint64)
        self.ctx.check(self.ctx.L.etp_shard_leaves_at(self.h, _p(idx), idx.size, _p(out)))
        return out

    @property
    def polynomials(self) -> np.ndarray:
        out = np.zeros((max(self.num_local_cols, 1), 1 << self.degree_log), dtype=np.uint64)
        self.ctx.check(self.ctx.L.etp_shard_download_coeffs(self.h, _p(out)))
        return out[:self.num_local_cols]

    def aux_columns_dev(self, table, lookup_challenges, ctl_challenges, aux_ptr: int) -> np.ndarray:
        """Every auxiliary polynomial of the table (values on the trace domain) into the device matrix `aux_ptr`
        (num_aux_columns x n), computed on THIS rank from the mapped LDE; returns ctl_zs_first."""
        lc = _u64(list(lookup_challenges) + [0])
        cc = _u64(ctl_challenges) if ctl_challenges is not None else None
        n_zs = max(int(self.ctx.L.etp_table_num_ctl_zs(self.ctx.h, table)), 0)
        zs = np.zeros(max(n_zs, 1), dtype=np.uint64)
        self.ctx.check(self.ctx.L.etp_shard_aux_columns_dev(self.h, table, _p(lc), len(lookup_challenges), _p(cc) if cc is not None else None,
                                                            C.c_void_p(aux_ptr), _p(zs)))
        return zs[:n_zs]

    def compute_quotient_polys_dev(self, table, aux_batch, challenge_scalars, public_inputs, alphas, out_ptr: int):
        """compute_quotient_polys over the split trace on THIS rank (peers' columns over NVLink); out: device matrix of
        num_challenges * quotient_degree_factor polynomials x n."""
        pi = _u64(list(public_inputs) + [0])
        sc = _u64(list(challenge_scalars) + [0])
        a = _u64(alphas)
        self.ctx.check(self.ctx.L.etp_shard_compute_quotient_polys_dev(self.h, table, aux_batch.h if aux_batch is not None else None, _p(sc),
                                                                       len(challenge_scalars), _p(pi), _p(a), a.size, C.c_void_p(out_ptr)))

    def eval_at_ext_points(self, z0, z1):
        """The local columns' polynomials at z0 and z1 -> two (num_local_cols, 2) arrays."""
        o0 = np.zeros((max(self.num_local_cols, 1), 2), dtype=np.uint64)
        o1 = np.zeros_like(o0)
        self.ctx.check(self.ctx.L.etp_shard_eval_at_ext_points(self.h, _p(_u64(z0)), _p(_u64(z1)), _p(o0), _p(o1)))
        return o0[:self.num_local_cols], o1[:self.num_local_cols]

    def fri_begin(self, extra_oracles, batches, ys, alpha, params: FriParams) -> "FriState":
        """The combination step of prove_openings with oracle 0 = this split table, oracles 1.. = `extra_oracles`
        (PolynomialBatches of this rank).  ys: per batch, the (n_polynomials, 2) claimed openings."""
        arr, keep = _fri_batches(batches)
        flat = _u64(np.concatenate([_u64(y).reshape(-1) for y in ys]) if ys else [])
        hs = (C.c_void_p * max(len(extra_oracles), 1))(*[o.h for o in extra_oracles])
        h = C.c_void_p()
        self.ctx.check(self.ctx.L.etp_shard_fri_begin(self.h, hs, len(extra_oracles), arr, len(batches), _p(flat), _p(_u64(alpha)),
                                                      C.byref(params), C.byref(h)))
        del keep
        return FriState(self.ctx, 0, params, handle=h)

    def close_peers(self):
        for p in self._opened:
            self.ctx.L.etp_ipc_close(self.ctx.h, p)
        self._opened = []

    def __del__(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.close_peers()
            self.ctx.L.etp_shard_free(self.h)
            self.h = None
