"""CPU tests of the product library's host side: the C-ABI shared library loads and exports every
symbol include/etp_b200.h declares, refuses to run without a GPU (no CPU fallback), and carries the
pinned Poseidon round constants.  No compute calls are made here."""
import ctypes
import hashlib
import os
import re
import struct
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "etp_b200.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(etp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import eth_tx_proof_b200 as etp

    L = etp.load_library()
    names = _declared_symbols()
    assert len(names) >= 45
    for name in names:
        assert hasattr(L, name), f"{name} is declared in include/etp_b200.h but not exported"
    # and the Python mirror binds exactly the declared surface
    assert sorted(L._etp_signatures) == names


def test_no_cpu_fallback_without_a_gpu():
    import eth_tx_proof_b200 as etp

    L = etp.load_library()
    if L.etp_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(etp.EtpError):
        etp.Context(0)
    h = ctypes.c_void_p()
    assert L.etp_ctx_create(0, ctypes.byref(h)) != 0 and not h.value


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "eth_tx_proof_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), fn
                assert not re.search(r"#\s*include\s*[\"<][^\">]*oracle", src), fn
                assert "liboracle" not in src and "orc_" not in src, fn


def test_embedded_round_constants_are_the_pinned_ones():
    import json

    with open(os.path.join(ROOT, "eth_tx_proof_b200", "csrc", "poseidon_constants.h")) as f:
        src = f.read()
    table = src.split("#define ETP_POSEIDON_RC_TABLE")[1].split("}")[0]
    vals = [int(x, 16) for x in re.findall(r"0x([0-9a-f]{16})ULL", table)]
    assert len(vals) == 360
    with open(os.path.join(ROOT, "tests", "golden", "poseidon_kat.json")) as f:
        kat = json.load(f)
    assert hashlib.sha256(b"".join(struct.pack("<Q", v) for v in vals)).hexdigest() == kat["round_constants_sha256_le_u64"]
    # the FP64-pipe start-value tables are the ones tools/gen_poseidon_constants.py derives from those constants
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_poseidon_constants as gen

    _, _, _, _, full, pair = gen.tables()
    for name, rows in (("ETP_POSEIDON_FULL_F64_TABLE", full), ("ETP_POSEIDON_PAIR_F64_TABLE", pair)):
        body = src.split("#define " + name)[1].split("}")[0]
        bits = [int(x, 16) for x in re.findall(r"0x([0-9a-f]{16})ULL", body)]
        want = [v for layer in rows for plane in layer for v in plane]
        assert len(bits) == len(want)
        for b, w in zip(bits, want):
            assert struct.unpack("<d", struct.pack("<Q", b))[0] == float(w) and abs(w) < 2**53


def test_synthetic_traces_are_deterministic_and_valid():
    import numpy as np

    import oracle
    from eth_tx_proof_b200 import synthetic as syn

    a, b = syn.memory_trace(9, seed=3), syn.memory_trace(9, seed=3)
    assert (a == b).all() and a.shape == (21, 512)
    assert oracle.check_constraints(oracle.TABLE_MEMORY, a) == -1
    c = syn.random_columns(3, 8)
    assert (c < np.uint64(0xFFFFFFFF00000001)).all()


def test_host_transcript_permutation_matches_the_known_answers_and_the_oracle():
    """The product's host-side Poseidon (csrc/host_poseidon.cpp, used by the Fiat-Shamir Challenger) needs no GPU: check it
    against the upstream known-answer vectors and against the oracle on edge and random states."""
    import json
    import random

    import numpy as np

    import eth_tx_proof_b200 as etp
    import oracle

    L = etp.load_library()
    fn = L.etp_host_poseidon_permute
    fn.argtypes = [ctypes.POINTER(ctypes.c_uint64)]
    fn.restype = None

    def perm(st):
        a = (ctypes.c_uint64 * 12)(*[int(x) for x in st])
        fn(a)
        return [int(x) for x in a]

    p = 0xFFFFFFFF00000001
    with open(os.path.join(ROOT, "tests", "golden", "poseidon_kat.json")) as f:
        kat = json.load(f)
    inputs = {"zeros": [0] * 12, "range12": list(range(12)), "neg_one": [p - 1] * 12}
    for v in kat["permutation"]:
        assert perm(inputs[v["input"]]) == [int(x, 16) for x in v["output"]]
    rng = random.Random(5)
    for st in [[2**64 - 1] * 12, [p] * 12, [2**32 - 1] * 12, [0xFFFFFFFF00000000] * 12] + [[rng.randrange(2**64) for _ in range(12)] for _ in range(300)]:
        assert perm(st) == [int(x) for x in oracle.poseidon_permute_naive(np.array(st, dtype=np.uint64))]


def test_challenger_c_abi_matches_the_oracle_and_compacts():
    """etp_challenger (plain data, host only): observe / get_challenge / compact == plonky2's Challenger as restated by the oracle."""
    import numpy as np

    import eth_tx_proof_b200 as etp
    import oracle

    g, o = etp.Challenger(), oracle.HostChallenger()
    rng = np.random.default_rng(5)
    for step in range(40):
        k = int(rng.integers(0, 20))
        vals = rng.integers(0, 2**63, size=k, dtype=np.uint64)
        g.observe(vals)
        o.observe(vals)
        m = int(rng.integers(0, 11))
        assert (g.get_n_challenges(m) == o.get_n(m)).all()
        if step % 7 == 3:
            assert (g.compact() == o.compact()).all()
        assert (g.words() == o.words()).all()
    c = g.clone()
    assert c.get_challenge() == g.get_challenge()


def test_fri_params_standard_configs():
    import eth_tx_proof_b200 as etp

    fp = etp.FriParams.make(22)  # StarkConfig::standard_fast_config
    assert (fp.rate_bits, fp.cap_height, fp.proof_of_work_bits, fp.num_query_rounds, fp.n_reductions) == (1, 4, 16, 84, 4)
    assert [fp.reduction_arity_bits[i] for i in range(4)] == [4, 4, 4, 4]
    fr = etp.FriParams.make(12, 3, 4, 16, 28)  # CircuitConfig::standard_recursion_config
    assert fr.n_reductions == 2
    assert etp.FriParams.make(5).n_reductions == 0 and etp.FriParams.make(6).n_reductions == 0 and etp.FriParams.make(8).n_reductions == 1


def test_cubin_disk_cache_round_trip(tmp_path):
    """ETP_CUBIN_CACHE: a compiled program is stored on disk and a FRESH process loads it instead of compiling (the analogue of
    the reference's persisted prover state); a corrupted file is ignored and rewritten."""
    import subprocess
    import time

    code = (
        "import sys, time, ctypes as C, numpy as np; sys.path.insert(0, %r)\n"
        "import eth_tx_proof_b200 as etp\n"
        "from eth_tx_proof_b200 import cprog\n"
        "L = etp.load_library(); w = np.ascontiguousarray(cprog.shape_program(300, 8).words)\n"
        "size = C.c_size_t(); err = C.create_string_buffer(256); t0 = time.perf_counter()\n"
        "rc = L.etp_cprog_compile_check(w.ctypes.data_as(C.POINTER(C.c_uint64)), w.size, C.byref(size), err, 256)\n"
        "print(rc, size.value, time.perf_counter() - t0)\n") % ROOT
    env = dict(os.environ, ETP_CUBIN_CACHE=str(tmp_path))

    def run():
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        rc, size, secs = r.stdout.split()
        assert rc == "0"
        return int(size), float(secs)

    size1, cold = run()
    files = list(tmp_path.glob("etp_*.cubin"))
    assert len(files) == 1 and files[0].stat().st_size == size1 + 24
    size2, warm = run()
    assert size2 == size1 and warm < cold / 3, (cold, warm)
    files[0].write_bytes(files[0].read_bytes()[:-7])  # truncated: must be ignored, recompiled and replaced
    size3, again = run()
    assert size3 == size1 and again > warm and files[0].stat().st_size == size1 + 24
