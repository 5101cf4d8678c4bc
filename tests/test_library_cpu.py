"""CPU tests of the product library's host side: the C-ABI shared library loads and exports every
symbol include/etp_b200.h declares, refuses to run without a GPU (no CPU fallback), and carries the
pinned Poseidon round constants.  No compute calls are made here."""
import ctypes
import hashlib
import os
import re
import struct

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
    f64 = src.split("#define ETP_POSEIDON_RC_F64_TABLE")[1].split("}")[0]
    bits = [int(x, 16) for x in re.findall(r"0x([0-9a-f]{16})ULL", f64)]
    assert len(bits) == 720
    as_f64 = lambda b: struct.unpack("<d", struct.pack("<Q", b))[0]
    for r in range(30):
        for h in range(2):
            ks = [((vals[12 * (r + 1) + i] >> (32 * h)) & 0xFFFFFFFF) if r < 29 else 0 for i in range(12)]
            for i in range(6):
                assert as_f64(bits[24 * r + 12 * h + i]) == float(2**51 + ks[i])
                assert as_f64(bits[24 * r + 12 * h + 6 + i]) == float(2**52 + ks[i + 6] - ks[i])


def test_synthetic_traces_are_deterministic_and_valid():
    import numpy as np

    import oracle
    from eth_tx_proof_b200 import synthetic as syn

    a, b = syn.memory_trace(9, seed=3), syn.memory_trace(9, seed=3)
    assert (a == b).all() and a.shape == (21, 512)
    assert oracle.check_constraints(oracle.TABLE_MEMORY, a) == -1
    c = syn.random_columns(3, 8)
    assert (c < np.uint64(0xFFFFFFFF00000001)).all()
