"""The drop-in boundary from a non-Python host: tests/cabi/host_only.c is compiled as plain C11 against include/etp_b200.h,
linked with libetp_b200.so and run (host-only entry points: no GPU needed).  What a cgo / Rust-FFI binding relies on — the header
parses as C, the symbols link, the plain-data structs have the declared layout, and nothing falls back to the CPU without a
device (the Rust binding itself is rust/etp_b200_sys, which cannot be compiled in this image)."""
import json
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_plain_c_host_links_and_runs(tmp_path):
    import eth_tx_proof_b200 as etp

    etp.load_library()  # builds nothing: the in-tree .so must already exist
    lib_dir = os.path.join(ROOT, "eth_tx_proof_b200")
    exe = str(tmp_path / "host_only")
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cabi", "host_only.c"), "-o", exe, "-L", lib_dir, "-letp_b200", f"-Wl,-rpath,{lib_dir}"],
                   check=True, capture_output=True, text=True)
    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "poseidon_kat.json")))["permutation"][0]
    assert kat["input"] == "zeros"
    r = subprocess.run([exe] + kat["output"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "all C-host checks passed" in r.stdout, r.stdout + r.stderr
    assert r.stdout.count("ok   ") == 8
