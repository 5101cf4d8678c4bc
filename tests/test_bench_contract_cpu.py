"""The bench line contract, checked on the committed end-of-round line (profiles/r01_bench_2p22x128_v10.json) and on the
reference-arm line: every key the driver reads is present and of the right kind, the numbers are consistent with each
other, and bench.py still parses.  No GPU needed."""
import ast
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        lines = [l for l in f.read().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_bench_line_contract():
    d = _load("r01_bench_2p22x128_v10.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "commit_hbm_gbs" and d["unit"] == "GB/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["dtype"] == "u64" and d["data"] == "synthetic" and d["scaling"] == "weak" and d["warmup"] >= 3
    assert "workload" in d["config"] and "2^22 x 128" in d["config"]["workload"] and "model" not in d["config"]
    # value = algorithmic bytes / device time
    assert abs(d["value"] - d["config"]["algorithmic_bytes_per_step"] / (d["ms_per_step"] / 1e3) / 1e9) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 8 * 128 * (1 << 22) and e["d2h_bytes_per_step"] == 512
    assert e["value"] < d["value"]  # host copies are inside the e2e region
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert d["gpu_launches"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    for bad in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"):
        assert bad not in d["clocks"]["reasons"]
    assert d["stark"]["prove_ms"] > 0 and d["stark"]["prove_host_ms"] > d["stark"]["prove_ms"] and d["tx"]["tx_per_min"] > 0


def test_reference_arm_line_contract():
    d = _load("r01_bench_reference_arm_v10.json")
    assert d["impl"] == "reference" and d["metric"] == "commit_hbm_gbs" and d["unit"] == "GB/s"
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_bench_script_parses_and_defaults():
    src = open(os.path.join(ROOT, "bench.py")).read()
    ast.parse(src)
    assert '"--impl"' in src and '"--gpus"' in src and '"--steps"' in src and '"--warmup"' in src


def test_final_round2_line_carries_the_real_recursion_leg():
    """profiles/r02_bench_2p22x128_v10_final.json (the default `python bench.py` at the end of round 2): the base contract keys,
    and the tx_real_recursion leg — 7 table STARKs + 15 circuit proofs per transaction, every circuit over the real proof below it —
    with its per-kind times adding up to less than the whole job, the in-run CPU parity figure, and the real block of 8."""
    d = _load("r02_bench_2p22x128_v10_final.json")
    assert d["metric"] == "commit_hbm_gbs" and d["n_gpus"] == 1 and d["warmup"] >= 3 and d["e2e"]["value"] < d["value"]
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
    r = d["tx_real_recursion"]
    assert "error" not in r and r["circuit_proofs_per_tx"] == 15 and len(r["circuits"]) == 15
    kinds = [c["kind"] for c in r["circuits"]]
    assert kinds.count("wrapper") == 7 and kinds.count("shrink") == 7 and kinds[-1] == "root"
    assert all(c["degree_bits"] == 13 for c in r["circuits"] if c["kind"] == "shrink")
    per_kind = sum(v["sum"] for v in r["circuit_prove_ms"].values())
    assert 0 < per_kind < r["tx_ms"] < 1000 and r["tx_per_min"] > 60000.0 / r["tx_ms"] * 0.9
    assert r["cpu_baseline"]["parity"].startswith("FRI proof words ==") and r["cpu_baseline"]["cpu_ms"] > 50 * r["cpu_baseline"]["gpu_ms"]
    assert [c["name"] for c in r["block_of_8_segments"]["circuits"]] == ["agg1", "agg2", "agg3", "block"]
    assert r["block_of_8_segments"]["ms"] > 8 * r["tx_ms"] * 0.5
