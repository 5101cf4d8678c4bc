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
