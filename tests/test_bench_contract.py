"""CPU: the committed bench lines (profiles/r01_bench_*.json, produced by bench.py on B200s) carry every key the
measurement contract asks for, and the numbers in them are self-consistent (value = proofs / time, frac = achieved / peak)."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINES = sorted(glob.glob(os.path.join(ROOT, "profiles", "r01_bench_*.json")))


@pytest.mark.parametrize("path", LINES, ids=[os.path.basename(p) for p in LINES])
def test_bench_line_contract(path):
    d = json.load(open(path))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
        assert k in d, k
    assert d["metric"] == "groth16_proofs_per_sec" and d["unit"] == "proofs/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    cfg = d["config"]
    assert "workload" in cfg and "model" not in cfg
    # value is the whole-job aggregate: proofs of all ranks / device time of the timed steps
    proofs = d["n_gpus"] * cfg["proofs_per_step_per_gpu"] * d["steps"]
    assert d["value"] == pytest.approx(proofs / (d["ms_per_step"] * d["steps"] / 1e3), rel=1e-6)
    e = d["e2e"]
    assert e["unit"] == "proofs/s" and 0 < e["value"] <= d["value"] * 1.02
    assert e["h2d_bytes_per_step"] == cfg["proofs_per_step_per_gpu"] * cfg["n_vars"] * 32 and e["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9) and 0 < r["frac"] < 1
    c = d["clocks"]
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert c["sm_mhz"] >= 0.9 * c["sm_max_mhz"]
    if "cpu_baseline" in d:
        b = d["cpu_baseline"]
        assert b["kind"] in ("port", "reference") and b["cores"] >= 1 and b["value"] > 0 and b["sample"]
        assert d["proof_matches_cpu_port"] is True


def test_reference_arm_line():
    d = json.load(open(os.path.join(ROOT, "profiles", "r01_ref_n1.json")))
    assert d["impl"] == "reference" and d["metric"] == "groth16_proofs_per_sec" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] and d["cpu_baseline"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
