"""The committed bench records (profiles/) carry every key of the bench.py JSON contract, for our arm and for the
reference arm.  Guards the contract without a GPU; the numbers themselves come from B200 runs."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "e2e"]


def _load(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads([line for line in f if line.startswith("{")][-1])


@pytest.mark.parametrize("name,n,scaling", [("r2_bench_n1.json", 1, "weak"), ("r2_bench_n2_c5.json", 2, "strong"),
                                            ("r2_bench_n8_c5.json", 8, "strong"), ("r2_bench_n8_c2w.json", 8, "weak")])
def test_our_arm_record_has_the_contract_keys(name, n, scaling):
    d = _load(name)
    for k in BASE + ["gpu_launches", "clocks", "roofline"]:
        assert k in d, (name, k)
    assert d["n_gpus"] == n and d["dtype"] == "f64" and d["scaling"] == scaling and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["value"] != d["value"]
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic", "fp64"} <= set(r) and r["bound"] == "hbm"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert d["gpu_launches"] > 0
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"])
    # whole-job throughput: elements of the whole mesh per step time
    assert abs(d["value"] - d["config"]["elements"] / d["ms_per_step"] / 1e3) < 1e-6 * d["value"]
    if n == 1:
        c = d["cpu_baseline"]
        assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] in ("port", "reference")
        assert d["config"]["workload"].startswith("C2") and d["parity"]["ok"] is True
        # the optimised CPU baseline (factored math, fused sweep, all cores) beside the faithful port
        assert d["cpu_baseline_opt"]["value"] > c["value"] and d["cpu_baseline_opt"]["kind"] == "port"
        assert d["c5_single_gpu"]["elements"] == 256**3
    else:
        assert ("C5" in d["config"]["workload"]) == (scaling == "strong")
        assert d["newton_step"]["pcg_iterations"] > 0


def test_reference_arm_record():
    d = _load("r2_bench_reference_arm.json")
    for k in BASE + ["impl", "cpu_baseline", "cpu_baseline_1thread"]:
        assert k in d, k
    assert d["impl"] == "reference"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] == d["cpu_baseline"]["value"]
    assert d["metric"] == _load("r2_bench_n1.json")["metric"] and d["unit"] == "Melem/s"
    # the whole C2 mesh per step on a stated number of threads, and the single-thread figure of the single-threaded reference
    assert d["config"]["sample_elements_per_step"] == 131072 and d["cpu_baseline"]["cores"] == d["config"]["threads"] > 1
    assert d["cpu_baseline_1thread"]["cores"] == 1
    assert d["cpu_baseline_opt"]["value"] > d["cpu_baseline"]["value"]
