"""The committed bench records (profiles/) carry every key of the bench.py JSON contract, for our arm and for the
reference arm.  Guards the contract without a GPU; the numbers themselves come from B200 runs."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "e2e"]


def _load(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads([line for line in f if line.startswith("{")][-1])


def test_our_arm_record_has_the_contract_keys():
    for name, n in (("r1_bench_n1.json", 1), ("r1_bench_n2.json", 2), ("r1_bench_n8.json", 8)):
        d = _load(name)
        for k in BASE + ["gpu_launches", "clocks", "roofline"]:
            assert k in d, (name, k)
        assert d["n_gpus"] == n and d["dtype"] == "f64" and d["scaling"] == "weak" and d["vs_baseline"] is None
        assert "workload" in d["config"] and "model" not in d["config"]
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["value"] != d["value"]
        r = d["roofline"]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm"
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
        assert d["gpu_launches"] > 0
        assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"])
        if n == 1:
            c = d["cpu_baseline"]
            assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] in ("port", "reference")
        # whole-job throughput: elements of all ranks per step time
        assert abs(d["value"] - d["config"]["elements"] / d["ms_per_step"] / 1e3) < 1e-6 * d["value"]


def test_reference_arm_record():
    d = _load("r1_bench_reference_arm.json")
    for k in BASE + ["impl", "cpu_baseline"]:
        assert k in d, k
    assert d["impl"] == "reference"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] == d["cpu_baseline"]["value"]
    assert d["metric"] == _load("r1_bench_n1.json")["metric"] and d["unit"] == "Melem/s"
