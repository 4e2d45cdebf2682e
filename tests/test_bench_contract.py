"""The bench.py JSON line committed with the newest capture carries every key the measurement contract names
(checked on the committed artefact, so it runs without a GPU), and the reference arm's line does too."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _newest_bench():
    paths = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*", "z_bench.json")))
    assert paths, "no committed bench line under profiles/"
    with open(paths[-1]) as f:
        return json.load(f)


def test_committed_bench_line_has_the_contract_keys():
    d = _newest_bench()
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "collision_checked_edges_per_sec" and d["higher_is_better"] is True
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"] and "l2" in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    for k in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert k in d["clocks"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert 0 < e["value"] < d["value"]                       # host buffers in and out can only be slower
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
    assert r["traffic"] is None or r["traffic"] >= 0.9 * r["algorithmic_bytes_per_launch"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["unit"] == d["unit"] and c["sample"]
    # value is what the step time says
    assert abs(d["value"] - d["edges_per_step"] / (d["ms_per_step"] / 1e3)) / d["value"] < 1e-9


def test_bench_argument_defaults_are_single_gpu_and_short():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True)
    assert out.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert flag in out.stdout
