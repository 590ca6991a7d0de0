"""The bench.py contract (one JSON line on stdout with the keys the driver reads), checked on both arms:
the reference arm runs on the CPU (oracle port), our arm needs the GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=300):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout,
                       cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must carry exactly one JSON line"
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert d["impl"] == "reference" and d["metric"] == "NTT polymul/s" and d["unit"] == "polymul/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "polymul/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "N=1024" in d["config"]["workload"] and "65537" in d["config"]["workload"]


@pytest.mark.gpu
def test_our_arm_line():
    d = _run("--steps", "3", "--warmup", "3", "--no-extras")
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    assert d["metric"] == "NTT polymul/s" and d["n_gpus"] == 1 and d["steps"] == 3 and d["scaling"] == "weak"
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0.3 < r["frac"] < 1.05 and r["traffic"] is None or r["traffic"] > 1e9
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["matches_device_result"]
    assert e["value"] < d["value"]  # host buffers and PCIe are inside the timed region
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and c["gpu_matches_on_sample"]
    assert d["gpu_launches"] == 3 and d["clocks"]["sm_mhz"] > 0
    assert d["bootstrap"]["value"] > 0 and d["bootstrap"]["e2e"]["matches_device_result"]
    # round 2: the keys the driver keeps carry the other formats and the measured peaks
    assert "17 bits" in e["call"] and e["u32_wire"]["matches_device_result"] and e["u64_wire"]["matches_device_result"]
    assert e["value"] > e["u32_wire"]["value"] > e["u64_wire"]["value"]  # PCIe-bound: fewer wire bytes, more polymul/s
    assert e["pcie_microbench_gbs_per_rank"]["h2d_gbs"][0] > 1
    assert r["u32_device_format"]["matches_u64_result"] and r["u32_device_format"]["value"] > 0
    assert 0.3 < r["int"]["frac"] < 1.2
    b = d["bootstrap"]["roofline"]
    assert b["peak_source"].startswith("measured") and 0.3 < b["frac"] < 1.05
