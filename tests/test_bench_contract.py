"""bench.py's reference arm (`--impl reference`: the CPU port of the reference's algorithm on the host cores) runs here
without a GPU; its JSON line must carry the contract's keys, on BASELINE.json's metric. The GPU arm's line is checked
on the GPU box (tests/test_gpu_pipeline.py) and by the driver."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_contract():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [l for l in p.stdout.strip().splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["impl"] == "reference"
    assert d["metric"].replace("x", "×") == base["metric"] or d["metric"] == base["metric"].replace("×", "x")
    assert d["value"] > 1e6 and d["unit"] == "steps/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["ms_per_step"] > 0
    assert d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic" and d["scaling"] in ("strong", "weak")
    assert "3840x2160" in d["config"]["workload"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_own_arm_fails_loudly_without_a_gpu():
    """No CPU fallback: on a box without an sm_100 device the product arm must exit non-zero, not fall back to the oracle."""
    import ctypes as C
    sys.path.insert(0, os.path.join(ROOT, "blackhole-simulation_b200"))
    import gravitas_b200 as g
    n = C.c_int32(0)
    g.lib().gvt_device_count(C.byref(n))
    if n.value > 0:
        import pytest
        pytest.skip("a GPU is present")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--no-cpu"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert p.returncode != 0 and "GravitasError" in (p.stderr + p.stdout)
    assert not any(l.startswith('{"metric"') for l in p.stdout.splitlines())
