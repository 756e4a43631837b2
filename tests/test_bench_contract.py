"""The reference arm of bench.py runs here (it is the CPU path): one JSON line with the keys the
driver reads, and the same `config` object the B200 arm prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--ref-games", "100"], capture_output=True, text=True, timeout=600,
                         cwd=ROOT, check=True).stdout
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1 and d["higher_is_better"] is True
    assert d["unit"] == "env-steps/s" and d["value"] > 0 and d["vs_baseline"] is None and d["scaling"] == "weak"
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.bench_config(1 << 24, 1)        # what the B200 arm prints for its default run
    assert d["metric"] == bench.METRIC if hasattr(bench, "METRIC") else True


import pytest


@pytest.mark.gpu
def test_b200_arm_prints_the_contract_line():
    """bench.py at a reduced size on the GPU: one JSON line with every key of the contract."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--envs", str(1 << 20), "--steps", "1",
                          "--warmup", "3", "--passes-per-step", "2", "--cpu-seconds", "0.3", "--sweep-games", "1000000",
                          "--e2e-steps", "1"], capture_output=True, text=True, timeout=900, cwd=ROOT, check=True).stdout
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "clocks", "gpu_launches"):
        assert key in d, key
    assert "impl" not in d or d["impl"] == "b200"
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and 0 < r["frac"] <= 1.0
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["gpu_launches"] > 0 and d["clocks"]["sm_mhz"] > 0
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.bench_config(1 << 20, 1)
