"""GPU: bench.py keeps the driver's contract (one JSON line, the required keys, both arms)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import REPO

pytestmark = pytest.mark.gpu


def _run(*args):
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), *args], capture_output=True, text=True, timeout=600, cwd=REPO)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


def test_b200_arm_contract():
    d = _run("--steps", "2", "--warmup", "3", "--size", "128", "--spp", "16")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["unit"] == "Mrays/s" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["value"] > 0 and d["gpu_launches"] == 4
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and d["e2e"]["d2h_bytes_per_step"] == 128 * 128 * 3 * 8
    r = d["roofline"]
    assert set(r) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and 0 < r["frac"] < 1.5 and r["unit"] == "TFLOP/s"
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
    assert d["value"] / c["value"] > 10          # sanity: the GPU arm is not a CPU fallback


def test_reference_arm_contract():
    d = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--size", "128", "--spp", "16")
    assert d["impl"] == "reference" and d["gpu_launches"] == 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["metric"].startswith("Mrays/s")
