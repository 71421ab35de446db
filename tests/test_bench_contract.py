"""CPU: the reference arm of bench.py (the one arm that runs without a GPU) prints ONE JSON line with the contract keys,
and the product arm refuses to run without CUDA instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600,
                          cwd=ROOT)


def test_reference_arm_json_line():
    r = _run("--impl", "reference", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["impl"] == "reference" and d["metric"] == base["metric"] and d["unit"] == "images/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["scaling"] == "strong" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert "workload" in d["config"] and "model" not in d["config"]


def test_product_arm_fails_loudly_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        import pytest

        pytest.skip("a GPU is present")
    r = _run("--steps", "1", "--warmup", "1")
    assert r.returncode != 0
    assert not [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
