"""CPU: the bench.py contract that can be checked without a GPU -- the reference arm's JSON line (it times the
reference's algorithm on the host cores) and the loud failure of the product arm when there is no CUDA device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT,
                          timeout=300)


def test_reference_arm_prints_one_contract_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--batch", "4", "--persons", "5")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["metric"].startswith("grouping images/sec") and d["unit"] == "images/s"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] >= 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["batch_per_gpu"] == 4 and d["config"]["persons"] == 5 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0  # nothing of ours runs in the reference arm


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    r = _run("--steps", "1", "--warmup", "1", "--batch", "4")
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
