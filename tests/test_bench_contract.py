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


def test_reference_arm_loads_nothing_of_the_product_and_reports_its_cpu_budget():
    """the reference arm used to dlopen libspgroup.so (through build()) and sized its pool with
    os.cpu_count(), which ignores the cgroup quota / affinity of the lease."""
    code = (
        "import runpy, sys, os\n"
        f"sys.argv = [{os.path.join(ROOT, 'bench.py')!r}, '--impl', 'reference', '--steps', '1', '--warmup', '1', '--batch', '4', '--persons', '5']\n"
        f"runpy.run_path({os.path.join(ROOT, 'bench.py')!r}, run_name='__main__')\n"
        "mods = [m for m in sys.modules if m in ('improved_body_parts_b200.grouping', 'improved_body_parts_b200.dropin', '__graft_entry__')]\n"
        "maps = open('/proc/self/maps').read()\n"
        "print('CHECK', mods, maps.count('libspgroup'))\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    check = next(l for l in r.stdout.splitlines() if l.startswith("CHECK"))
    assert check == "CHECK [] 0", check  # torch (and its own CUDA runtime) may be there for get_num_threads(); nothing of ours
    d = json.loads(next(l for l in r.stdout.splitlines() if l.startswith("{")))
    cb = d["cpu_baseline"]
    assert cb["cores"] == cb["cpus"]["usable"] <= cb["cpus"]["affinity"] and cb["single_process"]["value"] > 0
    assert cb["cpus"]["usable"] == len(os.sched_getaffinity(0)) or cb["cpus"]["cgroup_quota_cpus"] is not None
