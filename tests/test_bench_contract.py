"""bench.py's CPU-only legs: the reference arm prints one JSON line with the contract's keys, refuses to pretend without a
GPU, and the library baseline (OpenSSL) loads when it was built."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    env = dict(os.environ, OMP_NUM_THREADS=str(min(8, os.cpu_count() or 1)))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


def test_reference_arm_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-ranges", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "headers/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("headers/sec") and line["value"] > 0 and line["n_gpus"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "headers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _run("--steps", "1", "--warmup", "0", timeout=300)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)


def test_cpu_library_baseline_loads():
    sys.path.insert(0, ROOT)
    import bench
    if not os.path.exists(os.path.join(ROOT, "baseline", "_cpulib", "libbsx_cpulib.so")):
        pytest.skip("baseline/cpu_library.c not built (no OpenSSL development files)")
    b = bench.cpu_library_baseline()
    assert b and b["kind"] == "library" and b["unit"] == "headers/s" and b["value"] > 1000 and b["cores"] >= 1
