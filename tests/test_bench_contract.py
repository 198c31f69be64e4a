"""The parts of bench.py's contract that need no GPU: the reference arm prints exactly one JSON line with the agreed keys
(and runs the reference's own C/OpenMP functions when oracle/_ref is built), other ranks of a torchrun launch print
nothing, and the GPU arm refuses to run without CUDA instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=900,
                          env=dict(os.environ, **(env or {})))


def test_reference_arm_prints_one_json_line():
    r = _bench("--impl", "reference", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mcell-updates/s" and d["higher_is_better"] is True
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config",
                "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["value"] > 0 and d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "workload" in d["config"]
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_fd2d_3_2.so")):
        assert cb["kind"] == "reference" and "test_3_2.c" in cb["sample"]


def test_reference_arm_other_ranks_stay_silent():
    r = _bench("--impl", "reference", "--gpus", "2", "--steps", "1", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _bench("--steps", "1", "--no-cpu", "--no-e2e")
    assert r.returncode != 0 and r.stdout.strip() == "" and "CUDA" in r.stderr
