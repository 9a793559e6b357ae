"""The driver-facing contract of bench.py that can be checked without a GPU: the reference arm prints one JSON line with the agreed
keys (it times the CPU port of the reference path: the one place besides tests/ and smoke() that may execute oracle/), and the
product arm fails loudly instead of falling back to a CPU path when there is no CUDA device."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=240):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-streams", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"] == "audio-sec/sec" and d["unit"] == "audio-s/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["config"]["workload"].startswith("configs[3]")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None                       # BASELINE.md publishes no number for this metric


def test_product_arm_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _run("--steps", "1", "--warmup", "1", "--no-sweep", "--no-cpu-baseline", timeout=120)
    assert r.returncode != 0                              # no CPU fallback: the product path needs the CUDA library and a device
    assert not any(ln.strip().startswith("{") for ln in r.stdout.splitlines()), "no result line may be printed without a GPU"
