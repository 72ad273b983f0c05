"""bench.py contract checks that need no GPU: the reference arm (the oracle timed on the host cores) prints one JSON line
with the keys the driver reads, for the small workloads; the GPU arm refuses to run without a CUDA device (no CPU path)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)


@pytest.mark.parametrize("workload", ["S2", "S1"])
def test_reference_arm_prints_the_contract_line(workload):
    r = _run("--impl", "reference", "--workload", workload, "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith(workload)


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _run("--workload", "S2", "--steps", "1")
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout) or "CUDA" in (r.stderr + r.stdout)
