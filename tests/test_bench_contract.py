"""bench.py's driver contract on the CPU: the reference arm (the oracle port timed on the host cores) runs without a GPU and
prints ONE JSON line with the keys the driver reads; the GPU arm fails loudly when there is no device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    r = _run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--rows", "300000"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["unit"] == "rows/s"
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["vs_baseline"] is None           # BASELINE.md publishes no number for this metric


def test_q3_and_q95_reference_arms_answer():
    for wl in ("q3", "q95"):
        r = _run(["--impl", "reference", "--workload", wl])
        assert r.returncode == 0
        d = json.loads(r.stdout.strip().splitlines()[-1])
        assert d["impl"] == "reference" and "unavailable" in d


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = _run(["--steps", "1", "--warmup", "1", "--rows", "100000", "--no-e2e", "--no-cpu-baseline"])
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)       # no CPU fallback behind the product arm
