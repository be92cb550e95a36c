"""The C++ host layer (starrocks_b200/host): GPU operators behind the reference's pipeline::Operator interface.
CPU: it compiles and links against the C-ABI library, and fails loudly without a device.  GPU: the pipeline test binary
drives scan -> build / probe x4 -> aggregate (per-operator and fused) with the PipelineDriver protocol and checks the
groups against a row-at-a-time evaluation."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "starrocks_b200", "host")
BIN = os.path.join(HOST, "tests", "pipeline_q41_test")


def _build(gpu):
    gpu.lib()
    subprocess.check_call(["make", "-C", HOST, "-s"])
    assert os.path.exists(BIN)


def test_host_layer_builds_and_fails_loudly_without_a_device(gpu):
    import torch
    _build(gpu)
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([BIN, "1000"], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_pipeline_q41_through_the_operator_interface(gpu):
    _build(gpu)
    r = subprocess.run([BIN, "1500000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "PIPELINE_Q41_OK" in r.stdout
