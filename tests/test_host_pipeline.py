"""The C++ host layer (starrocks_b200/host): GPU operators behind the reference's pipeline::Operator interface.
CPU: it compiles and links against the C-ABI library, and fails loudly without a device.  GPU: the pipeline test binary
(tests/cpp) drives scan -> build / probe x4 -> aggregate (per-operator, fused, with runtime filters), the exchange sink and
the two-phase aggregate with the PipelineDriver protocol and checks the groups against the CPU oracle; the operator
end-to-end benchmark (starrocks_b200/host/bench) runs DOP pipeline drivers on host threads into one shared fragment."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "starrocks_b200", "host")
CPP = os.path.join(ROOT, "tests", "cpp")
BIN = os.path.join(CPP, "pipeline_q41_test")
BENCH = os.path.join(HOST, "bench", "operator_e2e_bench")


def _build(gpu, oracle):
    gpu.lib()
    oracle.lib()
    subprocess.check_call(["make", "-C", HOST, "-s"])
    subprocess.check_call(["make", "-C", CPP, "-s"])
    assert os.path.exists(BIN) and os.path.exists(BENCH)


def test_host_layer_builds_and_fails_loudly_without_a_device(gpu, oracle):
    import torch
    _build(gpu, oracle)
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    for binary in (BIN, BENCH):
        r = subprocess.run([binary, "1000"], capture_output=True, text=True)
        assert r.returncode == 3 and "no CPU fallback" in r.stderr


def test_binary_column_and_dictionary_path_on_the_host():
    """the host string column (BinaryColumn) and the low-cardinality dictionary path that keeps strings out
    of the GPU operators -- local segment codes -> global ids (TYPE_INT, what the C-ABI sees) -> DictDecodeOperator.  The
    binary reproduces the reference's binary_column_test.cpp cases (incl. the xor_checksum golden) and needs no GPU."""
    subprocess.check_call(["make", "-C", CPP, "-s", "binary_column_test"])
    r = subprocess.run([os.path.join(CPP, "binary_column_test")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "BINARY_COLUMN_TEST_OK" in r.stdout


@pytest.mark.gpu
def test_pipeline_q41_through_the_operator_interface(gpu, oracle):
    _build(gpu, oracle)
    r = subprocess.run([BIN, "1500000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "PIPELINE_Q41_OK" in r.stdout
    assert "BuildHashTableTime" in r.stdout and "SearchHashTableTime" in r.stdout and "AggComputeTime" in r.stdout   # _unique_metrics


@pytest.mark.gpu
def test_operator_e2e_bench_multi_driver(gpu, oracle):
    # 4 pipeline drivers on 4 host threads push 4096-row chunks into one shared fragment through asynchronous pinned
    # batches (small batches so that every driver cycles through both of its buffers many times)
    _build(gpu, oracle)
    r = subprocess.run([BENCH, "6000000", "4", "65536"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["matches_row_at_a_time_evaluation"] is True and line["groups"] == 35 and line["fragment_batches"] >= 90
