"""e2e experiment: the fused fragment on pinned host columns, pushed as ONE batch vs many small batches (what the C++ operator
path does).  python tools/e2e_batches.py [--rows 600000000] [--batch 4194304]"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from starrocks_b200 import abi, gpu, ssb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=600_000_000)
    ap.add_argument("--batch", type=int, default=1 << 22)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = gpu.Context(0, stream=stream.cuda_stream)
    sz = ssb.sizes(100.0)
    dims = ssb.gen_dims(100.0)
    gjoins, gkeep = ssb.build_dims(gpu, dims, ssb.dim_plans_q41(), ctx=ctx)
    n = a.rows
    cols = bench.gen_lineorder_device(torch, dev, n, sz, ssb.SEED)
    host = {nm: torch.empty(n, dtype=torch.int32, pin_memory=True) for nm in ssb.Q41_FACT_COLS}
    for nm in ssb.Q41_FACT_COLS:
        host[nm].copy_(cols[nm])
    torch.cuda.synchronize()
    frag = gpu.Fragment(ctx, abi.ScanDesc(), gjoins, ssb.q41_agg_desc())
    out = {}
    for label, bs in (("one_batch", n), ("batches", a.batch), ("batches_x4", a.batch * 4), ("batches_x16", a.batch * 16)):
        for rep in range(2):
            frag.reset()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record(stream)
            for lo in range(0, n, bs):
                hi = min(n, lo + bs)
                ch = abi.Chunk([(ssb.LO_SLOTS[nm], host[nm][lo:hi].data_ptr(), None, abi.TYPE_INT) for nm in ssb.Q41_FACT_COLS], num_rows=hi - lo,
                               mem=abi.MEM_HOST_PINNED)
                frag.push(ch)
            t1 = time.perf_counter()
            e1.record(stream)
            torch.cuda.synchronize()
            out[label] = {"batch_rows": bs, "device_ms": e0.elapsed_time(e1), "enqueue_ms": 1e3 * (t1 - t0), "rows_passed": frag.rows_passed}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
