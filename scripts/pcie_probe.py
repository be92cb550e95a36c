"""Scratch measurement (not part of the product or the tests): how fast can kernels read pinned host memory in place,
and what do the fragment passes cost when the fact columns live in pinned host memory?
  python scripts/pcie_probe.py [rows]
"""
import sys
import time

import torch

sys.path.insert(0, ".")
from starrocks_b200 import abi, gpu, ssb  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 600_000_000
    sf = 100.0
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = gpu.Context(0, stream=stream.cuda_stream)
    quick = len(sys.argv) > 2 and sys.argv[2] == "quick"
    # dense zero-copy read bandwidth
    buf = torch.empty(1 << 30 if not quick else 1 << 20, dtype=torch.int32, pin_memory=True)
    buf.fill_(1)
    for rep in range(3 if not quick else 0):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.bandwidth_probe(buf.data_ptr(), buf.numel() * 4)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"zero-copy dense read: {buf.numel() * 4 / ms / 1e6:.1f} GB/s ({ms:.1f} ms)")
    d = torch.empty(1 << 30, dtype=torch.int32, device=dev)
    for rep in range(2 if not quick else 0):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        d.copy_(buf, non_blocking=True)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"cudaMemcpyAsync H2D: {buf.numel() * 4 / ms / 1e6:.1f} GB/s ({ms:.1f} ms)")
    del d, buf

    sz = ssb.sizes(sf)
    dims = ssb.gen_dims(sf)
    gjoins, gkeep = ssb.build_dims(gpu, dims, ssb.dim_plans_q41(), ctx=ctx)
    import bench
    cols = bench.gen_lineorder_device(torch, dev, n, sz, 7)
    host_cols = {nm: torch.empty(n, dtype=torch.int32, pin_memory=True) for nm in ssb.Q41_FACT_COLS}
    for nm in ssb.Q41_FACT_COLS:
        host_cols[nm].copy_(cols[nm])
    torch.cuda.synchronize()
    for mode in ((2, 1) if not quick else (2,)):
        frag = gpu.Fragment(ctx, abi.ScanDesc(), gjoins, ssb.q41_agg_desc(), mode=mode)
        dchunk = ssb.fact_chunk(cols, ssb.Q41_FACT_COLS, mem=abi.MEM_DEVICE)
        frag.push(dchunk)   # plan on device data
        ctx.sync()
        hchunk = abi.Chunk([(ssb.LO_SLOTS[nm], host_cols[nm].data_ptr(), None, abi.TYPE_INT) for nm in ssb.Q41_FACT_COLS],
                           num_rows=n, mem=abi.MEM_HOST_PINNED)
        for rep in range(3 if not quick else 1):
            frag.reset()
            t0 = time.perf_counter()
            frag.push(hchunk)
            ctx.sync()
            dt = (time.perf_counter() - t0) * 1e3
            print(f"mode {mode} in-place push: {dt:.1f} ms, passes {frag.last_pass_ms()}")
        frag.close()


if __name__ == "__main__":
    main()
