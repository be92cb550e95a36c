"""Scratch measurement (not part of the product): where does the N>1 tail of a bench step spend its time?
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/nccl_probe.py
"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from starrocks_b200 import abi, gpu, ssb  # noqa: E402
from starrocks_b200.distributed import all_reduce_dense_state, dense_state_views  # noqa: E402


def timeit(name, fn, rank, iters=200, sync=True):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
        if sync:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / iters * 1e6
    if rank == 0:
        print(f"{name}: {dt:.1f} us/iter", flush=True)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = gpu.Context(local, stream=stream.cuda_stream)
    t = torch.zeros(525, dtype=torch.int64, device=dev)
    timeit("all_reduce 525 x int64 (sync each)", lambda: dist.all_reduce(t), rank)
    timeit("all_reduce 525 x int64 (async, 200 back to back)", lambda: dist.all_reduce(t), rank, sync=False)
    a, b, c = (torch.zeros(175, dtype=torch.int64, device=dev) for _ in range(3))

    def cat_path():
        flat = torch.cat([a, b, c])
        dist.all_reduce(flat)
        a.copy_(flat[:175]); b.copy_(flat[175:350]); c.copy_(flat[350:])
    timeit("cat + all_reduce + 3 copies", cat_path, rank)

    def three():
        dist.all_reduce(a); dist.all_reduce(b); dist.all_reduce(c)
    timeit("3 separate all_reduce", three, rank)

    # the real thing on a tiny fragment
    sf = 1.0
    dims = ssb.gen_dims(sf)
    gjoins, gkeep = ssb.build_dims(gpu, dims, ssb.dim_plans_q41(), ctx=ctx)
    lo = ssb.gen_lineorder(sf, n=1_000_000)
    cols = {k: torch.from_numpy(v).to(dev) for k, v in lo.items()}
    chunk = ssb.fact_chunk(cols, ssb.Q41_FACT_COLS, mem=abi.MEM_DEVICE)
    frag = gpu.Fragment(ctx, abi.ScanDesc(), gjoins, ssb.q41_agg_desc())
    frag.push(chunk)
    ctx.sync()
    timeit("agg.dense_state()", lambda: frag.agg.dense_state(), rank, sync=False)
    st = frag.agg.dense_state()
    timeit("dense_state_views()", lambda: dense_state_views(st, dev), rank, sync=False)
    timeit("all_reduce_dense_state()", lambda: all_reduce_dense_state(frag.agg.dense_state(), dev), rank)

    def full_step():
        frag.reset()
        frag.push(chunk)
        all_reduce_dense_state(frag.agg.dense_state(), dev)
        frag.agg.finish()
        if rank == 0:
            gpu.chunk_out_to_host(ctx, frag.agg.pull(mem=abi.MEM_HOST))
    timeit("full step on 1 M rows, allreduce merge", full_step, rank)

    def local_step():
        frag.reset()
        frag.push(chunk)
        frag.agg.finish()
        gpu.chunk_out_to_host(ctx, frag.agg.pull(mem=abi.MEM_HOST))
    timeit("full step on 1 M rows, no merge (every rank pulls)", local_step, rank)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
