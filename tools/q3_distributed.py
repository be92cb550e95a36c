"""TPC-H Q3 across N GPUs of one box with an NCCL hash shuffle (BASELINE.json config 3).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/q3_distributed.py --sf 10

Plan per rank (one process per GPU; the fragment instances of the reference's distributed plan):
    customer (replicated, broadcast join)   -> J1
    orders shard  -- o_orderdate < D, SEMI J1 -> HASH_PARTITIONED exchange on o_orderkey  -> local build of J2
    lineitem shard -- l_shipdate > D          -> HASH_PARTITIONED exchange on l_orderkey  -> probe J2 + aggregate
The exchange is the reference's ExchangeSink hash step (FNV + ReduceOp, exchange_sink_operator.cpp:586-637) done by
sr_xchg_partition on the device, followed by all_to_all_single over NCCL (starrocks_b200.distributed).  Groups are
disjoint across ranks after the shuffle, so the result is the union of the per-rank results.
Rank 0 checks the union against the CPU oracle run on the whole tables (bit-exact), then prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from starrocks_b200 import abi, gpu, tpch  # noqa: E402
from starrocks_b200.distributed import device_view, exchange_partitions  # noqa: E402


def shuffle(ctx, xchg, chunk, dev, keep):
    """partition `chunk` (device) by the xchg's key and exchange -> (list of received tensors, rows sent off-rank bytes)"""
    out, offs = xchg.partition(chunk)
    cols = [device_view(out.cols[k].data, out.num_rows, abi.TYPE_WIDTH[out.cols[k].type], dev) for k in range(out.num_cols)]
    meta = [(out.cols[k].slot_id, out.cols[k].type) for k in range(out.num_cols)]
    recv = exchange_partitions(cols, offs.tolist())
    rank = dist.get_rank()
    sent_rows = int(out.num_rows - (offs[rank + 1] - offs[rank]))
    nbytes = sent_rows * sum(abi.TYPE_WIDTH[t] for _, t in meta)
    keep.append(recv)
    return abi.Chunk([(meta[k][0], recv[k], None, meta[k][1]) for k in range(len(meta))], num_rows=int(recv[0].numel()),
                     mem=abi.MEM_DEVICE), nbytes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sf", type=float, default=1.0)
    ap.add_argument("--no-check", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = gpu.Context(local, stream=stream.cuda_stream)

    t = tpch.gen_tables(args.sf)
    orders = {k: torch.from_numpy(v[rank::world].copy()).to(dev) for k, v in t["orders"].items()}
    lineitem = {k: torch.from_numpy(v[rank::world].copy()).to(dev) for k, v in t["lineitem"].items()}
    n_li, n_ord = int(lineitem["l_orderkey"].numel()), int(orders["o_orderkey"].numel())
    cust_scan, j1d, ord_scan, j2d, _ = tpch.q3_descs()
    li_scan = abi.ScanDesc(preds=[abi.make_pred(tpch.L_SHIPDATE, abi.PRED_GT, tpch.CUTOFF)],
                           out_slots=[tpch.L_ORDERKEY, tpch.L_EXTENDEDPRICE, tpch.L_DISCOUNT])
    payload = [tpch.O_ORDERDATE, tpch.O_SHIPPRIORITY]
    cust_chunk = tpch.table_chunk(t["customer"], tpch.CUSTOMER_COLS)
    customer_dev = {k: torch.from_numpy(v).to(dev) for k, v in t["customer"].items()}
    cust_chunk = tpch.table_chunk(customer_dev, tpch.CUSTOMER_COLS, mem=abi.MEM_DEVICE)

    def run_plan():
        """one execution of the plan with fresh operator handles; -> (result, stats, phase seconds)"""
        keep = []
        phases = {}
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        mark = [t0]

        def phase(name):
            torch.cuda.synchronize()
            now = time.perf_counter()
            phases[name] = phases.get(name, 0.0) + now - mark[0]
            mark[0] = now

        # --- build side ---
        s1 = gpu.Scan(ctx, cust_scan)
        b1 = tpch._dev_chunk(s1.filter(cust_chunk))
        j1 = gpu.Join(ctx, j1d)
        j1.append_build(b1)
        j1.build_finish()
        phase("customer scan + J1 build")
        s2 = gpu.Scan(ctx, ord_scan)
        o_f = tpch._dev_chunk(s2.filter(tpch.table_chunk(orders, tpch.ORDERS_COLS, mem=abi.MEM_DEVICE)))
        o_j = tpch._dev_chunk(j1.probe(o_f))
        phase("orders scan + semi-join probe")
        x_ord = gpu.Xchg(ctx, abi.make_part_desc([tpch.O_ORDERKEY], world))
        o_local, bytes_o = shuffle(ctx, x_ord, o_j, dev, keep)
        phase("orders partition + all-to-all")
        j2 = gpu.Join(ctx, j2d)
        j2.append_build(o_local)
        j2.build_finish()
        phase("J2 build")
        # --- probe side ---
        s3 = gpu.Scan(ctx, li_scan)
        l_f = tpch._dev_chunk(s3.filter(tpch.table_chunk(lineitem, tpch.LINEITEM_COLS, mem=abi.MEM_DEVICE)))
        phase("lineitem scan")
        x_li = gpu.Xchg(ctx, abi.make_part_desc([tpch.L_ORDERKEY], world))
        l_local, bytes_l = shuffle(ctx, x_li, l_f, dev, keep)
        phase("lineitem partition + all-to-all")
        frag = gpu.Fragment(ctx, abi.ScanDesc(), [(j2, tpch.L_ORDERKEY, payload)], tpch.q3_agg_desc())
        if l_local.num_rows > 0:
            frag.push(l_local)
        res = frag.agg.result()
        phase("probe J2 + aggregate + result")
        dist.barrier()
        dt = time.perf_counter() - t0
        st = (bytes_o + bytes_l, j2.info().build_rows, l_local.num_rows)
        frag.close()
        for h in (s1, s2, s3, j1, j2, x_ord, x_li):
            h.close()
        return res, st, dt, phases

    run_plan()                                   # warm-up: NCCL channels, allocator, kernel modules
    res, st, dt, phases = run_plan()
    bytes_total, j2_rows, probe_rows = st

    from tests.helpers import gpu_rows
    rows = gpu_rows(res)
    stats = torch.tensor([len(rows), sum(r[3] for r in rows) % (1 << 62), bytes_total, j2_rows, probe_rows], dtype=torch.int64, device=dev)
    dist.all_reduce(stats)
    gathered = [None] * world if rank == 0 else None
    if not args.no_check:
        dist.gather_object(rows, gathered, dst=0)
    if rank == 0:
        line = {"query": "TPC-H Q3", "sf": args.sf, "n_gpus": world, "lineitem_rows": len(t["lineitem"]["l_orderkey"]),
                "orders_rows": len(t["orders"]["o_orderkey"]), "groups": int(stats[0]), "shuffled_bytes": int(stats[2]),
                "j2_build_rows": int(stats[3]), "probe_rows_after_shuffle": int(stats[4]), "seconds": dt,
                "lineitem_rows_per_s": len(t["lineitem"]["l_orderkey"]) / dt,
                "phase_ms_rank0": {k: round(v * 1e3, 3) for k, v in phases.items()}}
        if not args.no_check:
            from oracle import oracle
            from tests.helpers import oracle_rows
            oj2, okeep = tpch.q3_build_oracle(oracle, t)
            _, _, _, _, full_scan = tpch.q3_descs()
            ores, _ = oracle.fragment_run(full_scan, [(oj2, tpch.L_ORDERKEY, payload)], tpch.q3_agg_desc(),
                                          tpch.table_chunk(t["lineitem"], tpch.LINEITEM_COLS), num_threads=os.cpu_count() or 1)
            union = sorted(r for part in gathered for r in part)
            exp = oracle_rows(ores)
            line["bit_exact_vs_oracle"] = union == exp
            line["oracle_groups"] = len(exp)
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
