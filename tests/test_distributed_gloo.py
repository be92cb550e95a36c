"""N > 1 host logic on CPU: world-size-2 gloo runs of (a) the two-phase aggregate (partial states gathered to rank 0
and merged) and (b) the HASH_PARTITIONED exchange (FNV + ReduceOp partition, all_to_all, local join + aggregate).
The per-rank compute is done by the oracle here (no GPU in this container); on the GPU box bench.py drives the same
starrocks_b200.distributed functions over NCCL with the CUDA operators."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from starrocks_b200 import abi, ssb
from starrocks_b200.abi import Chunk


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _init(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _two_phase_worker(rank, world, port, q):
    from oracle import oracle
    from starrocks_b200.distributed import all_gather_partial_states, gather_partial_states
    from tests.helpers import oracle_rows
    _init(rank, world, port)
    dims = ssb.gen_dims(0.05)
    n = 200_000
    lo = ssb.gen_lineorder(0.05, n=n)                       # every rank generates the same table and takes its shard
    shard = {k: v[rank::world].copy() for k, v in lo.items()}
    ojoins, keep = ssb.build_dims(oracle, dims, ssb.dim_plans_q41())      # broadcast join: dimensions replicated
    part, _ = oracle.fragment_run(abi.ScanDesc(), ojoins, ssb.q41_agg_desc(), ssb.fact_chunk(shard, ssb.Q41_FACT_COLS), num_threads=1)
    out = part.output()
    cols = [torch.from_numpy(np.ascontiguousarray(o[1]).astype(np.int64)) for o in out]
    gathered = gather_partial_states(cols, max_rows=175, dst=0)
    packed = all_gather_partial_states(cols, max_rows=175, dst=0)      # same exchange, one collective (bench.py's path)
    if rank == 0:
        for k in range(len(cols)):
            assert torch.equal(packed[k], torch.cat([g[k] for g in gathered]))
    else:
        assert packed is None and gathered is None
    if rank == 0:
        final_desc = abi.make_agg_desc([ssb.D_YEAR, ssb.C_NATION], [abi.TYPE_INT, abi.TYPE_INT],
                                       fns=[(abi.AGG_SUM, abi.TYPE_BIGINT, 50, [("col", 50)]), (abi.AGG_SUM, abi.TYPE_BIGINT, 51, [("col", 51)])])
        final = oracle.Agg(final_desc)
        for g in gathered:
            final.push(Chunk([(ssb.D_YEAR, g[0].numpy().astype(np.int32), None), (ssb.C_NATION, g[1].numpy().astype(np.int32), None),
                              (50, g[2].numpy().copy(), None), (51, g[3].numpy().copy(), None)]))
        whole, _ = oracle.fragment_run(abi.ScanDesc(), ojoins, ssb.q41_agg_desc(), ssb.fact_chunk(lo, ssb.Q41_FACT_COLS), num_threads=1)
        q.put(oracle_rows(final) == oracle_rows(whole))
    dist.barrier()
    dist.destroy_process_group()


def _intermediate_format_worker(rank, world, port, q):
    """general two-phase aggregate over the intermediate format (sr_agg_two_phase_descs): AVG / MIN / MAX / COUNT of a
    nullable column, first phase per rank (half of its rows pre-aggregated, half passed through), state chunks gathered
    to rank 0 in one collective, merge phase there; equals the single-phase result on the whole table"""
    from oracle import oracle
    from starrocks_b200.distributed import all_gather_partial_states
    from tests.helpers import assert_rows_equal, oracle_rows, rand_nulls
    _init(rank, world, port)
    rng = np.random.default_rng(5)
    n = 60_000
    key = rng.integers(0, 3000, n, dtype=np.int32)
    v = rng.integers(-10**6, 10**6, n, dtype=np.int32)
    vn = rand_nulls(rng, n, 0.3)
    x = rng.normal(0, 100, n)
    d = abi.make_agg_desc([0], [abi.TYPE_INT], fns=[(abi.AGG_AVG, abi.TYPE_INT, 10, [("col", 1)]), (abi.AGG_COUNT, abi.TYPE_INT, 11, [("col", 1)]),
                                                    (abi.AGG_MIN, abi.TYPE_INT, 12, [("col", 1)]), (abi.AGG_MAX, abi.TYPE_DOUBLE, 13, [("col", 2)]),
                                                    (abi.AGG_SUM, abi.TYPE_DOUBLE, 14, [("col", 2)]), (abi.AGG_COUNT_STAR, 0, 15, None)])
    p1, p2 = abi.two_phase_descs(d)
    mine = slice(rank, None, world)
    k_, v_, vn_, x_ = key[mine].copy(), v[mine].copy(), vn[mine].copy(), x[mine].copy()
    half = len(k_) // 2
    pre = oracle.Agg(p1)
    pre.push(Chunk([(0, k_[:half].copy(), None), (1, v_[:half].copy(), vn_[:half].copy()), (2, x_[:half].copy(), None)]))
    slots = [p1.group_slots[0]] + [p1.fns[f].out_slot for f in range(p1.num_fns)]
    a = [(s, t, dd, nl) for s, (t, dd, nl) in zip(slots, pre.output())]
    b = oracle.convert_to_states(p1, Chunk([(0, k_[half:].copy(), None), (1, v_[half:].copy(), vn_[half:].copy()), (2, x_[half:].copy(), None)]))
    # ship every column as int64 words (doubles bit-cast) plus its null bytes
    cols = []
    for (s, t, da, na), (_, _, db, nb) in zip(a, b):
        data = np.concatenate([np.asarray(da), np.asarray(db)])
        bits = data.view(np.int64) if data.dtype == np.float64 else data.astype(np.int64)
        nul = np.concatenate([np.zeros(len(da), np.uint8) if na is None else na, np.zeros(len(db), np.uint8) if nb is None else nb])
        cols += [torch.from_numpy(np.ascontiguousarray(bits)), torch.from_numpy(nul.astype(np.int64))]
    packed = all_gather_partial_states(cols, max_rows=len(k_), dst=0)
    if rank == 0:
        final = oracle.Agg(p2)
        chunk_cols = []
        for i, (s, t, _, _) in enumerate(a):
            bits = packed[2 * i].numpy()
            data = bits.view(np.float64).copy() if t == abi.TYPE_DOUBLE else bits.astype(abi.TYPE_NUMPY[t])
            chunk_cols.append((s, np.ascontiguousarray(data), packed[2 * i + 1].numpy().astype(np.uint8), t))
        final.push(Chunk(chunk_cols))
        whole = oracle.Agg(d)
        whole.push(Chunk([(0, key, None), (1, v, vn), (2, x, None)]))
        try:
            assert_rows_equal(oracle_rows(final), oracle_rows(whole), float_cols=(1, 5))
            q.put(True)
        except AssertionError as e:
            print(e)
            q.put(False)
    dist.barrier()
    dist.destroy_process_group()


def _global_runtime_filter_worker(rank, world, port, q):
    """partitioned join: each rank builds the PARTIAL filter of its build-side shard (sized for the global row count),
    the directories are OR-ed and min / max / has_null reduced across ranks; the merged filter equals the filter built
    from the whole build side, byte for byte"""
    from oracle import oracle
    from starrocks_b200.distributed import all_reduce_runtime_filter
    from tests.helpers import rand_nulls
    _init(rank, world, port)
    rng = np.random.default_rng(3)
    n = 50_000
    keys = rng.integers(-10**7, 10**7, n, dtype=np.int64)
    nulls = rand_nulls(rng, n, 0.01)
    part = oracle.RuntimeFilter(abi.TYPE_BIGINT, n)                      # expected_rows = GLOBAL build rows
    part.insert(Chunk([(0, keys[rank::world].copy(), nulls[rank::world].copy())]), 0, insert_nulls=(rank == 1))
    inf = part.info()
    directory = torch.from_numpy(part.directory().view(np.int32).copy())
    mn, mx, cnt, has_null = all_reduce_runtime_filter(directory, inf.min_value, inf.max_value, inf.num_inserted, inf.has_null != 0)
    whole = oracle.RuntimeFilter(abi.TYPE_BIGINT, n)
    whole.insert(Chunk([(0, keys, nulls)]), 0, insert_nulls=True)
    w = whole.info()
    ok = np.array_equal(directory.numpy().view(np.uint32), whole.directory())
    # rank 0 inserted no NULLs, rank 1 did: the union has them; counts add up to the non-NULL keys
    ok = ok and (mn, mx, cnt, has_null) == (w.min_value, w.max_value, w.num_inserted, True)
    # and the merged bytes fed into an empty filter evaluate like the whole-side filter
    merged = oracle.RuntimeFilter(abi.TYPE_BIGINT, n)
    other = oracle.RuntimeFilter(abi.TYPE_BIGINT, n)
    other.insert(Chunk([(0, keys, nulls)]), 0, insert_nulls=True)
    merged.merge(other)
    probe = Chunk([(0, rng.integers(-10**7, 10**7, 20_000, dtype=np.int64), rand_nulls(rng, 20_000, 0.02))])
    ok = ok and np.array_equal(merged.evaluate(probe, 0), whole.evaluate(probe, 0)) and np.array_equal(merged.directory(), directory.numpy().view(np.uint32))
    # the IN part of a global filter: the union of the partial key lists, gone as soon as one partial filter has none or the
    # union exceeds the row limit
    from starrocks_b200.distributed import all_gather_in_values
    small = np.arange(rank * 300, rank * 300 + 400, dtype=np.int64)         # overlapping ranges: 700 distinct keys over 2 ranks
    got = all_gather_in_values(small)
    ok = ok and got is not None and np.array_equal(got, np.arange(0, 300 * (world - 1) + 400))
    ok = ok and all_gather_in_values(small if rank == 0 else None) is None
    ok = ok and all_gather_in_values(np.arange(rank * 1000, rank * 1000 + 600, dtype=np.int64)) is None
    q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def _shuffle_worker(rank, world, port, q):
    from oracle import oracle
    from starrocks_b200.distributed import exchange_partitions, gather_partial_states
    from tests.helpers import oracle_rows
    _init(rank, world, port)
    rng = np.random.default_rng(77)
    n = 100_000
    keys = rng.integers(0, 5000, n, dtype=np.int32)
    vals = rng.integers(-1000, 1000, n, dtype=np.int64)
    my_k, my_v = keys[rank::world].copy(), vals[rank::world].copy()
    pd_ = abi.make_part_desc([0], world)                    # HASH_PARTITIONED: FNV + ReduceOp
    hv, ch, ri, st = oracle.hash_partition(pd_, Chunk([(0, my_k, None), (1, my_v, None)]))
    recv = exchange_partitions([torch.from_numpy(my_k[ri]), torch.from_numpy(my_v[ri])], st.tolist())
    rk, rv = recv[0].numpy(), recv[1].numpy()
    # every key now lives on exactly one rank: ReduceOp(fnv(key), world) == rank for all received rows
    hv2, ch2, _, _ = oracle.hash_partition(pd_, Chunk([(0, rk.copy(), None), (1, rv.copy(), None)]))
    ok = bool((ch2 == rank).all())
    d = abi.make_agg_desc([0], [abi.TYPE_INT], fns=[(abi.AGG_SUM, abi.TYPE_BIGINT, 10, [("col", 1)]), (abi.AGG_COUNT_STAR, 0, 11, None)])
    local = oracle.Agg(d)
    local.push(Chunk([(0, rk.copy(), None), (1, rv.copy(), None)]))
    out = local.output()
    ng = torch.tensor([local.num_groups], dtype=torch.int64)
    dist.all_reduce(ng)
    cols = [torch.from_numpy(np.ascontiguousarray(o[1]).astype(np.int64)) for o in out]
    gathered = gather_partial_states(cols, max_rows=5000, dst=0)
    if rank == 0:
        whole = oracle.Agg(d)
        whole.push(Chunk([(0, keys, None), (1, vals, None)]))
        rows = sorted(tuple(int(x) for x in r) for g in gathered for r in zip(g[0].tolist(), g[1].tolist(), g[2].tolist()))
        q.put(ok and int(ng.item()) == whole.num_groups and rows == [tuple(r) for r in oracle_rows(whole)])
    else:
        q.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def _q95_worker(rank, world, port, q):
    """the distributed TPC-DS Q95 plan (tools/q95_distributed.py) on the CPU: each rank generates its block of orders, both tables
    are HASH_PARTITIONED on the order number (oracle.hash_partition = FNV + ReduceOp) and exchanged, the local plan runs on
    the oracle engine, the three results are summed -- and must equal the join-free evaluation of the whole query"""
    from oracle import oracle
    from starrocks_b200 import tpcds
    from starrocks_b200.distributed import exchange_partitions
    _init(rank, world, port)
    g = tpcds.Q95Gen(0.4)
    blk = (g.n_orders + world - 1) // world
    lo, hi = min(g.n_orders, rank * blk), min(g.n_orders, (rank + 1) * blk)
    ws, wr = g.web_sales_of_orders(lo, hi), g.web_returns_of_orders(lo, hi)

    def shuffle(table, cols, key_slot):
        chunk = tpcds.table_chunk(table, cols)
        _, _, ri, st = oracle.hash_partition(abi.make_part_desc([key_slot], world), chunk)
        recv = exchange_partitions([torch.from_numpy(np.ascontiguousarray(table[name][ri])) for name, _, _ in cols], st.tolist())
        return Chunk([(slot, r.numpy(), None, typ) for (_, slot, typ), r in zip(cols, recv)])

    ws_local = shuffle(ws, tpcds.WS_COLS, tpcds.WS_ORDER)
    wr_local = shuffle(wr, [("wr_order_number", tpcds.WS_ORDER, abi.TYPE_BIGINT)], tpcds.WS_ORDER)
    dims = {"date": Chunk([(tpcds.D_DATE_SK, g.date_keys(), None)]), "addr": Chunk([(tpcds.CA_ADDRESS_SK, g.address_keys(), None)]),
            "site": Chunk([(tpcds.WEB_SITE_SK, g.site_keys(), None)])}
    res, st = tpcds.q95_local_plan(tpcds.OracleEngine(oracle), ws_local, wr_local, dims, morsel_rows=40_000)
    tot = torch.tensor(list(res) + [ws_local.num_rows], dtype=torch.int64)
    dist.all_reduce(tot)
    exp = g.expected(0, g.n_orders)
    ok = tuple(int(x) for x in tot[:3]) == exp and exp[0] > 10
    ok = ok and int(tot[3]) == len(g.web_sales_of_orders(0, g.n_orders)["ws_order_number"])      # no row lost or duplicated by the exchange
    # every order is complete on exactly one rank: its order numbers hash to this rank
    _, ch, _, _ = oracle.hash_partition(abi.make_part_desc([tpcds.WS_ORDER], world), ws_local)
    ok = ok and bool((ch == rank).all())
    q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("worker", [_two_phase_worker, _intermediate_format_worker, _global_runtime_filter_worker, _shuffle_worker, _q95_worker])
def test_world_size_2_gloo(oracle, worker):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = []
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0, f"rank exited with {p.exitcode}"
    while not q.empty():
        results.append(q.get())
    assert results and all(results)
