"""Synthetic SSB tables (seeded) and the Q1.1 / Q4.1 plans expressed against the C-ABI descriptors.

Data follows SURVEY.md section 8d: all SSB columns are int32 (`int(11)`, test/common/sql/ssb/create.sql),
low-cardinality strings arrive as int32 global-dictionary codes.  seed = 20240921 + table_id.
The same plan builders drive the GPU library (starrocks_b200.gpu) and, in tests / the CPU baseline,
the oracle -- they only produce descriptors, they never compute.
"""
import datetime

import numpy as np

from . import abi

# slot ids (one namespace for the whole fragment, like a TupleDescriptor's slots)
LO_ORDERDATE, LO_CUSTKEY, LO_SUPPKEY, LO_PARTKEY, LO_REVENUE, LO_SUPPLYCOST = 0, 1, 2, 3, 4, 5
LO_QUANTITY, LO_DISCOUNT, LO_EXTENDEDPRICE = 6, 7, 8
C_CUSTKEY, C_REGION, C_NATION = 10, 11, 12
S_SUPPKEY, S_REGION = 20, 21
P_PARTKEY, P_MFGR = 30, 31
D_DATEKEY, D_YEAR = 40, 41
OUT_SUM_REVENUE, OUT_SUM_SUPPLYCOST, OUT_REVENUE = 50, 51, 52

AMERICA = 1  # dictionary code of c_region / s_region 'AMERICA'
SEED = 20240921


def sizes(sf):
    """SSB cardinalities at scale factor sf (lineorder ~ 6 M x sf; SF100 customer 3 M, supplier 0.2 M, part 1.4 M)."""
    return {
        "lineorder": int(round(6_000_000 * sf)),
        "customer": max(1, int(round(30_000 * sf))),
        "supplier": max(1, int(round(2_000 * sf))),
        "part": 200_000 * (1 + int(np.floor(np.log2(sf)))) if sf >= 1 else max(1, int(round(200_000 * sf))),
    }


def gen_dates():
    d0 = datetime.date(1992, 1, 1)
    days = [d0 + datetime.timedelta(days=i) for i in range(2556)]
    datekey = np.array([d.year * 10000 + d.month * 100 + d.day for d in days], dtype=np.int32)
    year = np.array([d.year for d in days], dtype=np.int32)
    return {"d_datekey": datekey, "d_year": year}


def gen_dims(sf, seed=SEED):
    sz = sizes(sf)
    rc = np.random.default_rng(seed + 1)
    rs = np.random.default_rng(seed + 2)
    rp = np.random.default_rng(seed + 3)
    c_region = rc.integers(0, 5, sz["customer"], dtype=np.int32)
    cust = {"c_custkey": np.arange(1, sz["customer"] + 1, dtype=np.int32), "c_region": c_region,
            "c_nation": (c_region * 5 + rc.integers(0, 5, sz["customer"], dtype=np.int32)).astype(np.int32)}
    supp = {"s_suppkey": np.arange(1, sz["supplier"] + 1, dtype=np.int32),
            "s_region": rs.integers(0, 5, sz["supplier"], dtype=np.int32)}
    part = {"p_partkey": np.arange(1, sz["part"] + 1, dtype=np.int32),
            "p_mfgr": rp.integers(0, 5, sz["part"], dtype=np.int32)}
    return {"customer": cust, "supplier": supp, "part": part, "dates": gen_dates()}


def gen_lineorder(sf, n=None, seed=SEED, dims_sizes=None):
    """numpy generator (tests, CPU side).  bench.py generates the same distributions on the device."""
    sz = dims_sizes or sizes(sf)
    n = sz["lineorder"] if n is None else n
    r = np.random.default_rng(seed + 0)
    datekey = gen_dates()["d_datekey"]
    return {
        "lo_orderdate": datekey[r.integers(0, len(datekey), n)].astype(np.int32),
        "lo_custkey": r.integers(1, sz["customer"] + 1, n, dtype=np.int32),
        "lo_suppkey": r.integers(1, sz["supplier"] + 1, n, dtype=np.int32),
        "lo_partkey": r.integers(1, sz["part"] + 1, n, dtype=np.int32),
        "lo_revenue": r.integers(81_000, 10_400_001, n, dtype=np.int32),
        "lo_supplycost": r.integers(54_000, 125_001, n, dtype=np.int32),
        "lo_quantity": r.integers(1, 51, n, dtype=np.int32),
        "lo_discount": r.integers(0, 11, n, dtype=np.int32),
        "lo_extendedprice": r.integers(90_000, 10_494_951, n, dtype=np.int32),
    }


LO_SLOTS = {"lo_orderdate": LO_ORDERDATE, "lo_custkey": LO_CUSTKEY, "lo_suppkey": LO_SUPPKEY,
            "lo_partkey": LO_PARTKEY, "lo_revenue": LO_REVENUE, "lo_supplycost": LO_SUPPLYCOST,
            "lo_quantity": LO_QUANTITY, "lo_discount": LO_DISCOUNT, "lo_extendedprice": LO_EXTENDEDPRICE}

Q41_FACT_COLS = ["lo_orderdate", "lo_custkey", "lo_suppkey", "lo_partkey", "lo_revenue", "lo_supplycost"]
Q11_FACT_COLS = ["lo_orderdate", "lo_discount", "lo_quantity", "lo_extendedprice"]


def fact_chunk(cols, names, mem=abi.MEM_HOST):
    """cols: dict name -> numpy array / torch tensor"""
    return abi.Chunk([(LO_SLOTS[nm], cols[nm], None, abi.TYPE_INT) for nm in names], mem=mem)


# ---- dimension-side plans (each is: scan + filter -> join build) ---------------------------------
def dim_plans_q41():
    """(table, key column, key slot, payload columns, scan predicates) for the four Q4.1 builds.
    fe/fe-core/src/test/resources/sql/ssb/Q4.1.sql: c_region = 'AMERICA', s_region = 'AMERICA',
    p_mfgr in ('MFGR#1','MFGR#2'); dictionary codes: AMERICA = 1, MFGR#1 = 0, MFGR#2 = 1."""
    return [
        ("supplier", "s_suppkey", S_SUPPKEY, LO_SUPPKEY, [], [abi.make_pred(S_REGION, abi.PRED_EQ, AMERICA)],
         {"s_suppkey": S_SUPPKEY, "s_region": S_REGION}),
        ("customer", "c_custkey", C_CUSTKEY, LO_CUSTKEY, [("c_nation", C_NATION)],
         [abi.make_pred(C_REGION, abi.PRED_EQ, AMERICA)],
         {"c_custkey": C_CUSTKEY, "c_region": C_REGION, "c_nation": C_NATION}),
        ("part", "p_partkey", P_PARTKEY, LO_PARTKEY, [], [abi.make_pred(P_MFGR, abi.PRED_IN, in_list=[0, 1])],
         {"p_partkey": P_PARTKEY, "p_mfgr": P_MFGR}),
        ("dates", "d_datekey", D_DATEKEY, LO_ORDERDATE, [("d_year", D_YEAR)], [],
         {"d_datekey": D_DATEKEY, "d_year": D_YEAR}),
    ]


def q41_agg_desc():
    """group by d_year, c_nation; sum(lo_revenue), sum(lo_supplycost) (profit = difference, final project)."""
    return abi.make_agg_desc(
        [D_YEAR, C_NATION], [abi.TYPE_INT, abi.TYPE_INT],
        fns=[(abi.AGG_SUM, abi.TYPE_INT, OUT_SUM_REVENUE, [("col", LO_REVENUE)]),
             (abi.AGG_SUM, abi.TYPE_INT, OUT_SUM_SUPPLYCOST, [("col", LO_SUPPLYCOST)])],
        ranges=[(1992, 1998), (0, 24)])


def q11_scan_preds():
    """SSB flat Q1.1: lo_orderdate in 1993, lo_discount between 1 and 3, lo_quantity < 25."""
    return [abi.make_pred(LO_ORDERDATE, abi.PRED_BETWEEN, 19930101, 19931231),
            abi.make_pred(LO_DISCOUNT, abi.PRED_BETWEEN, 1, 3),
            abi.make_pred(LO_QUANTITY, abi.PRED_LT, 25)]


def q11_agg_desc():
    """select sum(lo_extendedprice * lo_discount) as revenue"""
    return abi.make_agg_desc(fns=[(abi.AGG_SUM, abi.TYPE_BIGINT, OUT_REVENUE,
                                   [("col", LO_EXTENDEDPRICE), ("col", LO_DISCOUNT), "*"])])


def build_dims(impl, dims, plans, ctx=None):
    """Run the build side of every join through `impl` ('gpu' module or 'oracle' module).

    GPU: dimension chunk -> sr_scan_filter (device output) -> sr_join_append_build -> build_finish.
    Oracle: orc_scan_filter -> orc_join_append_build -> orc_join_build.
    returns list of (join, probe_key_slot, [payload slots]) in plan order, plus objects to keep alive.
    """
    joins, keep = [], []
    for table, key_col, key_slot, probe_slot, payload, preds, slotmap in plans:
        cols = dims[table]
        chunk = abi.Chunk([(slotmap[nm], cols[nm], None, abi.TYPE_INT) for nm in cols])
        out_slots = [key_slot] + [s for _, s in payload]
        sd = abi.ScanDesc(preds=preds, out_slots=out_slots)
        jd = abi.make_join_desc(abi.JOIN_INNER, [key_slot], [probe_slot], [abi.TYPE_INT],
                                build_out=[s for _, s in payload], probe_out=[])
        if ctx is not None:  # GPU
            scan = impl.Scan(ctx, sd)
            out = scan.filter(chunk)
            bchunk = abi.Chunk([(out.cols[k].slot_id, out.cols[k].data, out.cols[k].nulls, out.cols[k].type)
                                for k in range(out.num_cols)], num_rows=out.num_rows, mem=abi.MEM_DEVICE)
            j = impl.Join(ctx, jd)
            j.append_build(bchunk)
            j.build_finish()
            keep.append((scan, chunk, bchunk))
        else:
            rows, res = impl.scan_filter(sd, chunk)
            bchunk = abi.Chunk([(s, res[s][0], res[s][1], abi.TYPE_INT) for s in out_slots])
            j = impl.Join(jd)
            j.append_build(bchunk)
            j.build()
            keep.append((chunk, bchunk))
        joins.append((j, probe_slot, [s for _, s in payload]))
    return joins, keep
