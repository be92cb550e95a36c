"""Synthetic TPC-H tables for Q3 (BASELINE.json config "TPC-H Q3: lineitem JOIN orders JOIN customer, hash join +
aggregate") and the Q3 plan expressed through either implementation of the operator interface (the GPU library or
the CPU oracle).  Shapes follow SURVEY.md section 8d: int32 keys, money = decimal64(15,2) (scaled int64), dates =
int32 julian day numbers (types/date_value.h:120), sparse o_orderkey (8 of every 32 values, TPC-H spec 4.2.3).

Plan (fe/fe-core/src/test/resources/sql/tpch/q3.sql):
    customer  -- c_mktsegment = 'BUILDING'              --> build J1 (c_custkey)
    orders    -- o_orderdate < 1995-03-15, SEMI J1      --> build J2 (o_orderkey; payload o_orderdate, o_shippriority)
    lineitem  -- l_shipdate  > 1995-03-15, INNER J2     --> GROUP BY l_orderkey, o_orderdate, o_shippriority
                                                            SUM(l_extendedprice * (1 - l_discount))   [decimal128]
The 12-byte group key exercises the 16-byte packed-key hash table (serialized-fixed-size key, aggregator.cpp:1485-1566).
"""
import numpy as np

from . import abi

SEED = 20240921
JULIAN_1992_01_01 = 2448623
CUTOFF = JULIAN_1992_01_01 + 1169        # 1995-03-15
ORDERDATE_SPAN = 2406                    # 1992-01-01 .. 1998-08-02
BUILDING = 1

C_CUSTKEY, C_MKTSEGMENT = 100, 101
O_ORDERKEY, O_CUSTKEY, O_ORDERDATE, O_SHIPPRIORITY = 110, 111, 112, 113
L_ORDERKEY, L_EXTENDEDPRICE, L_DISCOUNT, L_SHIPDATE = 120, 121, 122, 123
OUT_REVENUE = 130

CUSTOMER_COLS = [("c_custkey", C_CUSTKEY, abi.TYPE_INT), ("c_mktsegment", C_MKTSEGMENT, abi.TYPE_INT)]
ORDERS_COLS = [("o_orderkey", O_ORDERKEY, abi.TYPE_INT), ("o_custkey", O_CUSTKEY, abi.TYPE_INT),
               ("o_orderdate", O_ORDERDATE, abi.TYPE_DATE), ("o_shippriority", O_SHIPPRIORITY, abi.TYPE_INT)]
LINEITEM_COLS = [("l_orderkey", L_ORDERKEY, abi.TYPE_INT), ("l_extendedprice", L_EXTENDEDPRICE, abi.TYPE_DECIMAL64),
                 ("l_discount", L_DISCOUNT, abi.TYPE_DECIMAL64), ("l_shipdate", L_SHIPDATE, abi.TYPE_DATE)]


def gen_tables(sf, seed=SEED):
    """-> dict table -> dict column -> numpy array.  customer 150 000 x sf, orders 10 x customers, lineitem 1..7 per order."""
    rng = np.random.default_rng(seed)
    nc = max(3, int(150_000 * sf))
    no = nc * 10
    customer = {"c_custkey": np.arange(1, nc + 1, dtype=np.int32), "c_mktsegment": rng.integers(0, 5, nc, dtype=np.int32)}
    i = np.arange(no, dtype=np.int64)
    o_orderkey = ((i // 8) * 32 + (i % 8) + 1).astype(np.int32)
    ck = rng.integers(1, nc + 1, no, dtype=np.int32)
    ck = np.where(ck % 3 == 0, np.maximum(ck - 1, 1), ck).astype(np.int32)      # a third of the customers place no order
    o_orderdate = (JULIAN_1992_01_01 + rng.integers(0, ORDERDATE_SPAN, no)).astype(np.int32)
    orders = {"o_orderkey": o_orderkey, "o_custkey": ck, "o_orderdate": o_orderdate, "o_shippriority": np.zeros(no, dtype=np.int32)}
    per = rng.integers(1, 8, no)
    l_orderkey = np.repeat(o_orderkey, per)
    nl = len(l_orderkey)
    lineitem = {"l_orderkey": l_orderkey,
                "l_extendedprice": rng.integers(90_000, 10_494_951, nl, dtype=np.int64),
                "l_discount": rng.integers(0, 11, nl, dtype=np.int64),
                "l_shipdate": (np.repeat(o_orderdate, per) + rng.integers(1, 122, nl)).astype(np.int32)}
    # the fact table is not clustered by order in a shuffled plan: permute it
    p = rng.permutation(nl)
    lineitem = {k: np.ascontiguousarray(v[p]) for k, v in lineitem.items()}
    return {"customer": customer, "orders": orders, "lineitem": lineitem}


def table_chunk(cols, spec, mem=abi.MEM_HOST, rows=None):
    if rows is None:
        return abi.Chunk([(slot, cols[nm], None, typ) for nm, slot, typ in spec], mem=mem)
    lo, hi = rows
    return abi.Chunk([(slot, cols[nm][lo:hi], None, typ) for nm, slot, typ in spec], mem=mem)


def q3_agg_desc(expected_groups=0):
    return abi.make_agg_desc([L_ORDERKEY, O_ORDERDATE, O_SHIPPRIORITY], [abi.TYPE_INT, abi.TYPE_DATE, abi.TYPE_INT],
                             fns=[(abi.AGG_SUM, abi.TYPE_DECIMAL64, OUT_REVENUE,
                                   [("col", L_EXTENDEDPRICE), ("i", 100), ("col", L_DISCOUNT), "-", "*"])],
                             expected_groups=expected_groups)


def q3_descs():
    cust_scan = abi.ScanDesc(preds=[abi.make_pred(C_MKTSEGMENT, abi.PRED_EQ, BUILDING)], out_slots=[C_CUSTKEY])
    j1 = abi.make_join_desc(abi.JOIN_LEFT_SEMI, [C_CUSTKEY], [O_CUSTKEY], [abi.TYPE_INT],
                            probe_out=[O_ORDERKEY, O_ORDERDATE, O_SHIPPRIORITY])
    ord_scan = abi.ScanDesc(preds=[abi.make_pred(O_ORDERDATE, abi.PRED_LT, CUTOFF)],
                            out_slots=[O_ORDERKEY, O_CUSTKEY, O_ORDERDATE, O_SHIPPRIORITY])
    j2 = abi.make_join_desc(abi.JOIN_INNER, [O_ORDERKEY], [L_ORDERKEY], [abi.TYPE_INT], build_out=[O_ORDERDATE, O_SHIPPRIORITY])
    li_scan = abi.ScanDesc(preds=[abi.make_pred(L_SHIPDATE, abi.PRED_GT, CUTOFF)])
    return cust_scan, j1, ord_scan, j2, li_scan


def _dev_chunk(out):
    return abi.Chunk([(out.cols[k].slot_id, out.cols[k].data, out.cols[k].nulls, out.cols[k].type) for k in range(out.num_cols)],
                     num_rows=out.num_rows, mem=abi.MEM_DEVICE)


def q3_build_gpu(gpu, ctx, tables):
    """customer -> J1, orders SEMI J1 -> J2 on the GPU.  -> (J2, objects to keep alive / close)"""
    cust_scan, j1d, ord_scan, j2d, _ = q3_descs()
    keep = []
    s1 = gpu.Scan(ctx, cust_scan)
    c_chunk = table_chunk(tables["customer"], CUSTOMER_COLS)
    b1 = _dev_chunk(s1.filter(c_chunk))
    j1 = gpu.Join(ctx, j1d)
    j1.append_build(b1)
    j1.build_finish()
    s2 = gpu.Scan(ctx, ord_scan)
    o_chunk = table_chunk(tables["orders"], ORDERS_COLS)
    o_f = _dev_chunk(s2.filter(o_chunk))
    o_j = _dev_chunk(j1.probe(o_f))
    j2 = gpu.Join(ctx, j2d)
    j2.append_build(o_j)
    j2.build_finish()
    keep += [s1, s2, j1, c_chunk, o_chunk, b1, o_f, o_j]
    return j2, keep


def q3_build_oracle(oracle, tables):
    cust_scan, j1d, ord_scan, j2d, _ = q3_descs()
    _, res = oracle.scan_filter(cust_scan, table_chunk(tables["customer"], CUSTOMER_COLS))
    j1 = oracle.Join(j1d)
    j1.append_build(abi.Chunk([(C_CUSTKEY, res[C_CUSTKEY][0], res[C_CUSTKEY][1], abi.TYPE_INT)]))
    j1.build()
    _, ores = oracle.scan_filter(ord_scan, table_chunk(tables["orders"], ORDERS_COLS))
    types = {slot: typ for _, slot, typ in ORDERS_COLS}
    o_f = abi.Chunk([(s, ores[s][0], ores[s][1], types[s]) for s in (O_ORDERKEY, O_CUSTKEY, O_ORDERDATE, O_SHIPPRIORITY)])
    pi, bi = j1.probe_all(o_f)
    out = j1.output(o_f, pi, bi)
    o_j = abi.Chunk([(s, d, None, types[s]) for s, d, _ in out])
    j2 = oracle.Join(j2d)
    j2.append_build(o_j)
    j2.build()
    return j2, [j1, o_f, o_j]


# ---------------------------------------------------------------------------------------------------------------------
# Shardable generator (SF300 on 8 GPUs: no rank ever holds a whole table).  Every column value is a pure function of the
# row's logical index: counter-based hashing (splitmix64 of index and column id) instead of a sequential RNG, so that any
# rank can produce any slice on its own device, the ranks need not agree on anything but (sf, world), and the host can
# reproduce the same tables with numpy for the oracle check.  Same shapes / distributions as gen_tables above.
# ---------------------------------------------------------------------------------------------------------------------
_M64 = (1 << 64) - 1


def _sm64_np(x):
    """splitmix64 output function on uint64 numpy arrays (wrapping)"""
    with np.errstate(over="ignore"):
        z = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)).astype(np.uint64)
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)).astype(np.uint64)
        return z ^ (z >> np.uint64(31))


def _sm64_torch(x):
    """the same on int64 torch tensors (two's complement wrapping; logical shifts emulated by masking)"""
    z = x + (-7046029254386353131)
    z = (z ^ ((z >> 30) & ((1 << 34) - 1))) * (-4658895280553007687)
    z = (z ^ ((z >> 27) & ((1 << 37) - 1))) * (-7723592293110705685)
    return z ^ ((z >> 31) & ((1 << 33) - 1))


class HashGen:
    """column values as functions of the row index.  xp = "np" (host, oracle check) or a torch device."""

    def __init__(self, sf, device=None):
        self.nc = max(3, int(150_000 * sf))
        self.no = self.nc * 10
        self.dev = device

    def _h(self, col, idx):
        """63-bit non-negative hash of (column id, index)"""
        if self.dev is None:
            return (_sm64_np(idx.astype(np.uint64) * np.uint64(16) + np.uint64(col)) >> np.uint64(1)).astype(np.int64)
        return (_sm64_torch(idx * 16 + col) >> 1) & ((1 << 62) - 1 | (1 << 62))

    def _arange(self, lo, hi):
        if self.dev is None:
            return np.arange(lo, hi, dtype=np.int64)
        import torch
        return torch.arange(lo, hi, dtype=torch.int64, device=self.dev)

    def _i32(self, x):
        return x.astype(np.int32) if self.dev is None else x.to(dtype=__import__("torch").int32)

    def customer(self, lo=0, hi=None):
        hi = self.nc if hi is None else hi
        i = self._arange(lo, hi)
        return {"c_custkey": self._i32(i + 1), "c_mktsegment": self._i32(self._h(1, i) % 5)}

    def order_cols(self, i):
        """columns of the orders with logical indexes `i` (int64 array / tensor)"""
        okey = (i // 8) * 32 + (i % 8) + 1
        ck = 1 + self._h(2, i) % self.nc
        if self.dev is None:
            ck = np.where(ck % 3 == 0, np.maximum(ck - 1, 1), ck)       # a third of the customers place no order
        else:
            import torch
            ck = torch.where(ck % 3 == 0, torch.clamp(ck - 1, min=1), ck)
        odate = JULIAN_1992_01_01 + self._h(3, i) % ORDERDATE_SPAN
        return okey, ck, odate

    def orders(self, idx):
        okey, ck, odate = self.order_cols(idx)
        zeros = np.zeros(len(idx), dtype=np.int32) if self.dev is None else __import__("torch").zeros(idx.numel(), dtype=__import__("torch").int32, device=self.dev)
        return {"o_orderkey": self._i32(okey), "o_custkey": self._i32(ck), "o_orderdate": self._i32(odate), "o_shippriority": zeros}

    def lineitem_of_orders(self, lo, hi):
        """the line items of the orders with logical indexes [lo, hi): 1..7 per order"""
        i = self._arange(lo, hi)
        per = 1 + self._h(4, i) % 7
        okey, _, odate = self.order_cols(i)
        if self.dev is None:
            rep = np.repeat(np.arange(hi - lo, dtype=np.int64), per)
            first = np.cumsum(per) - per
            line = np.arange(len(rep), dtype=np.int64) - first[rep]
        else:
            import torch
            rep = torch.repeat_interleave(torch.arange(hi - lo, dtype=torch.int64, device=self.dev), per)
            first = torch.cumsum(per, 0) - per
            line = torch.arange(rep.numel(), dtype=torch.int64, device=self.dev) - first[rep]
        key = (i[rep]) * 8 + line
        price = 90_000 + self._h(5, key) % (10_494_951 - 90_000)
        disc = self._h(6, key) % 11
        ship = odate[rep] + 1 + self._h(7, key) % 121
        return {"l_orderkey": self._i32(okey[rep]), "l_extendedprice": price, "l_discount": disc, "l_shipdate": self._i32(ship)}

    def order_qualifies(self, i):
        """Q3's build-side conditions for the orders with logical indexes i: o_orderdate < cutoff and the customer's segment"""
        _, ck, odate = self.order_cols(i)
        seg = self._h(1, ck - 1) % 5
        return (odate < CUTOFF) & (seg == BUILDING)


def gen_tables_hashed(sf):
    """the whole tables on the host from HashGen (for the oracle check of the sharded runs)"""
    g = HashGen(sf)
    return {"customer": g.customer(), "orders": g.orders(np.arange(g.no, dtype=np.int64)), "lineitem": g.lineitem_of_orders(0, g.no)}
