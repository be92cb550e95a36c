"""GPU parity tests: the CUDA path (through the C-ABI of libsr_gpu.so) against the CPU oracle on
the same seeded inputs.  Bit-exact for integer / index / COUNT / SUM(int64) / decimal work; 1e-6
relative for double SUM / AVG (the tolerance BASELINE.json states).  Join and group-by outputs
are compared as sorted multisets (row order is an implementation artefact in the reference too).
"""
import numpy as np
import pytest

from starrocks_b200 import abi, ssb
from starrocks_b200.abi import Chunk
from tests.helpers import assert_rows_equal, gpu_rows, oracle_rows, rand_nulls, rows_sorted, col_to_py

pytestmark = pytest.mark.gpu

SIZES = [0, 1, 31, 255, 256, 1025, 4096, 100003]


# ---------------------------------------------------------------------------------------------
# K5 JoinKeyHash on the device, pinned by the reference's golden vectors
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,stride,expect", [
    (np.int32, 3, (0, 11)), (np.int32, 7, (0, 14)), (np.int32, 1, (4, 6)),
    (np.int64, 3, (3, 7)), (np.int64, 7, (4, 7)), (np.int64, 1, (4, 6)),
])
def test_join_key_hash_golden(gpu, ctx, dtype, stride, expect):
    # be/test/exec/join_hash_map_test.cpp:924-1009
    nb = 1 << 16
    keys = np.arange(0, nb * stride * 5, stride, dtype=dtype)
    buckets = ctx.join_key_hash(keys, 16)
    counts = np.bincount(buckets, minlength=nb)
    assert (counts.min(), counts.max()) == expect
    assert ctx.join_key_hash(np.array([1, 2, 3, 4], dtype=np.int32), 2).tolist() == [2, 0, 3, 1]


# ---------------------------------------------------------------------------------------------
# scan predicates + order-preserving compaction (K1-K4)
# ---------------------------------------------------------------------------------------------
def _scan_inputs(n, seed=7):
    rng = np.random.default_rng(seed)
    return Chunk([
        (0, rng.integers(-50, 50, n, dtype=np.int32), None),
        (1, rng.integers(-10**12, 10**12, n, dtype=np.int64), rand_nulls(rng, n)),
        (2, rng.normal(0, 10, n), rand_nulls(rng, n, 0.05)),
        (3, rng.integers(-100, 100, n).astype(np.int16), None),
        (4, rng.integers(0, 2, n).astype(np.uint8), None, abi.TYPE_BOOLEAN),
        (5, rng.random(n).astype(np.float32), None),
    ])


SCAN_CASES = [
    dict(preds=[abi.make_pred(0, abi.PRED_BETWEEN, -10, 20)]),
    dict(preds=[abi.make_pred(0, abi.PRED_GE, 0), abi.make_pred(1, abi.PRED_LT, 0), abi.make_pred(3, abi.PRED_NE, 7)]),
    dict(preds=[abi.make_pred(2, abi.PRED_GT, 1.5, is_double=True), abi.make_pred(5, abi.PRED_LE, 0.5, is_double=True)]),
    dict(preds=[abi.make_pred(0, abi.PRED_IN, in_list=[1, 2, 3, -7]), abi.make_pred(1, abi.PRED_IS_NOT_NULL)]),
    dict(preds=[abi.make_pred(1, abi.PRED_IS_NULL)]),
    dict(preds=[abi.make_pred(3, abi.PRED_NOT_IN, in_list=[0, 1])]),
    dict(preds=[abi.make_pred(0, abi.PRED_EQ, 1000)]),                      # nothing passes
    dict(preds=[]),                                                         # everything passes
    dict(exprs=[abi.make_expr([("col", 0), ("i", 2), "*", ("col", 3), "+", ("i", 10), ">"])]),
    dict(exprs=[abi.make_expr([("col", 1), ("i", 0), ">", ("col", 2), ("d", 0.0), "<", "or"])]),   # 3-valued OR
    dict(exprs=[abi.make_expr([("col", 1), "isnull", "not", ("col", 0), ("col", 3), "<=", "and"])]),
    dict(preds=[abi.make_pred(4, abi.PRED_EQ, 1)], exprs=[abi.make_expr([("col", 2), ("col", 5), "/", ("d", 3.0), ">"])]),
]


@pytest.mark.parametrize("case", range(len(SCAN_CASES)))
@pytest.mark.parametrize("n", SIZES)
def test_scan_filter_parity(gpu, ctx, oracle, case, n):
    c = SCAN_CASES[case]
    chunk = _scan_inputs(n)
    sd = abi.ScanDesc(preds=c.get("preds", []), filter_exprs=c.get("exprs", []), out_slots=[0, 1, 2, 3, 4, 5])
    scan = gpu.Scan(ctx, sd)
    try:
        sel = scan.evaluate(chunk)
        assert np.array_equal(sel, oracle.scan_evaluate(sd, chunk))
        out = gpu.chunk_out_to_host(ctx, scan.filter(chunk))
        rows, exp = oracle.scan_filter(sd, chunk)
        assert len(out) == 6
        for slot, typ, data, nulls in out:
            assert len(data) == rows
            assert np.array_equal(data.view(np.uint8), exp[slot][0].view(np.uint8)), f"slot {slot}"  # bit-exact, in order
            if exp[slot][1] is not None:
                assert np.array_equal(nulls, exp[slot][1])
    finally:
        scan.close()


# ---------------------------------------------------------------------------------------------
# hash join build + probe (K6-K12)
# ---------------------------------------------------------------------------------------------
def _join_pairs(gpu, ctx, oracle, desc, build_chunks, probe_chunk, expect_method=None):
    gj = gpu.Join(ctx, desc)
    oj = oracle.Join(desc)
    try:
        for bc in build_chunks:
            gj.append_build(bc)
            oj.append_build(bc)
        gj.build_finish()
        oj.build()
        if expect_method is not None:
            assert gj.info().method == expect_method
        out = gj.probe(probe_chunk)
        n = out.num_rows
        gpi, gbi = gj.probe_indexes(n)
        opi, obi = oj.probe_all(probe_chunk, cap=max(1024, 4 * probe_chunk.num_rows + 16))
        # exactly the reference's pair sequence: probe order, and inside a duplicate chain descending build index
        assert np.array_equal(gpi, opi) and np.array_equal(gbi, obi)
        gout = gpu.chunk_out_to_host(ctx, out)
        if desc.join_type in (abi.JOIN_RIGHT_SEMI, abi.JOIN_RIGHT_ANTI):
            assert n == 0 and len(opi) == 0          # these emit build rows only, all of them in the POST_PROBE phase
        else:
            oout = oj.output(probe_chunk, opi, obi)
            assert [s for s, _, _, _ in gout] == [s for s, _, _ in oout]
            grow = rows_sorted([col_to_py(t, d, nl) for _, t, d, nl in gout])
            orow = rows_sorted([col_to_py(desc_type(desc, oj, probe_chunk, s), d, nl) for s, d, nl in oout])
            assert grow == orow
        if desc.join_type in POST_PROBE_JOIN_TYPES:
            # a second prober marks more build rows, then POST_PROBE: never-matched (RIGHT SEMI: matched) build rows in
            # build order, probe columns NULL (probe_remain, join_hash_map.hpp:136-143,420-457)
            half = probe_chunk.num_rows // 2
            tail = Chunk([(sl, d[half:].copy(), None if nl is None else nl[half:].copy()) for sl, d, nl in probe_chunk.columns()])
            gj.probe(tail, prober_id=1)
            oj.probe_all(tail, cap=max(1024, 4 * tail.num_rows + 16))
            ptypes = [probe_chunk.types[probe_chunk.slots.index(desc.probe_out_slots[k])] for k in range(desc.num_probe_out)]
            grem = gpu.chunk_out_to_host(ctx, gj.probe_remain())
            orem = oj.probe_remain(ptypes)
            assert [s for s, _, _, _ in grem] == [s for s, _, _ in orem]
            for (gs, gt, gd, gn), (os_, od, on) in zip(grem, orem):
                exp_null = on if gn is not None else None
                assert col_to_py(gt, gd, gn) == col_to_py(gt, od, exp_null), f"remain slot {gs}"      # same rows, same (build) order
                if gn is None:
                    assert not on.any()
        return gj.info(), n
    finally:
        gj.close()


def desc_type(desc, oj, probe_chunk, slot):
    if slot in probe_chunk.slots:
        return probe_chunk.types[probe_chunk.slots.index(slot)]
    return oj.build_types[slot]


POST_PROBE_JOIN_TYPES = (abi.JOIN_RIGHT_OUTER, abi.JOIN_RIGHT_SEMI, abi.JOIN_RIGHT_ANTI, abi.JOIN_FULL_OUTER)
JOIN_TYPES = [abi.JOIN_INNER, abi.JOIN_LEFT_OUTER, abi.JOIN_LEFT_SEMI, abi.JOIN_LEFT_ANTI] + list(POST_PROBE_JOIN_TYPES)


@pytest.mark.parametrize("join_type", JOIN_TYPES)
@pytest.mark.parametrize("kind", ["range_unique", "range_dup_nulls", "sparse_i64", "tiny_i8", "small_i16", "two_keys", "wide_two_i64", "wide_i64_i32_i16"])
def test_join_parity(gpu, ctx, oracle, join_type, kind):
    rng = np.random.default_rng(11)
    nb, npr = 5000, 20011
    expect = None
    if kind == "range_unique":
        bkey = rng.permutation(np.arange(100, 100 + nb, dtype=np.int32))
        pkey = rng.integers(0, 100 + 2 * nb, npr, dtype=np.int32)
        bn = pn = None
        ktypes, expect = [abi.TYPE_INT], abi.JOIN_METHOD_RANGE_DIRECT_MAPPING
    elif kind == "range_dup_nulls":
        bkey = rng.integers(0, 800, nb, dtype=np.int32)
        pkey = rng.integers(-50, 900, npr, dtype=np.int32)
        bn, pn = rand_nulls(rng, nb), rand_nulls(rng, npr)
        ktypes, expect = [abi.TYPE_INT], abi.JOIN_METHOD_RANGE_DIRECT_MAPPING
    elif kind == "sparse_i64":
        pool = rng.integers(-2**62, 2**62, 3000, dtype=np.int64)
        pool[0] = np.iinfo(np.int64).min  # the table's EMPTY sentinel value is a legal key
        bkey = pool[rng.integers(0, 3000, nb)]
        pkey = np.concatenate([pool[rng.integers(0, 3000, npr // 2)], rng.integers(-2**62, 2**62, npr - npr // 2, dtype=np.int64)])
        bn, pn = None, rand_nulls(rng, npr, 0.02)
        ktypes, expect = [abi.TYPE_BIGINT], abi.JOIN_METHOD_LINEAR_CHAINED
    elif kind == "tiny_i8":
        bkey = rng.integers(-128, 128, 300).astype(np.int8)
        pkey = rng.integers(-128, 128, npr).astype(np.int8)
        bn = pn = None
        nb = 300
        ktypes, expect = [abi.TYPE_TINYINT], abi.JOIN_METHOD_DIRECT_MAPPING
    elif kind == "small_i16":
        bkey = rng.integers(-3000, 3000, nb).astype(np.int16)
        pkey = rng.integers(-4000, 4000, npr).astype(np.int16)
        bn, pn = rand_nulls(rng, nb, 0.05), None
        ktypes, expect = [abi.TYPE_SMALLINT], abi.JOIN_METHOD_DIRECT_MAPPING
    elif kind.startswith("wide"):
        # packed keys of 16 / 14 bytes (SERIALIZED_FIXED_SIZE_LARGEINT, join_hash_table.cpp:221-222): the table is keyed by a
        # fingerprint, every chain entry's full key is compared; duplicates, NULLs in any key column, keys that differ only in
        # the high column
        widths = [np.int64, np.int64] if kind == "wide_two_i64" else [np.int64, np.int32, np.int16]
        tys = {np.int64: abi.TYPE_BIGINT, np.int32: abi.TYPE_INT, np.int16: abi.TYPE_SMALLINT}
        pool = [rng.integers(-2**40, 2**40, 40).astype(w) if w == np.int64 else rng.integers(-100, 100, 12).astype(w) for w in widths]
        bcols = [p[rng.integers(0, len(p), nb)] for p in pool]
        pcols = [p[rng.integers(0, len(p), npr)] for p in pool]
        pcols[0][::7] = rng.integers(-2**62, 2**62, len(pcols[0][::7]))          # probe keys the build side never had
        bpay = rng.integers(0, 10**6, nb, dtype=np.int32)
        ppay = rng.integers(0, 10**9, npr, dtype=np.int64)
        bslots, pslots = [10, 13, 14][:len(widths)], [0, 3, 4][:len(widths)]
        d = abi.make_join_desc(join_type, bslots, pslots, [tys[w] for w in widths], build_out=[11, bslots[-1]], probe_out=[1, pslots[0]])
        bn1, pn1 = rand_nulls(rng, nb, 0.03), rand_nulls(rng, npr, 0.03)
        half = nb // 2
        def bchunk(lo, hi):
            return Chunk([(sl, c[lo:hi].copy(), bn1[lo:hi].copy() if k == 1 else None) for k, (sl, c) in enumerate(zip(bslots, bcols))] + [(11, bpay[lo:hi].copy(), None)])
        probe = Chunk([(sl, c, pn1 if k == 0 else None) for k, (sl, c) in enumerate(zip(pslots, pcols))] + [(1, ppay, None)])
        _join_pairs(gpu, ctx, oracle, d, [bchunk(0, half), bchunk(half, nb)], probe, expect_method=abi.JOIN_METHOD_LINEAR_CHAINED)
        return
    else:  # two int32 keys packed into one 8-byte key (SERIALIZED_FIXED_SIZE_BIGINT)
        bkey = rng.integers(0, 60, nb, dtype=np.int32)
        pkey = rng.integers(0, 70, npr, dtype=np.int32)
        bn = pn = None
        ktypes, expect = [abi.TYPE_INT, abi.TYPE_INT], abi.JOIN_METHOD_LINEAR_CHAINED
    bpay = rng.integers(0, 10**6, nb, dtype=np.int32)
    bpay2 = rng.normal(size=nb)
    ppay = rng.integers(0, 10**9, npr, dtype=np.int64)
    if kind == "two_keys":
        bkey2 = rng.integers(-5, 5, nb, dtype=np.int32)
        pkey2 = rng.integers(-6, 6, npr, dtype=np.int32)
        d = abi.make_join_desc(join_type, [10, 13], [0, 3], ktypes, build_out=[11, 12], probe_out=[0, 1])
        half = nb // 2
        builds = [Chunk([(10, bkey[:half].copy(), None), (13, bkey2[:half].copy(), None), (11, bpay[:half].copy(), None), (12, bpay2[:half].copy(), None)]),
                  Chunk([(10, bkey[half:].copy(), None), (13, bkey2[half:].copy(), None), (11, bpay[half:].copy(), None), (12, bpay2[half:].copy(), None)])]
        probe = Chunk([(0, pkey, None), (3, pkey2, None), (1, ppay, None)])
    else:
        d = abi.make_join_desc(join_type, [10], [0], ktypes, build_out=[11, 12, 10], probe_out=[0, 1])
        half = nb // 3
        builds = [Chunk([(10, bkey[:half].copy(), None if bn is None else bn[:half].copy()), (11, bpay[:half].copy(), None), (12, bpay2[:half].copy(), None)]),
                  Chunk([(10, bkey[half:].copy(), None if bn is None else bn[half:].copy()), (11, bpay[half:].copy(), None), (12, bpay2[half:].copy(), None)])]
        probe = Chunk([(0, pkey, pn), (1, ppay, None)])
    _join_pairs(gpu, ctx, oracle, d, builds, probe, expect_method=expect)


@pytest.mark.parametrize("join_type", JOIN_TYPES)
@pytest.mark.parametrize("shape", ["dups_int", "nullable_mixed_types"])
def test_join_other_conjunct_parity(gpu, ctx, oracle, join_type, shape):
    # other-join conjunct (exec/hash_joiner.h:314-329): key-matched pairs only count when the expression over the probe row
    # and the build row is true; every join type, duplicate chains, NULLs on either side of the expression (NULL = false)
    rng = np.random.default_rng(41)
    nb, npr = 3000, 15_001
    bkey = rng.integers(0, 500, nb, dtype=np.int32)
    pkey = rng.integers(-20, 560, npr, dtype=np.int32)
    if shape == "dups_int":
        bval, pval = rng.integers(0, 100, nb, dtype=np.int32), rng.integers(0, 100, npr, dtype=np.int32)
        bvn = pvn = None
        conj = [("col", 1), ("col", 11), "<"]
    else:
        bval, pval = rng.normal(size=nb), rng.integers(-3, 3, npr, dtype=np.int64)
        bvn, pvn = rand_nulls(rng, nb, 0.1), rand_nulls(rng, npr, 0.1)
        conj = [("col", 1), "todouble", ("col", 11), ">", ("col", 0), ("i", 100), "<", "or"]
    d = abi.make_join_desc(join_type, [10], [0], [abi.TYPE_INT], build_out=[11, 10], probe_out=[0, 1], other_conjunct=conj)
    builds = [Chunk([(10, bkey[:1000].copy(), None), (11, bval[:1000].copy(), None if bvn is None else bvn[:1000].copy())]),
              Chunk([(10, bkey[1000:].copy(), None), (11, bval[1000:].copy(), None if bvn is None else bvn[1000:].copy())])]
    probe = Chunk([(0, pkey, rand_nulls(rng, npr, 0.02)), (1, pval, pvn)])
    _, n = _join_pairs(gpu, ctx, oracle, d, builds, probe)
    if join_type in (abi.JOIN_INNER, abi.JOIN_LEFT_OUTER, abi.JOIN_LEFT_SEMI, abi.JOIN_LEFT_ANTI, abi.JOIN_FULL_OUTER):
        assert n > 0


def test_join_one_key_golden(gpu, ctx):
    # be/test/exec/join_hash_map_test.cpp:2042-2088 OneKeyJoinHashTable through the CUDA path
    d = abi.make_join_desc(abi.JOIN_INNER, [3], [0], [abi.TYPE_INT], build_out=[3, 4, 5], probe_out=[0, 1, 2])
    j = gpu.Join(ctx, d)
    try:
        j.append_build(Chunk([(3 + k, np.arange(10 * k, 10 * k + 10, dtype=np.int32), None) for k in range(3)]))
        j.build_finish()
        out = gpu.chunk_out_to_host(ctx, j.probe(Chunk([(k, np.arange(1 + 10 * k, 6 + 10 * k, dtype=np.int32), None) for k in range(3)])))
        assert len(out) == 6
        for k, (slot, _, data, _) in enumerate(out):
            assert slot == k and data.tolist() == list(range(1 + 10 * (k % 3), 6 + 10 * (k % 3)))
        first, nxt = j.copy_table()
        assert first.tolist() == list(range(1, 11)) and not nxt.any()   # first[key - min] = build row (1-based)
    finally:
        j.close()


def test_join_sql_golden_counts(gpu, ctx):
    # test/sql/test_join/R/test_join_range_direct_mapping (see tests/test_oracle_golden.py)
    from tests.test_oracle_golden import sql_golden_t1
    idx, null13, big, null14 = sql_golden_t1()
    for keyc, nulls, ktype, jt, expect in ((idx, None, abi.TYPE_INT, abi.JOIN_INNER, 1280000),
                                           (idx, null13, abi.TYPE_INT, abi.JOIN_INNER, 98461),
                                           (big, null14, abi.TYPE_BIGINT, abi.JOIN_LEFT_OUTER, 1280000)):
        j = gpu.Join(ctx, abi.make_join_desc(jt, [1], [0], [ktype], build_out=[1], probe_out=[0]))
        try:
            j.append_build(Chunk([(1, keyc, nulls)]))
            j.build_finish()
            out = j.probe(Chunk([(0, keyc, nulls)]))
            assert out.num_rows == expect
            if jt == abi.JOIN_LEFT_OUTER:
                _, bi = j.probe_indexes(out.num_rows)
                assert int((bi != 0).sum()) == 91428
        finally:
            j.close()
    w1 = np.concatenate([idx, idx])
    j = gpu.Join(ctx, abi.make_join_desc(abi.JOIN_INNER, [1], [0], [abi.TYPE_INT], build_out=[], probe_out=[0]))
    try:
        j.append_build(Chunk([(1, w1, None)]))
        j.build_finish()
        assert j.info().has_duplicates == 1
        assert j.probe(Chunk([(0, w1, None)])).num_rows == 5120000
    finally:
        j.close()


def test_join_empty_sides(gpu, ctx):
    d = abi.make_join_desc(abi.JOIN_INNER, [1], [0], [abi.TYPE_INT], build_out=[1], probe_out=[0])
    j = gpu.Join(ctx, d)
    try:
        j.append_build(Chunk([(1, np.empty(0, dtype=np.int32), None)]))
        j.build_finish()
        assert j.probe(Chunk([(0, np.arange(10, dtype=np.int32), None)])).num_rows == 0
        assert j.probe(Chunk([(0, np.empty(0, dtype=np.int32), None)])).num_rows == 0
    finally:
        j.close()
    j = gpu.Join(ctx, abi.make_join_desc(abi.JOIN_LEFT_ANTI, [1], [0], [abi.TYPE_INT], build_out=[], probe_out=[0]))
    try:
        j.append_build(Chunk([(1, np.empty(0, dtype=np.int32), None)]))
        j.build_finish()
        assert j.probe(Chunk([(0, np.arange(10, dtype=np.int32), None)])).num_rows == 10
    finally:
        j.close()


@pytest.mark.parametrize("join_type", [abi.JOIN_INNER, abi.JOIN_LEFT_OUTER, abi.JOIN_LEFT_SEMI, abi.JOIN_LEFT_ANTI])
def test_join_build_side_never_appended(gpu, ctx, join_type):
    # A dimension scan that filters everything out hands the build operator no chunk at all: build_finish without any
    # append.  With the build schema declared in the desc (the reference knows it from the plan's row descriptor,
    # exec/hash_joiner.h:66-133) the probe returns zero rows (INNER / SEMI), every probe row (ANTI), or every probe row
    # padded with typed NULL build columns (LEFT OUTER) -- not an error.
    d = abi.make_join_desc(join_type, [1], [0], [abi.TYPE_INT], build_out=[1, 2], probe_out=[0],
                           build_out_types=[abi.TYPE_INT, abi.TYPE_BIGINT])
    j = gpu.Join(ctx, d)
    try:
        j.build_finish()
        keys = np.arange(10, dtype=np.int32)
        out = j.probe(Chunk([(0, keys, None)]))
        semi = join_type in (abi.JOIN_LEFT_SEMI, abi.JOIN_LEFT_ANTI)
        assert out.num_rows == (0 if join_type in (abi.JOIN_INNER, abi.JOIN_LEFT_SEMI) else 10)
        assert out.num_cols == (1 if semi else 3)
        got = gpu.chunk_out_to_host(ctx, out)
        if not semi:
            assert [g[1] for g in got] == [abi.TYPE_INT, abi.TYPE_INT, abi.TYPE_BIGINT]
        if join_type == abi.JOIN_LEFT_OUTER:
            assert got[0][2].tolist() == keys.tolist()
            assert got[1][3].tolist() == [1] * 10 and got[2][3].tolist() == [1] * 10      # NULL-padded build columns
        if join_type == abi.JOIN_LEFT_ANTI:
            assert got[0][2].tolist() == keys.tolist()
    finally:
        j.close()
    # without declared types the old, loud failure stays
    j = gpu.Join(ctx, abi.make_join_desc(abi.JOIN_LEFT_OUTER, [1], [0], [abi.TYPE_INT], build_out=[1], probe_out=[0]))
    try:
        j.build_finish()
        with pytest.raises(gpu.GpuError) as ei:
            j.probe(Chunk([(0, np.arange(4, dtype=np.int32), None)]))
        assert ei.value.code == abi.SR_ERR_STATE
    finally:
        j.close()


def test_divide_by_zero_is_null_gpu(gpu, ctx, oracle):
    # ArithmeticRightZeroCheck (be/src/exprs/arithmetic_operation.h:638): x / 0 -> NULL, skipped by SUM / COUNT / MIN
    rng = np.random.default_rng(3)
    n = 100_003
    a = rng.integers(-50, 50, n, dtype=np.int64)
    b = rng.integers(-3, 4, n, dtype=np.int64)
    g = rng.integers(0, 7, n, dtype=np.int32)
    q = [("col", 0), ("col", 1), "/"]
    d = abi.make_agg_desc([2], [abi.TYPE_INT], fns=[(abi.AGG_SUM, abi.TYPE_DOUBLE, 10, q), (abi.AGG_COUNT, abi.TYPE_DOUBLE, 11, q),
                                                     (abi.AGG_MIN, abi.TYPE_DOUBLE, 12, q), (abi.AGG_COUNT_STAR, 0, 13, None)], ranges=[(0, 6)])
    ch = Chunk([(0, a, None), (1, b, None), (2, g, None)])
    ga, oa = gpu.Agg(ctx, d), oracle.Agg(d)
    try:
        ga.push(ch)
        oa.push(ch)
        rows = gpu_rows(ga.result())
        assert_rows_equal(rows, oracle_rows(oa), float_cols=(1,))
        assert sum(r[2] for r in rows) == int((b != 0).sum())      # COUNT(a / b) counts the rows with a non-zero divisor
    finally:
        ga.close()


# ---------------------------------------------------------------------------------------------
# hash aggregate (K13-K17)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("typ,np_t", [(abi.TYPE_SMALLINT, np.int16), (abi.TYPE_INT, np.int32), (abi.TYPE_BIGINT, np.int64),
                                      (abi.TYPE_FLOAT, np.float32), (abi.TYPE_DOUBLE, np.float64)])
def test_agg_sum_golden(gpu, ctx, typ, np_t):
    # be/test/exprs/agg/aggregate_test.cpp:61-83 test_sum: 524076 / 2499500 / merged 3023576
    col1 = np.array(list(range(1024)) + [100, 200], dtype=np_t)
    col2 = np.arange(2000, 3000, dtype=np_t)
    d = abi.make_agg_desc(fns=[(abi.AGG_SUM, typ, 10, [("col", 0)])])
    a1, a2 = gpu.Agg(ctx, d), gpu.Agg(ctx, d)
    try:
        a1.push(Chunk([(0, col1, None)]))
        a2.push(Chunk([(0, col2, None)]))
        a2.merge(a1)
        assert a1.result()[0][2][0] == 524076
        assert a2.result()[0][2][0] == 3023576
    finally:
        a1.close()
        a2.close()


def _agg_case(name, n, rng):
    k1 = rng.integers(0, 7, n, dtype=np.int32)
    k2 = rng.integers(-3, 4, n).astype(np.int16)
    k3 = rng.integers(-10**9, 10**9, n, dtype=np.int64)
    v32 = rng.integers(-10**6, 10**6, n, dtype=np.int32)
    v64 = rng.integers(-10**15, 10**15, n, dtype=np.int64)
    vd = rng.normal(100, 50, n)
    vf = rng.random(n).astype(np.float32)
    dec = rng.integers(-2**62, 2**62, n, dtype=np.int64)
    if "lowcard" in name:
        k3 = k3 % 5
    if "full16" in name:          # the all-ones 16-byte key (-1, -1) collides with the empty marker: special slot
        k3[::97] = -1
        v64[::97] = -1
        k3[1::2] = k3[:-1:2][:len(k3[1::2])]
        v64[1::2] = v64[:-1:2][:len(v64[1::2])]
    cols = [(0, k1, rand_nulls(rng, n, 0.05) if "nullkey" in name else None), (1, k2, None), (2, k3, None),
            (3, v32, rand_nulls(rng, n, 0.2) if "nullval" in name else None), (4, v64, None), (5, vd, None), (6, vf, None),
            (7, dec, None, abi.TYPE_DECIMAL64)]
    fns = [(abi.AGG_SUM, abi.TYPE_INT, 20, [("col", 3)]), (abi.AGG_COUNT, abi.TYPE_INT, 21, [("col", 3)]),
           (abi.AGG_COUNT_STAR, abi.TYPE_INT, 22, None), (abi.AGG_AVG, abi.TYPE_INT, 23, [("col", 3)]),
           (abi.AGG_MIN, abi.TYPE_BIGINT, 24, [("col", 4)]), (abi.AGG_MAX, abi.TYPE_INT, 25, [("col", 3)]),
           (abi.AGG_SUM, abi.TYPE_DOUBLE, 26, [("col", 5)]), (abi.AGG_SUM, abi.TYPE_DECIMAL64, 27, [("col", 7)])]
    fns2 = [(abi.AGG_SUM, abi.TYPE_BIGINT, 30, [("col", 3), ("col", 4), "*", ("i", 3), "+"]),
            (abi.AGG_MIN, abi.TYPE_DOUBLE, 31, [("col", 5)]), (abi.AGG_MAX, abi.TYPE_FLOAT, 32, [("col", 6)]),
            (abi.AGG_AVG, abi.TYPE_DOUBLE, 33, [("col", 5), ("col", 6), "*"])]
    if name.startswith("nogroup"):
        d = abi.make_agg_desc(fns=fns if "a" in name.split("_")[1] else fns2)
        fl = (3, 6) if "a" in name.split("_")[1] else (1, 3)
    elif name.startswith("dense"):
        d = abi.make_agg_desc([0, 1], [abi.TYPE_INT, abi.TYPE_SMALLINT], fns=fns if "_a" in name else fns2,
                              ranges=[(0, 6), (-3, 3)], group_nullable=[1 if "nullkey" in name else 0, 0])
        fl = (2 + 3, 2 + 6) if "_a" in name else (2 + 1, 2 + 3)
    elif name.startswith("wide_full16"):   # two int64 keys: exactly 16 packed bytes (128-bit CAS claim)
        d = abi.make_agg_desc([2, 4], [abi.TYPE_BIGINT, abi.TYPE_BIGINT], fns=fns2)
        fl = (2 + 1, 2 + 3)
    elif name.startswith("wide"):          # int32 + int64 + int16 (+ null flags): 14..15 packed bytes
        d = abi.make_agg_desc([0, 2, 1], [abi.TYPE_INT, abi.TYPE_BIGINT, abi.TYPE_SMALLINT], fns=fns,
                              group_nullable=[1 if "nullkey" in name else 0, 0, 0])
        fl = (3 + 3, 3 + 6)
    elif name.startswith("hash2"):
        d = abi.make_agg_desc([0, 1], [abi.TYPE_INT, abi.TYPE_SMALLINT], fns=fns, group_nullable=[1 if "nullkey" in name else 0, 0])
        fl = (2 + 3, 2 + 6)
    else:  # hash on a high-cardinality int64 key
        d = abi.make_agg_desc([2], [abi.TYPE_BIGINT], fns=fns2)
        fl = (1 + 1, 1 + 3)
    return d, cols, fl


@pytest.mark.parametrize("name", ["nogroup_a", "nogroup_b", "nogroup_a_nullval", "dense_a", "dense_b", "dense_a_nullkey_nullval",
                                  "hash2_a", "hash2_a_nullkey_nullval", "hash1_b", "wide_a", "wide_lowcard_nullkey_nullval",
                                  "wide_full16"])
@pytest.mark.parametrize("n", [0, 1, 1000, 70001])
def test_agg_parity(gpu, ctx, oracle, name, n):
    rng = np.random.default_rng(5)
    d, cols, fl = _agg_case(name, n, rng)
    # push in two chunks to exercise state carry-over
    cut = n // 3
    def sub(lo, hi):
        return Chunk([(c[0], c[1][lo:hi].copy(), None if c[2] is None else c[2][lo:hi].copy()) + tuple(c[3:]) for c in cols])
    ga, oa = gpu.Agg(ctx, d), oracle.Agg(d)
    try:
        for lo, hi in ((0, cut), (cut, n)):
            ch = sub(lo, hi)
            ga.push(ch)
            oa.push(ch)
        assert_rows_equal(gpu_rows(ga.result()), oracle_rows(oa), float_cols=fl)
    finally:
        ga.close()


@pytest.mark.parametrize("shape", ["nogroup", "dense", "hash", "hash_nullkey", "wide"])
@pytest.mark.parametrize("n", [0, 1, 5000, 400_001, 2_500_000])
def test_agg_count_distinct_parity(gpu, ctx, oracle, shape, n):
    # COUNT(DISTINCT col) next to ordinary functions (reference: distinct.h:48-62,410-424,590-594: a hash set per group
    # state, NULL inputs skipped, result = set size).  2.5 M rows: the (group, value) set is filled by the
    # radix-partitioned push; two COUNT(DISTINCT)s of different widths share one operator.
    rng = np.random.default_rng(77)
    cols = [(0, rng.integers(0, 40, n, dtype=np.int32), rand_nulls(rng, n, 0.03) if shape == "hash_nullkey" else None),
            (1, rng.integers(-3000, 3000, n, dtype=np.int64), rand_nulls(rng, n, 0.1)),
            (2, rng.integers(0, 1 << 40, n, dtype=np.int64) >> rng.integers(0, 30, n), None),
            (3, rng.integers(0, 2000, n, dtype=np.int16), None),
            (4, rng.integers(0, 100_000, n, dtype=np.int32), None)]
    fns = [(abi.AGG_COUNT_DISTINCT, abi.TYPE_BIGINT, 20, [("col", 1)]), (abi.AGG_SUM, abi.TYPE_BIGINT, 21, [("col", 1)]),
           (abi.AGG_COUNT_STAR, 0, 22, None), (abi.AGG_COUNT_DISTINCT, abi.TYPE_SMALLINT, 23, [("col", 3)]),
           (abi.AGG_COUNT, abi.TYPE_BIGINT, 24, [("col", 1)])]
    if shape == "nogroup":
        d = abi.make_agg_desc(fns=fns)
    elif shape == "dense":
        d = abi.make_agg_desc([0], [abi.TYPE_INT], fns=fns, ranges=[(0, 39)])
    elif shape == "wide":    # 4 + 4 key bytes + 8 value bytes = a 16-byte set key
        d = abi.make_agg_desc([0, 4], [abi.TYPE_INT, abi.TYPE_INT], fns=[(abi.AGG_COUNT_DISTINCT, abi.TYPE_BIGINT, 20, [("col", 2)]), (abi.AGG_COUNT_STAR, 0, 22, None)])
    else:
        d = abi.make_agg_desc([0], [abi.TYPE_INT], fns=fns, group_nullable=[1 if shape == "hash_nullkey" else 0])
    ga, oa = gpu.Agg(ctx, d), oracle.Agg(d)
    try:
        for lo, hi in ((0, n // 3), (n // 3, n)):
            ch = Chunk([(c[0], c[1][lo:hi].copy(), None if c[2] is None else c[2][lo:hi].copy()) for c in cols])
            ga.push(ch)
            oa.push(ch)
        got = gpu_rows(ga.result())
        assert_rows_equal(got, oracle_rows(oa))
        if shape == "nogroup" and n >= 5000:
            v = cols[1][1][cols[1][2] == 0]
            assert got[0][0] == len(np.unique(v)) and got[0][4] == len(v)
        # reset -> same operator, same answer (the sets are reset with it)
        ga.reset()
        ch = Chunk([(c[0], c[1], c[2]) for c in cols])
        ga.push(ch)
        assert_rows_equal(gpu_rows(ga.result()), oracle_rows(oa))
    finally:
        ga.close()


def test_agg_count_distinct_unsupported_forms(gpu, ctx):
    fn = [(abi.AGG_COUNT_DISTINCT, abi.TYPE_BIGINT, 20, [("col", 1), ("i", 1), "+"])]          # expression argument
    with pytest.raises(gpu.GpuError) as ei:
        gpu.Agg(ctx, abi.make_agg_desc(fns=fn))
    assert ei.value.code == abi.SR_ERR_NOT_SUPPORTED
    d = abi.make_agg_desc([0], [abi.TYPE_INT], fns=[(abi.AGG_COUNT_DISTINCT, abi.TYPE_BIGINT, 20, [("col", 1)])])
    a, b = gpu.Agg(ctx, d), gpu.Agg(ctx, d)
    try:
        ch = Chunk([(0, np.arange(10, dtype=np.int32), None), (1, np.arange(10, dtype=np.int64), None)])
        a.push(ch)
        b.push(ch)
        b.finish()
        with pytest.raises(gpu.GpuError) as ei:
            a.merge(b)
        assert ei.value.code == abi.SR_ERR_NOT_SUPPORTED
    finally:
        a.close()
        b.close()


@pytest.mark.parametrize("name", ["dense_b", "nogroup_b"])
def test_agg_dense_state_elementwise_merge(gpu, ctx, oracle, name):
    # two "fragment instances" aggregate disjoint halves; their dense tables are merged in place by the element-wise
    # reduction an all-reduce would apply (sr_agg_dense_state), then either instance yields the final result
    import torch
    from starrocks_b200.distributed import dense_state_views
    rng = np.random.default_rng(11)
    n = 50_001
    d, cols, fl = _agg_case(name, n, rng)
    def sub(lo, hi):
        return Chunk([(c[0], c[1][lo:hi].copy(), None if c[2] is None else c[2][lo:hi].copy()) + tuple(c[3:]) for c in cols])
    a, b, oa = gpu.Agg(ctx, d), gpu.Agg(ctx, d), oracle.Agg(d)
    try:
        a.push(sub(0, n // 2))
        b.push(sub(n // 2, n))
        oa.push(sub(0, n))
        ctx.sync()
        va, vb = dense_state_views(a.dense_state(), "cuda:0"), dense_state_views(b.dense_state(), "cuda:0")
        assert len(va) == len(vb) >= 2
        for (ta, ra), (tb, rb) in zip(va, vb):
            assert ra == rb and ta.shape == tb.shape and ta.dtype == tb.dtype
            merged = ta + tb if ra == abi.STATE_REDUCE_SUM else (torch.minimum(ta, tb) if ra == abi.STATE_REDUCE_MIN else torch.maximum(ta, tb))
            ta.copy_(merged)
        torch.cuda.synchronize()
        assert_rows_equal(gpu_rows(a.result()), oracle_rows(oa), float_cols=fl)
    finally:
        a.close()
        b.close()


@pytest.mark.parametrize("name", ["dense_a", "hash1_b"])
def test_agg_dense_state_refused_when_not_elementwise(gpu, ctx, name):
    rng = np.random.default_rng(12)
    d, cols, _ = _agg_case(name, 1000, rng)
    a = gpu.Agg(ctx, d)
    try:
        a.push(Chunk([(c[0], c[1], c[2]) + tuple(c[3:]) for c in cols]))
        with pytest.raises(gpu.GpuError) as ei:
            a.dense_state()
        assert ei.value.code == abi.SR_ERR_NOT_SUPPORTED     # 128-bit decimal sum / hash table
    finally:
        a.close()


def test_agg_hash_growth_two_pass(gpu, ctx, oracle):
    # 3 M rows, ~2.6 M distinct int64 keys: forces the find/insert + update two-pass path and a table growth
    rng = np.random.default_rng(9)
    n = 3_000_000
    keys = rng.integers(0, 20_000_000, n, dtype=np.int64)
    vals = rng.integers(0, 1000, n, dtype=np.int64)
    d = abi.make_agg_desc([0], [abi.TYPE_BIGINT], fns=[(abi.AGG_SUM, abi.TYPE_BIGINT, 10, [("col", 1)]),
                                                        (abi.AGG_COUNT_STAR, abi.TYPE_BIGINT, 11, None)])
    ch = Chunk([(0, keys, None), (1, vals, None)])
    ga, oa = gpu.Agg(ctx, d), oracle.Agg(d)
    try:
        ga.push(ch)
        ga.push(ch)
        oa.push(ch)
        oa.push(ch)
        res = ga.result()
        order = np.argsort(res[0][2], kind="stable")
        oo = oa.output()
        oorder = np.argsort(oo[0][1], kind="stable")
        assert len(res[0][2]) == len(oo[0][1]) == len(np.unique(keys))
        for k in range(3):
            assert np.array_equal(res[k][2][order], oo[k][1][oorder])
    finally:
        ga.close()


@pytest.mark.parametrize("wide,expected,nkeys,n,force_l2", [(False, 0, 300_000, 700_000, False), (True, 0, 300_000, 700_000, False),
                                                             (False, 0, 5_000_000, 1_700_000, False), (False, 16_000_000, 3_000_000, 5_000_000, False),
                                                             (False, 0, 40_000_000, 6_000_000, False), (True, 0, 300_000, 700_000, True),
                                                             (False, 0, 5_000_000, 1_700_000, True)])
def test_agg_partitioned_push_parity(gpu, ctx, oracle, monkeypatch, wide, expected, nkeys, n, force_l2):
    # Large batches into a hash table of more than one slice take the radix-partitioned push (sr_agg_part.cuh: bucket
    # histogram, one or two scatter levels, one CTA per table slice applying its bucket in shared memory); the row
    # threshold is lowered through the library's tuning knob so that small inputs reach it.  A following small batch takes
    # the direct path.  Fourth case: 2^15 buckets -> two scatter levels.  Fifth: ~5.6 M groups into the initial 2^21-slot
    # table: every slice overflows, the buckets are re-applied with global atomics after a growth.  force_l2: the fallback
    # for tables of more than 2^15 slices (buckets of several slices, global atomics on an L2-prefetched range).
    monkeypatch.setenv("SR_AGG_PARTITION_MIN_ROWS", "100000")
    if force_l2:
        monkeypatch.setenv("SR_AGG_PARTITION_FORCE_L2", "1")
    rng = np.random.default_rng(21)
    k64 = rng.integers(0, nkeys, n, dtype=np.int64) * 1_000_003 - 7
    k64[::1000] = -1            # the table's empty marker as a legal key (special slot)
    k32 = rng.integers(0, 3, n, dtype=np.int32)
    v = rng.integers(-1000, 1000, n, dtype=np.int64)
    vn = rand_nulls(rng, n, 0.1)
    fns = [(abi.AGG_SUM, abi.TYPE_BIGINT, 10, [("col", 2), ("i", 3), "*"]), (abi.AGG_COUNT_STAR, 0, 11, None),
           (abi.AGG_MIN, abi.TYPE_BIGINT, 12, [("col", 2)]), (abi.AGG_COUNT, abi.TYPE_BIGINT, 13, [("col", 2)])]
    if wide:
        d = abi.make_agg_desc([1, 0], [abi.TYPE_INT, abi.TYPE_BIGINT], fns=fns, expected_groups=expected)
    else:
        d = abi.make_agg_desc([0], [abi.TYPE_BIGINT], fns=fns, expected_groups=expected)
    def sub(lo, hi):
        return Chunk([(0, k64[lo:hi].copy(), None), (1, k32[lo:hi].copy(), None), (2, v[lo:hi].copy(), vn[lo:hi].copy())])
    ga, oa = gpu.Agg(ctx, d), oracle.Agg(d)
    try:
        launches0 = ctx.launches
        for lo, hi in ((0, n // 2), (n // 2, n - 50_000), (n - 50_000, n)):   # fresh table, loaded slices, direct push
            ga.push(sub(lo, hi))
            oa.push(sub(lo, hi))
        assert ctx.launches - launches0 >= 5          # 2 x (scatter [+ tiles + scatter] + apply) + the direct push
        assert_rows_equal(gpu_rows(ga.result()), oracle_rows(oa))
    finally:
        ga.close()


@pytest.mark.parametrize("shape", ["sum_count", "count_only", "decimal128", "double_minmax"])
def test_agg_partitioned_push_shapes(gpu, ctx, oracle, monkeypatch, shape):
    # record shapes of the partitioned push: the SIMPLE plan (one plain key, plain column inputs: the 1e9-row / 1e8-key
    # config of BASELINE.json), a key-only record, 128-bit sums (multi-word shared-memory adds with carries) and
    # double SUM / MIN / MAX (CAS-loop shared atomics); int32 key -> masked to its width in the record
    monkeypatch.setenv("SR_AGG_PARTITION_MIN_ROWS", "100000")
    rng = np.random.default_rng(5)
    n = 900_000
    k64 = rng.integers(0, 200_000, n, dtype=np.int64) * 7919 - 3
    k32 = rng.integers(-100_000, 100_000, n, dtype=np.int32)
    v = rng.integers(-(1 << 40), 1 << 40, n, dtype=np.int64)
    dv = rng.normal(0, 1e6, n)
    if shape == "sum_count":
        d = abi.make_agg_desc([0], [abi.TYPE_BIGINT], fns=[(abi.AGG_SUM, abi.TYPE_BIGINT, 10, [("col", 2)]), (abi.AGG_COUNT_STAR, 0, 11, None)])
        cols = [(0, k64, None), (2, v, None)]
    elif shape == "count_only":
        d = abi.make_agg_desc([1], [abi.TYPE_INT], fns=[(abi.AGG_COUNT_STAR, 0, 11, None)])
        cols = [(1, k32, None)]
    elif shape == "decimal128":
        d = abi.make_agg_desc([1], [abi.TYPE_INT], fns=[(abi.AGG_SUM, abi.TYPE_DECIMAL64, 10, [("col", 2)]), (abi.AGG_COUNT_STAR, 0, 11, None)])
        cols = [(1, k32, None), (2, v, None, abi.TYPE_DECIMAL64)]
    else:
        d = abi.make_agg_desc([0], [abi.TYPE_BIGINT], fns=[(abi.AGG_SUM, abi.TYPE_DOUBLE, 10, [("col", 3)]), (abi.AGG_MIN, abi.TYPE_DOUBLE, 11, [("col", 3)]),
                                                            (abi.AGG_MAX, abi.TYPE_DOUBLE, 12, [("col", 3)]), (abi.AGG_AVG, abi.TYPE_BIGINT, 13, [("col", 2)])])
        cols = [(0, k64, None), (2, v >> 20, None), (3, dv, None)]
    ga, oa = gpu.Agg(ctx, d), oracle.Agg(d)
    try:
        half = n // 2
        for lo, hi in ((0, half), (half, n)):
            ch = Chunk([(c[0], c[1][lo:hi].copy(), None) + tuple(c[3:]) for c in cols])
            ga.push(ch)
            oa.push(ch)
        assert_rows_equal(gpu_rows(ga.result()), oracle_rows(oa), float_cols=(1, 4) if shape == "double_minmax" else ())
    finally:
        ga.close()


def test_agg_pull_paging_and_device_output(gpu, ctx):
    keys = np.arange(10000, dtype=np.int32)
    d = abi.make_agg_desc([0], [abi.TYPE_INT], fns=[(abi.AGG_SUM, abi.TYPE_INT, 10, [("col", 0)])], ranges=[(0, 9999)])
    a = gpu.Agg(ctx, d)
    try:
        a.push(Chunk([(0, keys, None)]))
        a.finish()
        assert a.num_groups == 10000
        seen = []
        while True:
            out = a.pull(max_rows=4096, mem=abi.MEM_DEVICE)   # chunk_size paging like AggregateBlockingSourceOperator
            if out.num_rows == 0:
                break
            got = gpu.chunk_out_to_host(ctx, out)
            assert got[0][2].tolist() == got[1][2].tolist()
            seen += got[0][2].tolist()
        assert seen == list(range(10000))
    finally:
        a.close()


# ---------------------------------------------------------------------------------------------
# fused fragment: SSB Q4.1 and Q1.1 shapes against the chunk-at-a-time oracle pipeline
# ---------------------------------------------------------------------------------------------
def _run_q41(gpu, ctx, oracle, sf, n, pushes=1, mode=0, pinned=False):
    import torch
    dims = ssb.gen_dims(sf)
    lo = ssb.gen_lineorder(sf, n=n)
    gjoins, gkeep = ssb.build_dims(gpu, dims, ssb.dim_plans_q41(), ctx=ctx)
    ojoins, okeep = ssb.build_dims(oracle, dims, ssb.dim_plans_q41())
    agg_desc = ssb.q41_agg_desc()
    frag = gpu.Fragment(ctx, abi.ScanDesc(), gjoins, agg_desc, mode=mode)
    try:
        step = (n + pushes - 1) // pushes
        for p in range(pushes):
            part = {k: v[p * step:(p + 1) * step].copy() for k, v in lo.items()}
            if pinned:
                # page-locked host columns read in place by the fragment kernels (SR_MEM_HOST_PINNED)
                pin = {k: torch.from_numpy(part[k]).pin_memory() for k in ssb.Q41_FACT_COLS}
                rows = len(part[ssb.Q41_FACT_COLS[0]])
                frag.push(Chunk([(ssb.LO_SLOTS[k], pin[k].data_ptr(), None, abi.TYPE_INT) for k in ssb.Q41_FACT_COLS],
                                num_rows=rows, mem=abi.MEM_HOST_PINNED))
                ctx.sync()   # the buffers must outlive the asynchronous passes
                del pin
            else:
                frag.push(ssb.fact_chunk(part, ssb.Q41_FACT_COLS))
        got = gpu_rows(frag.agg.result())
        passed = frag.rows_passed
        ores, opassed = oracle.fragment_run(abi.ScanDesc(), ojoins, agg_desc, ssb.fact_chunk(lo, ssb.Q41_FACT_COLS), num_threads=4)
        assert passed == opassed
        assert_rows_equal(got, oracle_rows(ores))
        return got, passed
    finally:
        frag.close()
        for j, _, _ in gjoins:
            j.close()
        for item in gkeep:
            item[0].close()


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_fragment_q41_parity(gpu, ctx, oracle, mode):
    got, passed = _run_q41(gpu, ctx, oracle, sf=0.1, n=600_000, mode=mode)
    assert len(got) == 35 and passed > 0          # 7 years x 5 AMERICA nations
    assert sum(1 for _ in got) == 35


@pytest.mark.parametrize("mode", [1, 2])
def test_fragment_q41_multi_push_and_ragged(gpu, ctx, oracle, mode):
    _run_q41(gpu, ctx, oracle, sf=0.05, n=300_007, pushes=3, mode=mode)
    _run_q41(gpu, ctx, oracle, sf=0.01, n=5, pushes=1, mode=mode)


@pytest.mark.parametrize("mode", [1, 2])
def test_fragment_q41_pinned_host_columns_in_place(gpu, ctx, oracle, mode):
    _run_q41(gpu, ctx, oracle, sf=0.1, n=600_000, pushes=2, mode=mode, pinned=True)


def test_fragment_rejects_pageable_memory_tagged_pinned(gpu, ctx):
    lo = ssb.gen_lineorder(0.01, n=1000)
    frag = gpu.Fragment(ctx, abi.ScanDesc(), [], ssb.q11_agg_desc())
    try:
        ch = ssb.fact_chunk(lo, ssb.Q11_FACT_COLS, mem=abi.MEM_HOST_PINNED)
        with pytest.raises(gpu.GpuError) as ei:
            frag.push(ch)
        assert ei.value.code == abi.SR_ERR_INVALID_ARGUMENT
    finally:
        frag.close()


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_fragment_q11_parity(gpu, ctx, oracle, mode):
    lo = ssb.gen_lineorder(1, n=1_000_003)
    sd = abi.ScanDesc(preds=ssb.q11_scan_preds())
    agg_desc = ssb.q11_agg_desc()
    frag = gpu.Fragment(ctx, sd, [], agg_desc, mode=mode)
    try:
        ch = ssb.fact_chunk(lo, ssb.Q11_FACT_COLS)
        frag.push(ch)
        got = gpu_rows(frag.agg.result())
        ores, opassed = oracle.fragment_run(sd, [], agg_desc, ch, num_threads=2)
        assert frag.rows_passed == opassed
        assert_rows_equal(got, oracle_rows(ores))
        m = ((lo["lo_orderdate"] >= 19930101) & (lo["lo_orderdate"] <= 19931231) & (lo["lo_discount"] >= 1) &
             (lo["lo_discount"] <= 3) & (lo["lo_quantity"] < 25))
        assert got[0][0] == int((lo["lo_extendedprice"][m].astype(np.int64) * lo["lo_discount"][m]).sum())
    finally:
        frag.close()


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_fragment_hash_group_and_semi_join(gpu, ctx, oracle, mode):
    # group-by without declared ranges -> hash table path inside the fused kernel; a LEFT SEMI join with a
    # sparse (hash) build side; a generic filter expression on the fact table
    rng = np.random.default_rng(3)
    n = 400_000
    fact = Chunk([(0, rng.integers(0, 5000, n, dtype=np.int32), None), (1, rng.integers(0, 10**9, n, dtype=np.int64), rand_nulls(rng, n, 0.01)),
                  (2, rng.integers(0, 100, n, dtype=np.int32), None), (3, rng.integers(1, 1000, n, dtype=np.int32), None)])
    dim1 = Chunk([(10, np.arange(0, 5000, 2, dtype=np.int32), None), (11, rng.integers(0, 1000, 2500, dtype=np.int32), None)])
    dim2 = Chunk([(20, rng.integers(0, 10**9, 300_000, dtype=np.int64), None)])
    d1 = abi.make_join_desc(abi.JOIN_INNER, [10], [0], [abi.TYPE_INT], build_out=[11])
    d2 = abi.make_join_desc(abi.JOIN_LEFT_SEMI, [20], [1], [abi.TYPE_BIGINT])
    sd = abi.ScanDesc(filter_exprs=[abi.make_expr([("col", 2), ("col", 3), "+", ("i", 50), ">"])])
    agg_desc = abi.make_agg_desc([11, 2], [abi.TYPE_INT, abi.TYPE_INT],
                                 fns=[(abi.AGG_SUM, abi.TYPE_BIGINT, 30, [("col", 3), ("col", 11), "*"]), (abi.AGG_COUNT_STAR, 0, 31, None),
                                      (abi.AGG_MAX, abi.TYPE_BIGINT, 32, [("col", 1)])])
    gj1, gj2 = gpu.Join(ctx, d1), gpu.Join(ctx, d2)
    oj1, oj2 = oracle.Join(d1), oracle.Join(d2)
    frag = None
    try:
        for j, ch in ((gj1, dim1), (gj2, dim2), (oj1, dim1), (oj2, dim2)):
            j.append_build(ch)
        gj1.build_finish()
        gj2.build_finish()
        oj1.build()
        oj2.build()
        assert gj2.info().method == abi.JOIN_METHOD_LINEAR_CHAINED
        frag = gpu.Fragment(ctx, sd, [(gj1, 0, [11]), (gj2, 1, [])], agg_desc, mode=mode)
        frag.push(fact)
        got = gpu_rows(frag.agg.result())
        ores, opassed = oracle.fragment_run(sd, [(oj1, 0, [11]), (oj2, 1, [])], agg_desc, fact, num_threads=3)
        assert frag.rows_passed == opassed
        assert_rows_equal(got, oracle_rows(ores))
    finally:
        if frag:
            frag.close()
        gj1.close()
        gj2.close()


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_fragment_tpch_q3_parity(gpu, ctx, oracle, mode):
    # TPC-H Q3 shape: customer -> J1, orders SEMI J1 -> J2 (payload), lineitem INNER J2 -> group by a 12-byte key
    # (l_orderkey, o_orderdate, o_shippriority), SUM(decimal64 expression) -> decimal128
    from starrocks_b200 import tpch
    t = tpch.gen_tables(0.1)
    li = t["lineitem"]
    gj2, gkeep = tpch.q3_build_gpu(gpu, ctx, t)
    oj2, okeep = tpch.q3_build_oracle(oracle, t)
    assert gj2.info().build_rows == oj2.build_rows > 0
    _, _, _, _, li_scan = tpch.q3_descs()
    agg_desc = tpch.q3_agg_desc()
    payload = [tpch.O_ORDERDATE, tpch.O_SHIPPRIORITY]
    frag = gpu.Fragment(ctx, li_scan, [(gj2, tpch.L_ORDERKEY, payload)], agg_desc, mode=mode)
    try:
        n = len(li["l_orderkey"])
        for lo, hi in ((0, n // 2), (n // 2, n)):
            frag.push(tpch.table_chunk(li, tpch.LINEITEM_COLS, rows=(lo, hi)))
        got = gpu_rows(frag.agg.result())
        ores, opassed = oracle.fragment_run(li_scan, [(oj2, tpch.L_ORDERKEY, payload)], agg_desc,
                                            tpch.table_chunk(li, tpch.LINEITEM_COLS), num_threads=3)
        assert frag.rows_passed == opassed > 0
        assert_rows_equal(got, oracle_rows(ores))
        # independent restatement with numpy on a handful of groups
        cust_ok = set(t["customer"]["c_custkey"][t["customer"]["c_mktsegment"] == tpch.BUILDING].tolist())
        o = t["orders"]
        okeys = {int(k): int(d) for k, c, d in zip(o["o_orderkey"], o["o_custkey"], o["o_orderdate"]) if d < tpch.CUTOFF and int(c) in cust_ok}
        m = (li["l_shipdate"] > tpch.CUTOFF) & np.isin(li["l_orderkey"], np.fromiter(okeys.keys(), dtype=np.int32))
        rev = {}
        for k, p, d in zip(li["l_orderkey"][m].tolist(), li["l_extendedprice"][m].tolist(), li["l_discount"][m].tolist()):
            rev[k] = rev.get(k, 0) + p * (100 - d)
        assert len(got) == len(rev)
        for row in got[:50]:
            assert row[1] == okeys[row[0]] and row[2] == 0 and row[3] == rev[row[0]]
    finally:
        frag.close()
        gj2.close()
        for x in gkeep:
            if hasattr(x, "close"):
                x.close()


def test_fragment_tpch_q3_is_stable_over_repeated_runs(gpu, ctx, oracle):
    # regression: the final pass used its shared copy of the join descriptors before every warp had written it (missing
    # barrier) -- about one run in ten looked a payload up through an all-zero descriptor and produced a (key, 0, 0) group
    from starrocks_b200 import tpch
    t = tpch.gen_tables(0.1)
    li = t["lineitem"]
    gj2, gkeep = tpch.q3_build_gpu(gpu, ctx, t)
    oj2, okeep = tpch.q3_build_oracle(oracle, t)
    _, _, _, _, li_scan = tpch.q3_descs()
    agg_desc = tpch.q3_agg_desc()
    payload = [tpch.O_ORDERDATE, tpch.O_SHIPPRIORITY]
    try:
        ores, _ = oracle.fragment_run(li_scan, [(oj2, tpch.L_ORDERKEY, payload)], agg_desc, tpch.table_chunk(li, tpch.LINEITEM_COLS), num_threads=3)
        exp = oracle_rows(ores)
        n = len(li["l_orderkey"])
        for it in range(25):
            frag = gpu.Fragment(ctx, li_scan, [(gj2, tpch.L_ORDERKEY, payload)], agg_desc, mode=(0, 2)[it & 1])
            try:
                for lo, hi in ((0, n // 2), (n // 2, n)):
                    frag.push(tpch.table_chunk(li, tpch.LINEITEM_COLS, rows=(lo, hi)))
                assert gpu_rows(frag.agg.result()) == exp, f"run {it}"
            finally:
                frag.close()
    finally:
        gj2.close()
        for x in gkeep:
            if hasattr(x, "close"):
                x.close()


def test_fragment_hash_table_grows_when_the_sampled_estimate_is_wrong(gpu, ctx, oracle):
    # the plan samples the first 64 K rows: none of them passes the conjunct, so the hash table is sized for ~0 groups;
    # the remaining 2.9 M rows all pass and all are new groups -> refused rows are re-applied after the table has grown
    n = 3_000_000
    idx = np.arange(n, dtype=np.int32)
    rng = np.random.default_rng(31)
    fact = Chunk([(0, idx, None), (1, rng.permutation(n).astype(np.int32), None), (2, rng.integers(-50, 50, n, dtype=np.int32), None)])
    sd = abi.ScanDesc(preds=[abi.make_pred(0, abi.PRED_GE, 70_000)])
    agg_desc = abi.make_agg_desc([1], [abi.TYPE_INT], fns=[(abi.AGG_SUM, abi.TYPE_INT, 10, [("col", 2)]), (abi.AGG_COUNT_STAR, 0, 11, None)])
    frag = gpu.Fragment(ctx, sd, [], agg_desc, mode=2)
    try:
        frag.push(fact)
        assert frag.rows_passed == n - 70_000
        got = gpu_rows(frag.agg.result())
        ores, opassed = oracle.fragment_run(sd, [], agg_desc, fact, num_threads=4)
        assert opassed == n - 70_000 and len(got) == n - 70_000
        assert_rows_equal(got, oracle_rows(ores))
    finally:
        frag.close()


@pytest.mark.parametrize("shape", ["hash_group", "no_group", "dense_group", "two_expanding", "grow"])
def test_fragment_one_to_many_join(gpu, ctx, oracle, shape):
    # duplicate build keys (reference: the one-to-many walk of JoinHashMap::_probe_from_ht, join_hash_map.hpp:718-795):
    # a probe row counts once per build row of its key; payload columns come from every chain entry in turn
    rng = np.random.default_rng(91)
    n = 300_000 if shape != "grow" else 1_200_000
    nb = 40_000
    fact = Chunk([(0, rng.integers(0, 20_000, n, dtype=np.int32), rand_nulls(rng, n, 0.01)), (1, rng.integers(0, 3000, n, dtype=np.int32), None),
                  (2, rng.integers(-1000, 1000, n, dtype=np.int64), None), (3, rng.integers(0, 50, n, dtype=np.int32), None)])
    # build 1: keys 0..9999 with 0-7 rows each (skewed), payload differs per row; a few NULL keys that never match
    k1 = np.repeat(np.arange(10_000, dtype=np.int32), rng.integers(0, 8, 10_000))
    rng.shuffle(k1)
    dim1 = Chunk([(10, k1, rand_nulls(rng, len(k1), 0.01)), (11, rng.integers(0, 200, len(k1), dtype=np.int32), None),
                  (12, rng.integers(-5, 5, len(k1), dtype=np.int64), rand_nulls(rng, len(k1), 0.05))])
    k2 = rng.integers(0, 2500, nb // 8, dtype=np.int32)            # build 2: ~2 rows per key, sparse domain -> may be a hash table
    dim2 = Chunk([(20, k2 * 7, None), (21, rng.integers(0, 30, len(k2), dtype=np.int32), None)])
    d1 = abi.make_join_desc(abi.JOIN_INNER, [10], [0], [abi.TYPE_INT], build_out=[11, 12])
    d2 = abi.make_join_desc(abi.JOIN_INNER, [20], [1], [abi.TYPE_INT], build_out=[21])
    sd = abi.ScanDesc(preds=[abi.make_pred(3, abi.PRED_LT, 40)])
    fns = [(abi.AGG_SUM, abi.TYPE_BIGINT, 30, [("col", 2), ("col", 12), "*"]), (abi.AGG_COUNT_STAR, 0, 31, None), (abi.AGG_COUNT, abi.TYPE_BIGINT, 32, [("col", 12)]),
           (abi.AGG_MIN, abi.TYPE_INT, 33, [("col", 11)])]
    if shape == "no_group":
        agg_desc = abi.make_agg_desc(fns=fns)
    elif shape == "dense_group":
        agg_desc = abi.make_agg_desc([11], [abi.TYPE_INT], fns=fns, ranges=[(0, 199)])
    elif shape == "grow":
        # the 64 K-row sample sees no survivor (conjunct on the row number), the rest survive and nearly every
        # (fact value, payload) pair is a new group -> refused rows are replayed whole after the table has grown
        fact = Chunk([(0, fact.columns()[0][1], None), (1, fact.columns()[1][1], None), (2, np.arange(n, dtype=np.int64), None),
                      (3, np.where(np.arange(n) < 70_000, 45, 1).astype(np.int32), None)])
        agg_desc = abi.make_agg_desc([2, 11], [abi.TYPE_BIGINT, abi.TYPE_INT], fns=fns[1:])
    else:
        agg_desc = abi.make_agg_desc([11, 3], [abi.TYPE_INT, abi.TYPE_INT], fns=fns)
    two = shape == "two_expanding"
    gj1, gj2 = gpu.Join(ctx, d1), gpu.Join(ctx, d2)
    oj1, oj2 = oracle.Join(d1), oracle.Join(d2)
    frag = None
    try:
        for j, ch in ((gj1, dim1), (gj2, dim2), (oj1, dim1), (oj2, dim2)):
            j.append_build(ch)
        gj1.build_finish()
        gj2.build_finish()
        oj1.build()
        oj2.build()
        assert gj1.info().has_duplicates and gj2.info().has_duplicates
        gjoins = [(gj1, 0, [11, 12])] + ([(gj2, 1, [21])] if two else [])
        ojoins = [(oj1, 0, [11, 12])] + ([(oj2, 1, [21])] if two else [])
        if two:
            agg_desc = abi.make_agg_desc([11, 21], [abi.TYPE_INT, abi.TYPE_INT], fns=fns)
        frag = gpu.Fragment(ctx, sd, gjoins, agg_desc)
        half = n // 2 + 13
        for lo, hi in ((0, half), (half, n)):
            frag.push(Chunk([(sl, d[lo:hi], None if nl is None else nl[lo:hi]) for sl, d, nl in fact.columns()]))
        got = gpu_rows(frag.agg.result())
        ores, opassed = oracle.fragment_run(sd, ojoins, agg_desc, fact, num_threads=3)
        assert opassed > (n // 2 if not two else 50_000)   # the expansion really happened (about 3.5 build rows per matching key)
        assert frag.rows_passed == opassed
        assert_rows_equal(got, oracle_rows(ores))
    finally:
        if frag:
            frag.close()
        gj1.close()
        gj2.close()


def test_q95_plan_parity(gpu, ctx, oracle):
    # BASELINE config 4's shape on one GPU: the same plan (starrocks_b200.tpcds.q95_local_plan) through the CUDA operators and
    # through the oracle -- one-to-many self join with an other-conjunct filter, GROUP BY dedup, four IN / dimension semi
    # joins, COUNT(DISTINCT) + SUMs -- must give the same three numbers and the same intermediate row counts
    from starrocks_b200 import tpcds
    g = tpcds.Q95Gen(1.0)
    ws, wr = g.web_sales_of_orders(0, g.n_orders), g.web_returns_of_orders(0, g.n_orders)
    def dims(mem):
        return {"date": Chunk([(tpcds.D_DATE_SK, g.date_keys(), None)]), "addr": Chunk([(tpcds.CA_ADDRESS_SK, g.address_keys(), None)]),
                "site": Chunk([(tpcds.WEB_SITE_SK, g.site_keys(), None)])}
    ws_chunk, wr_chunk = tpcds.table_chunk(ws, tpcds.WS_COLS), Chunk([(tpcds.WS_ORDER, wr["wr_order_number"], None)])
    ores, ost = tpcds.q95_local_plan(tpcds.OracleEngine(oracle), ws_chunk, wr_chunk, dims(0), morsel_rows=200_000)
    eng = tpcds.GpuEngine(gpu, ctx)
    eng.mem = abi.MEM_HOST               # host chunks in: the operators stage them
    try:
        gres, gst = tpcds.q95_local_plan(eng, ws_chunk, wr_chunk, dims(0), morsel_rows=200_000, expected_orders=g.n_orders)
    finally:
        eng.close()
    assert gres == ores == g.expected(0, g.n_orders)
    assert gst == ost and gres[0] > 50


# ---------------------------------------------------------------------------------------------
# exchange partitioning (K18)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("hash_fn,reduce_op", [(abi.HASH_FNV, abi.REDUCE_MULHI), (abi.HASH_CRC32, abi.REDUCE_MODULO), (abi.HASH_XXH3, abi.REDUCE_MULHI)])
@pytest.mark.parametrize("nch", [1, 8, 37])
@pytest.mark.parametrize("n", [0, 1, 1023, 50_001])
def test_xchg_partition_parity(gpu, ctx, oracle, hash_fn, reduce_op, nch, n):
    rng = np.random.default_rng(21)
    chunk = Chunk([(0, rng.integers(-10**6, 10**6, n, dtype=np.int32), rand_nulls(rng, n, 0.05)),
                   (1, rng.integers(-10**15, 10**15, n, dtype=np.int64), None), (2, rng.normal(size=n), None)])
    d = abi.make_part_desc([0, 1], nch, hash_fn=hash_fn, reduce_op=reduce_op)
    x = gpu.Xchg(ctx, d)
    try:
        ohv, och, ori, ost = oracle.hash_partition(d, chunk)
        hv, ch = x.hash(chunk)
        assert np.array_equal(hv, ohv) and np.array_equal(ch, och)
        out, offs = x.partition(chunk)
        assert offs.tolist() == ost.tolist()
        got = gpu.chunk_out_to_host(ctx, out)
        for k, (slot, typ, data, nulls) in enumerate(got):
            src = chunk._keep[k][0]
            assert np.array_equal(data.view(np.uint8), src[ori].view(np.uint8))      # stable: same row order as the reference's counting sort
        if n > 0:
            assert np.array_equal(got[0][3], chunk._keep[0][1][ori])
    finally:
        x.close()


def test_xchg_xxh3_all_widths(gpu, ctx, oracle):
    # exchange_hash_function_version = 1 (XXH3) over every fixed width: 1 / 2 bytes take XXH3_len_1to3_64b, 4 / 8 bytes
    # XXH3_len_4to8_64b, 16 bytes (decimal128) XXH3_len_9to16_64b -- the oracle is pinned by the reference's goldens and by
    # the reference's own xxhash.h (tests/test_oracle_golden.py)
    rng = np.random.default_rng(23)
    n = 20_011
    dec = rng.integers(0, 256, (n, 16), dtype=np.uint8)
    chunk = Chunk([(0, rng.integers(-128, 128, n, dtype=np.int8), None), (1, rng.integers(-30000, 30000, n, dtype=np.int16), rand_nulls(rng, n, 0.1)),
                   (2, rng.integers(-10**9, 10**9, n, dtype=np.int32), None), (3, rng.integers(-10**17, 10**17, n, dtype=np.int64), None),
                   (4, dec.reshape(-1).view(np.dtype((np.void, 16))), None, abi.TYPE_DECIMAL128)])
    for slots in ([0], [1], [2], [3], [4], [0, 1, 2, 3], [4, 1]):
        d = abi.make_part_desc(slots, 8, hash_fn=abi.HASH_XXH3)
        x = gpu.Xchg(ctx, d)
        try:
            ohv, och, _, _ = oracle.hash_partition(d, chunk)
            hv, ch = x.hash(chunk)
            assert np.array_equal(hv, ohv) and np.array_equal(ch, och), slots
        finally:
            x.close()


@pytest.mark.parametrize("n", [0, 1, 4096, 70_001])
@pytest.mark.parametrize("device_columns", [False, True])
def test_chunk_wire_format_parity(gpu, ctx, oracle, n, device_columns):
    # ChunkPB.data: the device serialiser must produce the oracle's bytes (slices of a partitioned chunk are
    # what the exchange ships per channel), and deserialising them gives the columns back
    rng = np.random.default_rng(31)
    dec = rng.integers(0, 256, (n, 16), dtype=np.uint8)
    cols = [(3, rng.integers(-10**9, 10**9, n, dtype=np.int32), rand_nulls(rng, n, 0.2)), (4, rng.integers(-10**17, 10**17, n, dtype=np.int64), None),
            (5, rng.integers(-100, 100, n, dtype=np.int8), rand_nulls(rng, n, 0.5)), (6, rng.normal(size=n), None),
            (7, dec.reshape(-1).view(np.dtype((np.void, 16))), None, abi.TYPE_DECIMAL128)]
    host = Chunk(cols)
    src = host
    keep = []
    if device_columns and n > 0:          # run the columns through the partition step: its output lives on the device
        x = gpu.Xchg(ctx, abi.make_part_desc([3], 1))
        out, _ = x.partition(host)
        src = gpu.chunk_out_as_view(out)
        keep.append(x)
    try:
        for lo, hi in ((0, n), (n // 3, n - n // 4)):
            payload, meta = gpu.chunk_serialize(ctx, src, lo, hi)
            assert payload.tobytes() == oracle.chunk_serialize(host, lo, hi).tobytes()
            assert meta.serialized_size == payload.nbytes and meta.num_rows == hi - lo and list(meta.is_nulls[:5]) == [1, 0, 1, 0, 0]
            sd = gpu.Serde(ctx)
            try:
                back = gpu.chunk_out_to_host(ctx, sd.deserialize(payload, meta))
                for k, c in enumerate(cols):
                    assert np.array_equal(back[k][2].view(np.uint8), c[1][lo:hi].view(np.uint8)) and back[k][0] == c[0]
                    assert (back[k][3] is None) == (c[2] is None) and (c[2] is None or np.array_equal(back[k][3], c[2][lo:hi]))
                if hi - lo > 0:            # a truncated payload must be refused, not read out of bounds
                    with pytest.raises(gpu.GpuError):
                        sd.deserialize(payload[:-3], meta)
            finally:
                sd.close()
    finally:
        for x in keep:
            x.close()


def test_large_batches_take_the_two_level_scan(gpu, ctx, oracle):
    # 5 M-row probe (19.5 K block counts) and a 37-channel partition of 3 M rows (> 16 K tile x channel counts): the
    # exclusive scans behind the ordered outputs switch from the single-block walk to the two-level form
    rng = np.random.default_rng(41)
    n = 5_000_000
    bkeys = rng.integers(0, 50_000, 60_000, dtype=np.int32)                # duplicates on the build side
    build = Chunk([(10, bkeys, None), (11, np.arange(60_000, dtype=np.int32), None)])
    probe = Chunk([(0, rng.integers(0, 100_000, n, dtype=np.int32), None), (1, np.arange(n, dtype=np.int32), None)])
    d = abi.make_join_desc(abi.JOIN_INNER, [10], [0], [abi.TYPE_INT], build_out=[11], probe_out=[1])
    gj, oj = gpu.Join(ctx, d), oracle.Join(d)
    x = gpu.Xchg(ctx, abi.make_part_desc([0], 37))
    try:
        gj.append_build(build)
        gj.build_finish()
        oj.append_build(build)
        oj.build()
        out = gj.probe(probe)
        pi, bi = oj.probe_all(probe)
        assert out.num_rows == len(pi)
        gpi, gbi = gj.probe_indexes(out.num_rows)
        assert np.array_equal(gpi, pi) and np.array_equal(gbi, bi)            # same pairs, same order (chains: descending build index)
        part = Chunk([(0, rng.integers(-10**6, 10**6, 3_000_000, dtype=np.int32), None)])
        ohv, och, ori, ost = oracle.hash_partition(abi.make_part_desc([0], 37), part)
        pout, offs = x.partition(part)
        assert offs.tolist() == ost.tolist()
        got = gpu.chunk_out_to_host(ctx, pout)
        assert np.array_equal(got[0][2], part._keep[0][0][ori])
    finally:
        gj.close()
        x.close()


# ---------------------------------------------------------------------------------------------
# loud failures (no silent fallback)
# ---------------------------------------------------------------------------------------------
def test_errors_are_loud(gpu, ctx):
    with pytest.raises(gpu.GpuError):
        gpu.Join(ctx, abi.make_join_desc(abi.JOIN_INNER, [1], [0], [abi.TYPE_DOUBLE]))
    j = gpu.Join(ctx, abi.make_join_desc(abi.JOIN_INNER, [1], [0], [abi.TYPE_INT]))
    try:
        with pytest.raises(gpu.GpuError) as ei:
            j.probe(Chunk([(0, np.arange(4, dtype=np.int32), None)]))
        assert ei.value.code == abi.SR_ERR_STATE
    finally:
        j.close()
    a = gpu.Agg(ctx, abi.make_agg_desc([0], [abi.TYPE_INT], fns=[(abi.AGG_COUNT_STAR, 0, 1, None)], ranges=[(0, 3)]))
    try:
        a.push(Chunk([(0, np.array([0, 1, 9], dtype=np.int32), None)]))   # 9 is outside the declared range
        with pytest.raises(gpu.GpuError):
            a.result()
    finally:
        a.close()
    s = gpu.Scan(ctx, abi.ScanDesc(preds=[abi.make_pred(5, abi.PRED_EQ, 1)], out_slots=[0]))
    try:
        with pytest.raises(gpu.GpuError):
            s.filter(Chunk([(0, np.arange(4, dtype=np.int32), None)]))   # slot 5 does not exist
    finally:
        s.close()


def test_agg_compressed_key_sql_goldens_gpu(gpu, ctx):
    # test/sql/test_agg/R/test_agg_compressed_key through the dense (range-declared) CUDA aggregate, in two pushes
    from tests.test_oracle_golden import _compressed_key_check

    def run(d, chunk):
        a = gpu.Agg(ctx, d)
        try:
            n = chunk.num_rows
            cols = chunk.columns()
            for lo, hi in ((0, n // 3), (n // 3, n)):
                a.push(Chunk([(s, arr[lo:hi].copy(), None if nl is None else nl[lo:hi].copy(), t) for (s, arr, nl), t in zip(cols, chunk.types)]))
            return gpu_rows(a.result())
        finally:
            a.close()
    _compressed_key_check(run)


# ---------------------------------------------------------------------------------------------
# segment data pages decoded on the device
# ---------------------------------------------------------------------------------------------
def _page_columns(dt, rng):
    info = np.iinfo(dt)
    yield "ssb_keys_22bit", rng.integers(1, 3_000_000, 100_003).astype(dt)
    yield "dates_ascending", np.sort(rng.integers(19920101, 19981231, 70_000)).astype(dt)          # format 1 frames
    yield "constant", np.full(1000, 42, dtype=dt)                                                     # bit width 0
    yield "negative_mixed", rng.integers(-10**6, 10**6, 12_345).astype(dt)
    yield "wide", rng.integers(info.min // 2 + 1, info.max // 2, 5000, dtype=dt)                      # widths 31 / 63
    yield "one_value", np.array([7], dtype=dt)
    yield "empty", np.zeros(0, dtype=dt)
    yield "original_values_single_frame", np.array([info.min, info.max, -1, 0, 1] * 20, dtype=dt)     # format 2 (see test_oracle_golden)


@pytest.mark.parametrize("dt,typ", [(np.int32, abi.TYPE_INT), (np.int64, abi.TYPE_BIGINT)])
@pytest.mark.parametrize("page_rows", [128, 1000, 65_536])
@pytest.mark.parametrize("mem", ["host", "device"])
def test_for_pages_decode_parity(gpu, ctx, oracle, dt, typ, page_rows, mem):
    # pages written by the (reference-pinned) oracle encoder, decoded by k_for_frames / k_for_decode: bit-exact with
    # ForDecoder as restated in the oracle; pages of one column are concatenated in order
    import torch
    rng = np.random.default_rng(5)
    dec = gpu.PageDecoder(ctx)
    try:
        for name, v in _page_columns(dt, rng):
            pages = [oracle.for_encode(v[lo:lo + page_rows]) for lo in range(0, max(len(v), 1), page_rows)]
            for pg, lo in zip(pages, range(0, max(len(v), 1), page_rows)):
                assert (oracle.for_decode(pg, dt) == v[lo:lo + page_rows]).all(), name
            out = torch.full((len(v) + 8,), -77, dtype=torch.int32 if dt == np.int32 else torch.int64, device="cuda")
            if mem == "device":
                blobs = [torch.from_numpy(np.concatenate([np.zeros(k % 4, dtype=np.uint8), pg])).cuda() for k, pg in enumerate(pages)]   # odd alignments
                views = [(b.data_ptr() + k % 4, len(pg)) for k, (b, pg) in enumerate(zip(blobs, pages))]
                rows = dec.decode(abi.PAGE_FOR, typ, views, out.data_ptr(), len(v), mem=abi.MEM_DEVICE)
            else:
                rows = dec.decode(abi.PAGE_FOR, typ, pages, out.data_ptr(), len(v))
            ctx.sync()
            assert rows == len(v), name
            got = out.cpu().numpy()
            assert (got[:len(v)] == v).all(), name
            assert (got[len(v):] == -77).all(), name            # nothing written past the column
    finally:
        dec.close()


def test_plain_pages_and_malformed_pages(gpu, ctx, oracle):
    import torch
    rng = np.random.default_rng(6)
    dec = gpu.PageDecoder(ctx)
    try:
        v = rng.integers(-10**9, 10**9, 10_001).astype(np.int64)
        pages = [oracle.plain_encode(v[lo:lo + 999]) for lo in range(0, len(v), 999)]
        out = torch.zeros(len(v), dtype=torch.int64, device="cuda")
        assert dec.decode(abi.PAGE_PLAIN, abi.TYPE_BIGINT, pages, out.data_ptr(), len(v)) == len(v)
        ctx.sync()
        assert (out.cpu().numpy() == v).all()
        good = oracle.for_encode(v[:300])
        for bad in (good[:3], np.concatenate([good[:-4], np.array([255, 255, 0, 0], dtype=np.uint8)]),     # too short; value count beyond the page
                    np.concatenate([good[:-7], np.array([0, 200], dtype=np.uint8), good[-5:]])):            # bit width 200
            with pytest.raises(gpu.GpuError) as ei:
                dec.decode(abi.PAGE_FOR, abi.TYPE_BIGINT, [bad], out.data_ptr(), len(v))
            assert ei.value.code == abi.SR_ERR_INVALID_ARGUMENT
        with pytest.raises(gpu.GpuError) as ei:                                                               # output too small
            dec.decode(abi.PAGE_FOR, abi.TYPE_BIGINT, [good], out.data_ptr(), 100)
        assert ei.value.code == abi.SR_ERR_INVALID_ARGUMENT
        with pytest.raises(gpu.GpuError) as ei:
            dec.decode(abi.PAGE_FOR, abi.TYPE_DOUBLE, [good], out.data_ptr(), len(v))
        assert ei.value.code == abi.SR_ERR_NOT_SUPPORTED
    finally:
        dec.close()


def test_for_pages_feed_the_scan(gpu, ctx, oracle):
    # the decoded column is an ordinary device column: scan + filter over it equals the oracle's scan over the raw values
    import torch
    rng = np.random.default_rng(8)
    n = 200_000
    key = rng.integers(0, 1000, n).astype(np.int32)
    val = rng.integers(0, 10**6, n).astype(np.int64)
    dec = gpu.PageDecoder(ctx)
    try:
        dk, dv = torch.empty(n, dtype=torch.int32, device="cuda"), torch.empty(n, dtype=torch.int64, device="cuda")
        dec.decode(abi.PAGE_FOR, abi.TYPE_INT, [oracle.for_encode(key[lo:lo + 50_000]) for lo in range(0, n, 50_000)], dk.data_ptr(), n)
        dec.decode(abi.PAGE_FOR, abi.TYPE_BIGINT, [oracle.for_encode(val[lo:lo + 30_000]) for lo in range(0, n, 30_000)], dv.data_ptr(), n)
        sd = abi.ScanDesc(preds=[abi.make_pred(0, abi.PRED_LT, 100)], out_slots=[0, 1])
        sc = gpu.Scan(ctx, sd)
        got = gpu.chunk_out_to_host(ctx, sc.filter(Chunk([(0, dk, None), (1, dv, None)], num_rows=n, mem=abi.MEM_DEVICE)))
        rows, ores = oracle.scan_filter(sd, Chunk([(0, key, None), (1, val, None)]))
        gcols = {s: d for s, _, d, _ in got}
        assert len(gcols[0]) == rows and (gcols[0] == ores[0][0]).all() and (gcols[1] == ores[1][0]).all()
        sc.close()
    finally:
        dec.close()
