"""GPU parity tests for runtime filters: the CUDA min/max + split-block bloom filter against the CPU
restatement of MinMaxRuntimeFilter / SimdBlockFilter (oracle/sr_oracle.cpp, pinned by the reference's own
runtime_filter_core_test.cpp vectors in tests/test_oracle_golden.py).  Everything is integer/bit work: the bloom
DIRECTORY BYTES, the info block and every selection byte must be identical.
"""
import numpy as np
import pytest

from starrocks_b200 import abi
from starrocks_b200.abi import Chunk
from tests.helpers import rand_nulls

pytestmark = pytest.mark.gpu

KEY_CASES = [
    (np.int32, abi.TYPE_INT, -10**6, 10**6),
    (np.int64, abi.TYPE_BIGINT, -2**62, 2**62),
    (np.int16, abi.TYPE_SMALLINT, -30000, 30000),
    (np.int8, abi.TYPE_TINYINT, -128, 128),
    (np.int32, abi.TYPE_DATE, 700000, 740000),
]


def _same_info(a, b):
    for f in ("min_value", "max_value", "num_inserted", "has_null", "log_num_buckets", "key_type", "num_in_values"):
        assert getattr(a, f) == getattr(b, f), f


@pytest.mark.parametrize("dtype,typ,lo,hi", KEY_CASES)
@pytest.mark.parametrize("n", [0, 1, 33, 1000, 100_003])
@pytest.mark.parametrize("with_bloom", [True, False])
def test_rf_insert_builds_the_reference_directory(gpu, ctx, oracle, dtype, typ, lo, hi, n, with_bloom):
    rng = np.random.default_rng(5 + n)
    keys = rng.integers(lo, hi, n).astype(dtype)
    nulls = rand_nulls(rng, n, 0.1) if n > 1 else None
    chunk = Chunk([(7, keys, nulls, typ)])
    g = gpu.RuntimeFilter(ctx, typ, n, with_bloom)
    o = oracle.RuntimeFilter(typ, n, with_bloom)
    try:
        # two inserts (chunks arrive one at a time on the build side); the second one records NULLs
        g.insert(chunk, 7, insert_nulls=False)
        o.insert(chunk, 7, insert_nulls=False)
        _same_info(g.info(), o.info())
        g.insert(chunk, 7, insert_nulls=True)
        o.insert(chunk, 7, insert_nulls=True)
        _same_info(g.info(), o.info())
        assert np.array_equal(g.directory(), o.directory())
        # evaluate on a probe column that overlaps the key domain, with NULLs
        m = 50_021
        pk = np.concatenate([keys[rng.integers(0, max(n, 1), m // 2)] if n else np.zeros(m // 2, dtype=dtype),
                             rng.integers(lo, hi, m - m // 2).astype(dtype)])
        probe = Chunk([(3, pk, rand_nulls(rng, m, 0.05), typ)])
        sel = g.evaluate(probe, 3)
        assert np.array_equal(sel, o.evaluate(probe, 3))
        prior = rng.integers(0, 2, m).astype(np.uint8)
        assert np.array_equal(g.evaluate(probe, 3, prior), o.evaluate(probe, 3, prior))
        if n:  # no false negatives among the inserted (non-NULL) keys
            hit = Chunk([(3, keys, None, typ)])
            inserted = np.ones(n, dtype=bool) if nulls is None else nulls == 0
            assert g.evaluate(hit, 3)[inserted].all()
    finally:
        g.close()


def test_rf_merge_is_the_union(gpu, ctx, oracle):
    # SimdBlockFilter::merge (runtime_filter.h:126-140) + min/max/has_null union (RuntimeBloomFilter::merge)
    rng = np.random.default_rng(9)
    a = rng.integers(0, 10**6, 40_000, dtype=np.int32)
    b = rng.integers(-10**6, 0, 30_000, dtype=np.int32)
    bn = rand_nulls(rng, len(b), 0.01)
    ca, cb = Chunk([(0, a, None)]), Chunk([(0, b, bn)])
    ga, gb = gpu.RuntimeFilter(ctx, abi.TYPE_INT, 70_000), gpu.RuntimeFilter(ctx, abi.TYPE_INT, 70_000)
    oa, ob = oracle.RuntimeFilter(abi.TYPE_INT, 70_000), oracle.RuntimeFilter(abi.TYPE_INT, 70_000)
    try:
        ga.insert(ca, 0)
        oa.insert(ca, 0)
        gb.insert(cb, 0, insert_nulls=True)
        ob.insert(cb, 0, insert_nulls=True)
        # GPU filter <- directory built on the CPU (a filter shipped from a CPU BE) and vice versa
        ga.merge(ob.directory(), ob.info())
        oa.merge(ob)
        _same_info(ga.info(), oa.info())
        assert np.array_equal(ga.directory(), oa.directory())
        both = Chunk([(0, np.concatenate([a, b[bn == 0]]), None)])
        assert ga.evaluate(both, 0).all()
        # sizes must agree (SimdBlockFilter::merge DCHECK)
        small = gpu.RuntimeFilter(ctx, abi.TYPE_INT, 100)
        with pytest.raises(gpu.GpuError):
            small.merge(ob.directory(), ob.info())
        small.close()
    finally:
        ga.close()
        gb.close()


@pytest.mark.parametrize("kind", ["range_i32", "sparse_i64_nulls", "two_keys"])
def test_join_build_publishes_its_runtime_filter(gpu, ctx, oracle, kind):
    # HashJoinBuildOperator::set_finishing -> create_runtime_filters (hash_join_build_operator.cpp:100-215): the filter
    # of key k holds exactly the build rows' key-k values, sized by the table's row count
    rng = np.random.default_rng(21)
    nb = 30_000
    if kind == "range_i32":
        cols = [(0, rng.integers(1000, 90_000, nb, dtype=np.int32), None)]
        ktypes, kslots = [abi.TYPE_INT], [0]
    elif kind == "sparse_i64_nulls":
        cols = [(0, rng.integers(-2**60, 2**60, nb, dtype=np.int64), rand_nulls(rng, nb, 0.03))]
        ktypes, kslots = [abi.TYPE_BIGINT], [0]
    else:
        cols = [(0, rng.integers(0, 5000, nb, dtype=np.int32), None), (1, rng.integers(-300, 300, nb, dtype=np.int32), None)]
        ktypes, kslots = [abi.TYPE_INT, abi.TYPE_INT], [0, 1]
    build = Chunk(cols + [(5, rng.integers(0, 100, nb, dtype=np.int32), None)])
    d = abi.make_join_desc(abi.JOIN_INNER, kslots, [10 + s for s in kslots], ktypes, build_out=[5], probe_out=[10])
    j = gpu.Join(ctx, d)
    try:
        half = nb // 2
        j.append_build(Chunk([(s, a[:half], None if nl is None else nl[:half]) for s, a, nl in build.columns()]))
        j.append_build(Chunk([(s, a[half:], None if nl is None else nl[half:]) for s, a, nl in build.columns()]))
        j.build_finish()
        for k, slot in enumerate(kslots):
            for insert_nulls in (False, True):
                g = gpu.RuntimeFilter.from_join(j, k, True, insert_nulls)
                o = oracle.RuntimeFilter(ktypes[k], nb, True)
                o.insert(build, slot, insert_nulls)
                _same_info(g.info(), o.info())
                assert np.array_equal(g.directory(), o.directory())
                g.close()
    finally:
        j.close()


@pytest.mark.parametrize("fast", [True, False])
@pytest.mark.parametrize("n", [0, 257, 100_003, 1_000_003])
def test_scan_applies_attached_runtime_filters(gpu, ctx, oracle, fast, n):
    # ScanOperator side (scan_operator.h:212-225, RuntimeFilterProbeCollector::evaluate): conjuncts AND every filter
    rng = np.random.default_rng(31 + n)
    dim1 = rng.choice(200_000, 20_000, replace=False).astype(np.int32)
    dim2 = rng.integers(-2**40, 2**40, 5000, dtype=np.int64)
    k1 = rng.integers(0, 200_000, n, dtype=np.int32)
    k2 = np.where(rng.random(n) < 0.5, dim2[rng.integers(0, len(dim2), n)], rng.integers(-2**40, 2**40, n, dtype=np.int64))
    chunk = Chunk([
        (0, rng.integers(0, 100, n, dtype=np.int32), None),
        (1, k1, None),
        (2, k2, None if fast else rand_nulls(rng, n, 0.05)),
        (3, rng.integers(0, 10**6, n, dtype=np.int64), None),
    ])
    preds = [abi.make_pred(0, abi.PRED_BETWEEN, 10, 80)] if fast else [abi.make_pred(0, abi.PRED_NE, 7), abi.make_pred(3, abi.PRED_GE, 1000)]
    sd = abi.ScanDesc(preds=preds, filter_exprs=[], out_slots=[0, 1, 2, 3])
    g1, g2 = gpu.RuntimeFilter(ctx, abi.TYPE_INT, len(dim1)), gpu.RuntimeFilter(ctx, abi.TYPE_BIGINT, len(dim2))
    o1, o2 = oracle.RuntimeFilter(abi.TYPE_INT, len(dim1)), oracle.RuntimeFilter(abi.TYPE_BIGINT, len(dim2))
    c1 = Chunk([(0, dim1, None)])
    c2 = Chunk([(0, dim2, rand_nulls(rng, len(dim2), 0.01))])
    for f, c in ((g1, c1), (o1, c1), (g2, c2), (o2, c2)):
        f.insert(c, 0, insert_nulls=True)
    scan = gpu.Scan(ctx, sd)
    try:
        scan.add_runtime_filter(g1, 1)
        scan.add_runtime_filter(g2, 2)
        expect = oracle.scan_evaluate(sd, chunk)
        expect = o1.evaluate(chunk, 1, expect)
        expect = o2.evaluate(chunk, 2, expect)
        assert np.array_equal(scan.evaluate(chunk), expect)
        out = gpu.chunk_out_to_host(ctx, scan.filter(chunk))
        keep = expect.astype(bool)
        for (slot, typ, data, nulls), (s0, arr, nl) in zip(out, chunk.columns()):
            assert slot == s0 and np.array_equal(data, arr[keep])
            if nl is not None:
                assert np.array_equal(nulls, nl[keep])
        if n:  # the filters do drop rows the conjuncts keep
            assert expect.sum() < oracle.scan_evaluate(sd, chunk).sum()
    finally:
        scan.close()
        g1.close()
        g2.close()


@pytest.mark.parametrize("dtype,typ,lo,hi", KEY_CASES)
@pytest.mark.parametrize("n", [1, 40, 1024, 1025])
def test_runtime_in_filter(gpu, ctx, oracle, dtype, typ, lo, hi, n):
    # runtime IN filter (HashJoiner::_create_runtime_in_filters, hash_joiner.cpp:563-609; row limit 1024): a small build side
    # also publishes its distinct keys, membership becomes EXACT (no bloom false positive); beyond the limit the IN part is gone
    rng = np.random.default_rng(61 + n)
    keys = rng.integers(lo, hi, n).astype(dtype)
    half = n // 2
    g, o = gpu.RuntimeFilter(ctx, typ, n), oracle.RuntimeFilter(typ, n)
    first_nulls = rand_nulls(rng, half, 0.1) if half > 3 else None
    try:
        for f in (g, o):
            f.insert(Chunk([(0, keys[:half], first_nulls, typ)]), 0)
            f.insert(Chunk([(0, keys[half:], None, typ)]), 0)
        _same_info(g.info(), o.info())
        vals = g.in_values()
        if n <= abi.RF_IN_FILTER_ROW_LIMIT:
            assert vals is not None and len(vals) == g.info().num_in_values and (np.diff(vals) > 0).all()
        else:
            assert vals is None and g.info().num_in_values == -1
        probe = np.concatenate([keys, rng.integers(lo, hi, 50_000).astype(dtype)])
        pc = Chunk([(3, probe, rand_nulls(rng, len(probe), 0.02), typ)])
        sel = g.evaluate(pc, 3)
        assert np.array_equal(sel, o.evaluate(pc, 3))
        if vals is not None:      # exact: a row passes iff its value is one of the inserted keys
            nn = pc.columns()[0][2] == 0
            assert np.array_equal(sel[nn].astype(bool), np.isin(probe[nn].astype(np.int64), vals))
    finally:
        g.close()


def test_runtime_in_filter_merge(gpu, ctx, oracle):
    a, b, c = (gpu.RuntimeFilter(ctx, abi.TYPE_INT, 100) for _ in range(3))
    big = gpu.RuntimeFilter(ctx, abi.TYPE_INT, 5000)
    try:
        a.insert(Chunk([(0, np.arange(0, 600, dtype=np.int32), None)]), 0)
        b.insert(Chunk([(0, np.arange(500, 900, dtype=np.int32), None)]), 0)
        a.merge_in_values(b.in_values())
        assert np.array_equal(a.in_values(), np.arange(0, 900))
        c.insert(Chunk([(0, np.arange(2000, 2300, dtype=np.int32), None)]), 0)
        a.merge_in_values(c.in_values())          # 1200 distinct keys: over the limit -> no IN part
        assert a.in_values() is None
        big.insert(Chunk([(0, np.arange(5000, dtype=np.int32), None)]), 0)
        assert big.in_values() is None
        b.merge_in_values(big.in_values())        # a partial filter without an IN part -> the total has none
        assert b.in_values() is None
    finally:
        for f in (a, b, c, big):
            f.close()


def test_scan_stops_evaluating_an_unselective_filter(gpu, ctx, oracle):
    # RuntimeFilterProbeCollector::update_selectivity (runtime_filter_probe.cpp:408-480): a filter that lets more than half of
    # its rows through is switched off for the next 31 batches; the result is the same set of rows a plan without that
    # filter plus the join would keep, so parity here is against the oracle WITH the filters for the batches that used
    # them and the bookkeeping is checked through sr_scan_get_rf_stats
    rng = np.random.default_rng(71)
    n = 200_000
    useful = rng.choice(100_000, 5_000, replace=False).astype(np.int32)        # ~5 % of the probe keys
    useless = np.arange(0, 95_000, dtype=np.int32)                              # ~95 % of the probe keys
    chunk = Chunk([(0, rng.integers(0, 100_000, n, dtype=np.int32), None), (1, rng.integers(0, 100_000, n, dtype=np.int32), None)])
    sd = abi.ScanDesc(out_slots=[0, 1])
    g1, g2 = gpu.RuntimeFilter(ctx, abi.TYPE_INT, len(useful)), gpu.RuntimeFilter(ctx, abi.TYPE_INT, len(useless))
    o1, o2 = oracle.RuntimeFilter(abi.TYPE_INT, len(useful)), oracle.RuntimeFilter(abi.TYPE_INT, len(useless))
    for f, keys in ((g1, useful), (o1, useful), (g2, useless), (o2, useless)):
        f.insert(Chunk([(0, keys, None)]), 0)
    scan = gpu.Scan(ctx, sd)
    try:
        scan.add_runtime_filter(g1, 0)
        scan.add_runtime_filter(g2, 1)
        both = o2.evaluate(chunk, 1, o1.evaluate(chunk, 0)).astype(bool)
        only1 = o1.evaluate(chunk, 0).astype(bool)
        rows = [gpu.chunk_out_to_host(ctx, scan.filter(chunk))[0][2] for _ in range(4)]
        assert np.array_equal(rows[0], chunk.columns()[0][1][both])               # batch 0 samples both filters
        for r in rows[1:]:
            assert np.array_equal(r, chunk.columns()[0][1][only1])                # the unselective one is off afterwards
        s1, s2 = scan.rf_stats(0), scan.rf_stats(1)
        assert s1.batches_skipped == 0 and s1.rows_tested == 4 * n and s1.last_selectivity < 0.1
        assert s2.batches_skipped == 3 and s2.last_selectivity > 0.9 and s2.rows_tested == s1.rows_passed // 4
        scan.set_rf_adaptive(False)                                               # every filter on every batch again
        assert np.array_equal(gpu.chunk_out_to_host(ctx, scan.filter(chunk))[0][2], chunk.columns()[0][1][both])
    finally:
        scan.close()
        g1.close()
        g2.close()


def test_runtime_filter_errors_are_loud(gpu, ctx):
    with pytest.raises(gpu.GpuError):
        gpu.RuntimeFilter(ctx, abi.TYPE_DOUBLE, 10)
    rf = gpu.RuntimeFilter(ctx, abi.TYPE_INT, 10)
    with pytest.raises(gpu.GpuError):
        rf.insert(Chunk([(0, np.arange(4, dtype=np.int32), None)]), 9)  # slot not in the chunk
    scan = gpu.Scan(ctx, abi.ScanDesc(preds=[], filter_exprs=[], out_slots=[0]))
    for _ in range(4):
        scan.add_runtime_filter(rf, 0)
    with pytest.raises(gpu.GpuError):
        scan.add_runtime_filter(rf, 0)
    scan.close()
    rf.close()


def test_reference_evaluate_and_fill_vectors(gpu, ctx):
    # the known answers of runtime_filter_core_test.cpp:125-163,227-263 through the CUDA path
    from tests.test_oracle_golden import _rf_reference_evaluate_vectors
    made = []

    def make(n):
        made.append(gpu.RuntimeFilter(ctx, abi.TYPE_INT, n))
        return made[-1]
    try:
        _rf_reference_evaluate_vectors(make)
    finally:
        for f in made:
            f.close()
