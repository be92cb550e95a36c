"""Two-phase / streaming aggregation: the first phase pre-aggregates batches or passes rows through in the
intermediate format (AggregateStreamingSinkOperator, aggregate_streaming_sink_operator.cpp:80-372), the merge phase combines
the states (AggregateFunction::merge).  Whatever mix of the two legs the first phase takes, the merged result must equal
the single-phase result: bit-exact for integer work, 1e-6 relative for double SUM / AVG (summation order differs).

CPU tests run the oracle only; GPU tests run the CUDA path through the C-ABI against the oracle.
"""
import ctypes as C

import numpy as np
import pytest

from starrocks_b200 import abi
from starrocks_b200.abi import Chunk
from tests.helpers import assert_rows_equal, gpu_rows, oracle_rows, rand_nulls


def _case(name, n, rng):
    k1 = rng.integers(0, 7, n, dtype=np.int32)
    k2 = rng.integers(-3, 4, n).astype(np.int16)
    k3 = rng.integers(0, 5000, n, dtype=np.int64)
    v32 = rng.integers(-10**6, 10**6, n, dtype=np.int32)
    v64 = rng.integers(-10**12, 10**12, n, dtype=np.int64)
    vd = rng.normal(100, 50, n)
    vf = rng.random(n).astype(np.float32)
    nullv = "nullval" in name
    cols = [(0, k1, rand_nulls(rng, n, 0.05) if "nullkey" in name else None), (1, k2, None), (2, k3, None),
            (3, v32, rand_nulls(rng, n, 0.3) if nullv else None), (4, v64, None), (5, vd, rand_nulls(rng, n, 0.1) if nullv else None),
            (6, vf, None)]
    fns = [(abi.AGG_SUM, abi.TYPE_INT, 20, [("col", 3)]), (abi.AGG_COUNT, abi.TYPE_INT, 21, [("col", 3)]),
           (abi.AGG_COUNT_STAR, abi.TYPE_INT, 22, None), (abi.AGG_AVG, abi.TYPE_INT, 23, [("col", 3)]),
           (abi.AGG_MIN, abi.TYPE_BIGINT, 24, [("col", 4)]), (abi.AGG_MAX, abi.TYPE_FLOAT, 25, [("col", 6)]),
           (abi.AGG_AVG, abi.TYPE_DOUBLE, 26, [("col", 5), ("col", 6), "*"])]   # 9 first-phase functions would not fit: 7 + 2 AVG halves = 9
    fns = fns[:6] + [(abi.AGG_SUM, abi.TYPE_DOUBLE, 26, [("col", 5)])] if "sumd" in name else fns[1:]   # keep <= 8 first-phase fns
    if name.startswith("nogroup"):
        d = abi.make_agg_desc(fns=fns)
        nk = 0
    elif name.startswith("dense"):
        d = abi.make_agg_desc([0, 1], [abi.TYPE_INT, abi.TYPE_SMALLINT], fns=fns, ranges=[(0, 6), (-3, 3)],
                              group_nullable=[1 if "nullkey" in name else 0, 0])
        nk = 2
    else:
        d = abi.make_agg_desc([2, 0], [abi.TYPE_BIGINT, abi.TYPE_INT], fns=fns, group_nullable=[0, 1 if "nullkey" in name else 0])
        nk = 2
    fl = tuple(nk + i for i, f in enumerate(fns) if f[0] == abi.AGG_AVG or (f[0] == abi.AGG_SUM and f[1] == abi.TYPE_DOUBLE))
    return d, cols, fl


CASES = ["nogroup", "nogroup_nullval_sumd", "dense", "dense_nullkey_nullval", "hash", "hash_nullkey_nullval_sumd"]


def _batches(rng, n):
    cuts = sorted(set([0, n] + rng.integers(0, n + 1, 5).tolist()))
    return list(zip(cuts[:-1], cuts[1:])) or [(0, 0)]


def _sub(cols, lo, hi):
    return Chunk([(c[0], c[1][lo:hi].copy(), None if c[2] is None else c[2][lo:hi].copy()) + tuple(c[3:]) for c in cols])


def _state_chunk(cols4):
    """[(slot, type, data, nulls)] -> Chunk"""
    return Chunk([(s, np.ascontiguousarray(a), nl, t) for s, t, a, nl in cols4])


def _oracle_states_of_preagg(oracle, p1, chunk):
    a = oracle.Agg(p1)
    a.push(chunk)
    slots = [p1.group_slots[k] for k in range(p1.num_group_keys)] + [p1.fns[f].out_slot for f in range(p1.num_fns)]
    return [(s, t, d, nl) for s, (t, d, nl) in zip(slots, a.output())]


def test_python_and_c_plan_rewrite_agree(gpu):
    # sr_agg_two_phase_descs is host-only code: callable without a device
    rng = np.random.default_rng(1)
    for name in CASES:
        d, _, _ = _case(name, 8, rng)
        a1, a2 = abi.two_phase_descs(d)
        b1, b2 = gpu.two_phase_descs(d)
        assert bytes(a1) == bytes(b1) and bytes(a2) == bytes(b2), name
    dec = abi.make_agg_desc(fns=[(abi.AGG_SUM, abi.TYPE_DECIMAL64, 1, [("col", 0)])])
    with pytest.raises(NotImplementedError):
        abi.two_phase_descs(dec)       # 128-bit sum states
    with pytest.raises(gpu.GpuError):
        gpu.two_phase_descs(dec)
    many = abi.make_agg_desc(fns=[(abi.AGG_AVG, abi.TYPE_INT, 10 + k, [("col", 0)]) for k in range(5)])
    with pytest.raises(NotImplementedError):
        abi.two_phase_descs(many)      # 10 first-phase functions


def test_avg_merge_known_answer(oracle):
    # AVG over three first-phase batches, one of them all-NULL: (1 + 3 + 5) / 3, and a group that only ever saw NULLs
    d = abi.make_agg_desc([0], [abi.TYPE_INT], fns=[(abi.AGG_AVG, abi.TYPE_INT, 9, [("col", 1)]), (abi.AGG_COUNT, abi.TYPE_INT, 8, [("col", 1)])])
    p1, p2 = abi.two_phase_descs(d)
    assert [p1.fns[k].kind for k in range(p1.num_fns)] == [abi.AGG_SUM, abi.AGG_COUNT, abi.AGG_COUNT]
    assert p2.fns[0].kind == abi.AGG_AVG_MERGE and p2.fns[0].reserved == abi.agg_state_slot(9) and p2.fns[1].kind == abi.AGG_SUM
    batches = [([7, 7, 8], [1, 0, 0], [0, 1, 1]), ([7, 8], [3, 0, 0], [0, 1]), ([7], [5], [0]), ([7], [0], [1])]
    final = oracle.Agg(p2)
    for i, (k, v, nl) in enumerate(batches):
        ch = Chunk([(0, np.array(k, dtype=np.int32), None), (1, np.array(v, dtype=np.int32), np.array(nl, dtype=np.uint8))])
        states = oracle.convert_to_states(p1, ch) if i % 2 else _oracle_states_of_preagg(oracle, p1, ch)
        final.push(_state_chunk(states))
    assert oracle_rows(final) == [(7, 3.0, 3), (8, None, 0)]


# be/test/exprs/agg/aggregate_test.cpp:47-59 (count), :61-83 (sum), :115-132 (avg), :669-686 (max), :705-722 (min) through
# test_agg_function (base_aggregate_test.h:83-113,203-238): state1 = update(column1 = {0..1023, 100, 200}), state2 =
# update(column2 = {2000..2999}), then state1 is SERIALISED and MERGED into state2.  Expected (result1, result2, merged):
REFERENCE_MERGE_VECTORS = {
    abi.AGG_COUNT: (1026, 1000, 2026),
    abi.AGG_SUM: (524076, 2499500, 3023576),
    abi.AGG_AVG: (524076 / 1026.0, 2499500 / 1000.0, 3023576 / 2026.0),
    abi.AGG_MAX: (1023, 2999, 2999),
    abi.AGG_MIN: (0, 2000, 0),
}
REFERENCE_MERGE_TYPES = [(np.int16, abi.TYPE_SMALLINT), (np.int32, abi.TYPE_INT), (np.int64, abi.TYPE_BIGINT), (np.float32, abi.TYPE_FLOAT),
                         (np.float64, abi.TYPE_DOUBLE)]


def _reference_merge_case(kind, dtype, typ):
    col1 = np.concatenate([np.arange(1024), [100, 200]]).astype(dtype)
    col2 = np.arange(2000, 3000).astype(dtype)
    d = abi.make_agg_desc(fns=[(kind, typ, 9, [("col", 0)])])
    return d, Chunk([(0, col1, None, typ)]), Chunk([(0, col2, None, typ)])


@pytest.mark.parametrize("kind", sorted(REFERENCE_MERGE_VECTORS))
@pytest.mark.parametrize("dtype,typ", REFERENCE_MERGE_TYPES)
def test_reference_serialize_merge_vectors_oracle(oracle, kind, dtype, typ):
    d, c1, c2 = _reference_merge_case(kind, dtype, typ)
    r1, r2, merged = REFERENCE_MERGE_VECTORS[kind]
    for ch, want in ((c1, r1), (c2, r2)):
        a = oracle.Agg(d)
        a.push(ch)
        assert oracle_rows(a) == [(want,)]
    p1, p2 = abi.two_phase_descs(d)
    final = oracle.Agg(p2)
    for ch in (c2, c1):     # state2 first, then the serialised state1 merged into it
        final.push(_state_chunk(_oracle_states_of_preagg(oracle, p1, ch)))
    assert oracle_rows(final) == [(merged,)]


@pytest.mark.gpu
@pytest.mark.parametrize("kind", sorted(REFERENCE_MERGE_VECTORS))
@pytest.mark.parametrize("dtype,typ", REFERENCE_MERGE_TYPES)
def test_reference_serialize_merge_vectors_gpu(gpu, ctx, kind, dtype, typ):
    d, c1, c2 = _reference_merge_case(kind, dtype, typ)
    merged = REFERENCE_MERGE_VECTORS[kind][2]
    p1, p2 = gpu.two_phase_descs(d)
    first, final = gpu.Agg(ctx, p1), gpu.Agg(ctx, p2)
    try:
        for ch in (c2, c1):
            first.reset()
            first.push(ch)
            first.finish()
            final.push(gpu.chunk_out_as_view(first.pull(mem=abi.MEM_DEVICE)))
        assert gpu_rows(final.result()) == [(merged,)]
    finally:
        first.close()
        final.close()


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("n", [0, 1, 5000])
def test_oracle_two_phase_equals_single_phase(oracle, name, n):
    rng = np.random.default_rng(17 + n)
    d, cols, fl = _case(name, n, rng)
    single = oracle.Agg(d)
    single.push(_sub(cols, 0, n))
    p1, p2 = abi.two_phase_descs(d)
    final = oracle.Agg(p2)
    for lo, hi in _batches(rng, n):
        ch = _sub(cols, lo, hi)
        # a query without GROUP BY has a blocking first phase (one state row per instance, even for no input); the
        # pass-through leg only exists for grouped aggregates
        stream = p1.num_group_keys > 0 and rng.random() < 0.5
        states = oracle.convert_to_states(p1, ch) if stream else _oracle_states_of_preagg(oracle, p1, ch)
        final.push(_state_chunk(states))
    assert_rows_equal(oracle_rows(final), oracle_rows(single), float_cols=fl)


# ---------------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("n", [0, 1, 1000, 70001])
def test_gpu_convert_to_states_parity(gpu, ctx, oracle, name, n):
    # per-row projection: every byte must agree (doubles too: no reassociation involved)
    rng = np.random.default_rng(23 + n)
    d, cols, _ = _case(name, n, rng)
    p1, _ = abi.two_phase_descs(d)
    ch = _sub(cols, 0, n)
    a = gpu.Agg(ctx, p1)
    try:
        got = gpu.chunk_out_to_host(ctx, a.convert_to_states(ch))
        exp = oracle.convert_to_states(p1, ch)
        assert len(got) == len(exp)
        for (gs, gt, gd, gn), (es, et, ed, en) in zip(got, exp):
            assert (gs, gt) == (es, et)
            if en is not None:
                assert gn is not None and np.array_equal(gn, en)
                keep = en == 0
                assert np.array_equal(np.asarray(gd)[keep], np.asarray(ed)[keep])
            else:
                assert np.array_equal(np.asarray(gd), np.asarray(ed))
    finally:
        a.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("n", [0, 1, 1000, 200_003])
def test_gpu_two_phase_equals_oracle_single_phase(gpu, ctx, oracle, name, n):
    # first phase on the device (pre-aggregated batches pulled as DEVICE chunks, or rows passed through), merge phase on
    # the device fed with those device chunks, against the single-phase oracle
    rng = np.random.default_rng(29 + n)
    d, cols, fl = _case(name, n, rng)
    single = oracle.Agg(d)
    single.push(_sub(cols, 0, n))
    p1, p2 = gpu.two_phase_descs(d)
    first, final = gpu.Agg(ctx, p1), gpu.Agg(ctx, p2)
    try:
        for lo, hi in _batches(rng, n):
            ch = _sub(cols, lo, hi)
            if p1.num_group_keys > 0 and rng.random() < 0.5:
                out = first.convert_to_states(ch)
            else:
                first.reset()
                first.push(ch)
                first.finish()
                out = first.pull(mem=abi.MEM_DEVICE)
            if out.num_rows:
                final.push(gpu.chunk_out_as_view(out))
        assert_rows_equal(gpu_rows(final.result()), oracle_rows(single), float_cols=fl)
    finally:
        first.close()
        final.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("n", [1000, 150_001])
def test_gpu_selective_preaggregation(gpu, ctx, oracle, name, n):
    # SELECTIVE_PREAGG (aggregate_streaming_sink_operator.cpp:173-210): after a warm-up batch has filled the table, every
    # further batch is split by build_hash_map_with_selection -- rows of known groups are aggregated in place, the others
    # are streamed out in the intermediate format.  (1) the streamed rows are exactly the oracle's selection (bytes of the
    # oracle's convert_to_states at those rows, in order); (2) table + streamed rows merged = the single-phase result.
    rng = np.random.default_rng(31 + n)
    d, cols, fl = _case(name, n, rng)
    single = oracle.Agg(d)
    single.push(_sub(cols, 0, n))
    p1, p2 = gpu.two_phase_descs(d)
    first, final, ofirst = gpu.Agg(ctx, p1), gpu.Agg(ctx, p2), oracle.Agg(p1)
    try:
        warm = n // 5
        first.push(_sub(cols, 0, warm))
        ofirst.push(_sub(cols, 0, warm))
        streamed = 0
        for lo, hi in ((warm, n // 2), (n // 2, n)):
            ch = _sub(cols, lo, hi)
            sel = ofirst.streaming_selection(ch)
            if name.startswith("dense"):
                sel[:] = 0      # a range-declared table has every group's slot from the start: nothing is ever streamed
            exp = oracle.convert_to_states(p1, ch)
            out = first.push_selective(ch)
            assert out.num_rows == int(sel.sum())
            got = gpu.chunk_out_to_host(ctx, out)
            for (gs, gt, gd, gn), (es, et, ed, en) in zip(got, exp):
                assert (gs, gt) == (es, et)
                keep = sel == 1
                if en is not None:
                    assert np.array_equal(gn, en[keep])
                    ok = en[keep] == 0
                    assert np.array_equal(np.asarray(gd)[ok], np.asarray(ed)[keep][ok])
                else:
                    assert np.array_equal(np.asarray(gd), np.asarray(ed)[keep])
            # the oracle's table takes the rows of known groups (what compute_batch_agg_states_with_selection does)
            known = sel == 0
            if known.any():
                ofirst.push(Chunk([(c[0], c[1][lo:hi][known].copy(), None if c[2] is None else c[2][lo:hi][known].copy()) + tuple(c[3:]) for c in cols]))
            if out.num_rows:
                final.push(gpu.chunk_out_as_view(out))
            streamed += out.num_rows
        if p1.num_group_keys > 0 and not name.startswith("dense"):
            assert streamed > 0                     # the hash cases really stream rows of unseen groups
        first.finish()
        tab = first.pull(mem=abi.MEM_DEVICE)
        if tab.num_rows:
            final.push(gpu.chunk_out_as_view(tab))
        assert_rows_equal(gpu_rows(final.result()), oracle_rows(single), float_cols=fl)
    finally:
        first.close()
        final.close()


@pytest.mark.gpu
def test_avg_merge_errors_are_loud(gpu, ctx):
    bad = abi.make_agg_desc(fns=[(abi.AGG_AVG_MERGE, abi.TYPE_DOUBLE, 9, [("col", 1)])])
    bad.fns[0].reserved = 77      # count state slot that is not in the chunk
    a = gpu.Agg(ctx, bad)
    with pytest.raises(gpu.GpuError):
        a.push(Chunk([(1, np.zeros(4), None)]))
    a.close()
    single = gpu.Agg(ctx, abi.make_agg_desc(fns=[(abi.AGG_AVG, abi.TYPE_INT, 9, [("col", 1)])]))
    with pytest.raises(gpu.GpuError):
        single.convert_to_states(Chunk([(1, np.zeros(4, dtype=np.int32), None)]))   # needs a first-phase desc
    single.close()
