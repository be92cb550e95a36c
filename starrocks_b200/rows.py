"""Result-set helpers shared by the parity tests, bench.py, __graft_entry__.smoke() and the tools: columns handed back by the
C-ABI (or by the oracle) -> sorted python rows, and their comparison (bit-exact, or relative tolerance for DOUBLE sums)."""
import numpy as np

from starrocks_b200 import abi


def col_to_py(typ, data, nulls):
    """column -> list of python values (None for NULL); 128-bit values as python ints"""
    n = len(data)
    if abi.TYPE_WIDTH[typ] == 16:
        raw = data.tobytes()
        vals = [int.from_bytes(raw[i * 16:(i + 1) * 16], "little", signed=True) for i in range(n)]
    else:
        vals = data.tolist()
    if nulls is not None and len(nulls):
        vals = [None if nulls[i] else v for i, v in enumerate(vals)]
    return vals


def rows_sorted(cols):
    """cols: list of python-value lists -> rows sorted with None first"""
    rows = list(zip(*cols)) if cols else []
    return sorted(rows, key=lambda r: tuple((0, 0) if v is None else (1, v) for v in r))


def gpu_rows(result):
    """gpu.Agg.result() / chunk_out_to_host output -> sorted rows"""
    return rows_sorted([col_to_py(t, d, n) for _, t, d, n in result])


def oracle_rows(agg):
    out = agg.output()
    return rows_sorted([col_to_py(t, d, n) for t, d, n in out])


def assert_rows_equal(got, exp, float_cols=(), rel=1e-6):
    """bit-exact for every column except float_cols (relative tolerance, as BASELINE.json states)"""
    assert len(got) == len(exp), f"row count {len(got)} != {len(exp)}"
    for r, (g, e) in enumerate(zip(got, exp)):
        assert len(g) == len(e)
        for c, (gv, ev) in enumerate(zip(g, e)):
            if c in float_cols and gv is not None and ev is not None:
                assert abs(gv - ev) <= rel * max(abs(ev), 1e-300), f"row {r} col {c}: {gv} vs {ev}"
            else:
                assert gv == ev, f"row {r} col {c}: {gv} vs {ev}"


def rand_nulls(rng, n, frac=0.1):
    return (rng.random(n) < frac).astype(np.uint8)
