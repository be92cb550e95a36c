"""shared helpers for the parity tests (the implementation lives in the package: bench.py and smoke() use it too)"""
from starrocks_b200.rows import assert_rows_equal, col_to_py, gpu_rows, oracle_rows, rand_nulls, rows_sorted  # noqa: F401
