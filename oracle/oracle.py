"""ctypes binding of the CPU oracle (oracle/libsr_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (starrocks_b200) never does.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from starrocks_b200 import abi  # noqa: E402

_LIB = None


class orc_join_options(C.Structure):
    _fields_ = [("enable_range_direct_mapping", C.c_int32), ("enable_linear_chained", C.c_int32),
                ("l2_cache_size", C.c_int64), ("l3_cache_size", C.c_int64), ("force_method", C.c_int32),
                ("chunk_size", C.c_int32)]


class orc_probe_result(C.Structure):
    _fields_ = [("count", C.c_int64), ("has_remain", C.c_int32), ("match_flag", C.c_int32),
                ("cur_probe_index", C.c_int32), ("cur_row_match_count", C.c_int32)]


class orc_frag_join(C.Structure):
    _fields_ = [("join", C.c_void_p), ("probe_key_slot", C.c_int32), ("num_payload", C.c_int32),
                ("payload_build_slots", C.c_int32 * abi.SR_MAX_FRAG_PAYLOAD)]


class orc_fragment_desc(C.Structure):
    _fields_ = [("scan", abi.sr_scan_desc), ("num_joins", C.c_int32), ("reserved", C.c_int32),
                ("joins", orc_frag_join * abi.SR_MAX_FRAG_JOINS), ("agg", abi.sr_agg_desc)]


BUCKET_CHAINED, DIRECT_MAPPING, RANGE_DIRECT_MAPPING, RANGE_DIRECT_MAPPING_SET = 1, 2, 3, 4
DENSE_RANGE_DIRECT_MAPPING, LINEAR_CHAINED, LINEAR_CHAINED_SET = 5, 6, 7


def _source_hash():
    import hashlib
    h = hashlib.sha256()
    for f in ("sr_oracle.cpp", "sr_oracle.h", os.path.join("..", "include", "sr_gpu_ops.h")):
        with open(os.path.join(_HERE, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _recorded_hash():
    try:
        with open(os.path.join(_HERE, "libsr_oracle.so.srchash")) as fh:
            return fh.read().strip()
    except OSError:
        return None


def build():
    """compile oracle/libsr_oracle.so (g++, seconds) and, where the reference tree is present, oracle/_ref (the pieces of the
    path that compile from the reference's own sources: the vendored xxHash, the frame-of-reference page codec)."""
    import fcntl
    with open(os.path.join(_HERE, "libsr_oracle.so.lock"), "w") as lock:      # the ranks of a torchrun job may all land here
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not (os.path.exists(os.path.join(_HERE, "libsr_oracle.so")) and _recorded_hash() == _source_hash()):
                subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libsr_oracle.so"])
                with open(os.path.join(_HERE, "libsr_oracle.so.srchash"), "w") as fh:
                    fh.write(_source_hash())
            subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def ref_xxh3():
    """oracle/_ref/libxxh3_ref.so (HashUtil::xx_hash3_64 from the reference's own header), or None when it was never built"""
    path = os.path.join(_HERE, "_ref", "libxxh3_ref.so")
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    L.ref_xx_hash3_64.restype = C.c_uint64
    L.ref_xx_hash3_64.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
    return L


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(_HERE, "libsr_oracle.so")
    # stale = the sources' content differs from what the library was built from (file times do not survive a copy of the tree)
    if not os.path.exists(path) or _recorded_hash() != _source_hash():
        build()
    L = C.CDLL(path)
    u32, i32, i64, vp = C.c_uint32, C.c_int32, C.c_int64, C.c_void_p
    sig = {
        "orc_join_key_hash32": (u32, [u32, u32]),
        "orc_join_key_hash64": (u32, [C.c_uint64, u32]),
        "orc_join_key_hash_slice": (u32, [vp, i32, u32]),
        "orc_calc_bucket_size": (u32, [u32]),
        "orc_crc32c": (u32, [vp, i32, u32]),
        "orc_crc_hash_32": (u32, [vp, i32, u32]),
        "orc_zlib_crc32": (u32, [vp, i32, u32]),
        "orc_fnv_hash": (u32, [vp, i32, u32]),
        "orc_xxh3_64": (C.c_uint64, [vp, i32, C.c_uint64]),
        "orc_chunk_serialize": (C.c_int64, [vp, C.c_int64, C.c_int64, vp, C.c_int64]),
        "orc_xorshift32": (u32, [u32]),
        "orc_reduce_op": (u32, [u32, u32]),
        "orc_filter_range": (i64, [vp, vp, i32, i64, i64]),
        "orc_scan_evaluate": (i32, [vp, vp, vp]),
        "orc_scan_filter": (i64, [vp, vp, vp, vp]),
        "orc_eval_expr": (i32, [vp, vp, vp, vp, vp, vp]),
        "orc_join_create": (vp, [vp, vp]),
        "orc_join_destroy": (None, [vp]),
        "orc_join_append_build": (i32, [vp, vp]),
        "orc_join_build": (i32, [vp]),
        "orc_join_method": (i32, [vp]),
        "orc_join_build_rows": (i64, [vp]),
        "orc_join_bucket_size": (i64, [vp]),
        "orc_join_min_value": (i64, [vp]),
        "orc_join_max_value": (i64, [vp]),
        "orc_join_first": (vp, [vp]),
        "orc_join_next": (vp, [vp]),
        "orc_join_probe_chunk": (i32, [vp, vp, i32, vp, vp, vp]),
        "orc_join_probe_all": (i64, [vp, vp, vp, vp, i64]),
        "orc_join_output": (i32, [vp, vp, i64, vp, vp, vp, vp]),
        "orc_join_probe_remain": (i64, [vp, vp, i64]),
        "orc_join_output_remain": (i32, [vp, i64, vp, vp, vp, vp]),
        "orc_agg_create": (vp, [vp]),
        "orc_agg_destroy": (None, [vp]),
        "orc_agg_push": (i32, [vp, vp]),
        "orc_agg_num_groups": (i64, [vp]),
        "orc_agg_output": (i32, [vp, vp, vp]),
        "orc_agg_out_type": (i32, [vp, i32]),
        "orc_agg_num_out_cols": (i32, [vp]),
        "orc_agg_merge": (i32, [vp, vp]),
        "orc_agg_convert_to_states": (i32, [vp, vp, vp, vp]),
        "orc_hash_partition": (i32, [vp, vp, vp, vp, vp, vp]),
        "orc_fragment_run": (i32, [vp, vp, i32, vp, vp]),
        "orc_rf_create": (vp, [i32, i64, i32]),
        "orc_rf_destroy": (None, [vp]),
        "orc_rf_insert_hash": (None, [vp, C.c_uint64]),
        "orc_rf_test_hash": (i32, [vp, C.c_uint64]),
        "orc_rf_value_hash": (C.c_uint64, [i64]),
        "orc_rf_insert": (i32, [vp, vp, i32, i32]),
        "orc_rf_merge": (i32, [vp, vp]),
        "orc_rf_evaluate": (i32, [vp, vp, i32, vp, i32]),
        "orc_rf_get_info": (i32, [vp, vp]),
        "orc_rf_directory": (vp, [vp, vp]),
        "orc_last_error": (C.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = L
    return L


class OracleError(RuntimeError):
    pass


def _check(rc):
    if rc is not None and rc < 0:
        raise OracleError(f"oracle error {rc}: {lib().orc_last_error().decode()}")
    return rc


def _np_ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _alloc_col(typ, n):
    w = abi.TYPE_WIDTH[typ]
    if w == 16:
        return np.zeros(n, dtype=np.dtype((np.void, 16)))
    return np.zeros(n, dtype=abi.TYPE_NUMPY[typ])


def i128_to_py(arr):
    """numpy void16 array -> list of python ints (little endian two's complement)."""
    raw = arr.tobytes()
    return [int.from_bytes(raw[i * 16:(i + 1) * 16], "little", signed=True) for i in range(len(arr))]


def scan_evaluate(scan_desc, chunk):
    sel = np.zeros(chunk.num_rows, dtype=np.uint8)
    _check(lib().orc_scan_evaluate(scan_desc.ref(), chunk.ref(), _np_ptr(sel)))
    return sel


def scan_filter(scan_desc, chunk):
    """-> (rows, {slot: (data, nulls_or_None)})"""
    n = chunk.num_rows
    outs, nulls = [], []
    for s in scan_desc.out_slots:
        k = chunk.slots.index(s)
        outs.append(_alloc_col(chunk.types[k], n))
        nulls.append(np.zeros(n, dtype=np.uint8) if chunk.view.cols[k].nulls else None)
    pd = (C.c_void_p * max(1, len(outs)))(*[o.ctypes.data for o in outs])
    pn = (C.c_void_p * max(1, len(outs)))(*[(x.ctypes.data if x is not None else None) for x in nulls])
    rows = _check(lib().orc_scan_filter(scan_desc.ref(), chunk.ref(), pd, pn))
    res = {}
    for s, o, x in zip(scan_desc.out_slots, outs, nulls):
        res[s] = (o[:rows], None if x is None else x[:rows])
    return rows, res


def eval_expr(expr, chunk):
    n = chunk.num_rows
    oi = np.zeros(n, dtype=np.int64)
    od = np.zeros(n, dtype=np.float64)
    on = np.zeros(n, dtype=np.uint8)
    isd = C.c_int32(0)
    _check(lib().orc_eval_expr(C.byref(expr), chunk.ref(), _np_ptr(oi), _np_ptr(od), _np_ptr(on), C.byref(isd)))
    return (od if isd.value else oi), on


class Join:
    def __init__(self, desc, options=None, force_method=0, chunk_size=0, l2=1 << 20, l3=32 << 20):
        self.desc = desc
        if options is None:
            options = orc_join_options(desc.enable_range_direct_mapping, 1, l2, l3, force_method, chunk_size)
        self.chunk_size = chunk_size or 4096
        self.h = lib().orc_join_create(C.byref(desc), C.byref(options))
        if not self.h:
            raise OracleError(lib().orc_last_error().decode())
        self.build_types = {}

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_join_destroy(self.h)
            self.h = None

    def append_build(self, chunk):
        for s, t in zip(chunk.slots, chunk.types):
            self.build_types[s] = t
        _check(lib().orc_join_append_build(self.h, chunk.ref()))

    def build(self):
        _check(lib().orc_join_build(self.h))

    @property
    def method(self):
        return lib().orc_join_method(self.h)

    @property
    def build_rows(self):
        return lib().orc_join_build_rows(self.h)

    @property
    def bucket_size(self):
        return lib().orc_join_bucket_size(self.h)

    @property
    def min_value(self):
        return lib().orc_join_min_value(self.h)

    @property
    def max_value(self):
        return lib().orc_join_max_value(self.h)

    def first(self):
        n = self.bucket_size
        if self.method == DENSE_RANGE_DIRECT_MAPPING:
            n = self.build_rows + 1
        p = lib().orc_join_first(self.h)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(n,)).copy()

    def next(self):
        p = lib().orc_join_next(self.h)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(self.build_rows + 1,)).copy()

    def probe_chunk(self, chunk, first_probe=True):
        pi = np.zeros(self.chunk_size + 8, dtype=np.uint32)
        bi = np.zeros(self.chunk_size + 8, dtype=np.uint32)
        res = orc_probe_result()
        _check(lib().orc_join_probe_chunk(self.h, chunk.ref(), 1 if first_probe else 0, _np_ptr(pi), _np_ptr(bi),
                                          C.byref(res)))
        return pi[:res.count].copy(), bi[:res.count].copy(), res

    def probe_all(self, chunk, cap=None):
        cap = cap if cap is not None else max(1024, chunk.num_rows * 2)
        while True:
            pi = np.zeros(cap, dtype=np.uint32)
            bi = np.zeros(cap, dtype=np.uint32)
            n = lib().orc_join_probe_all(self.h, chunk.ref(), _np_ptr(pi), _np_ptr(bi), cap)
            if n >= 0:
                return pi[:n], bi[:n]
            if -n <= cap:
                _check(int(n))
            cap = int(-n)

    def output(self, chunk, pi, bi):
        """-> list of (slot, data, nulls) for probe_out_slots + build_out_slots"""
        d = self.desc
        n = len(pi)
        semi = d.join_type in (abi.JOIN_LEFT_SEMI, abi.JOIN_LEFT_ANTI)
        slots, outs, nulls = [], [], []
        for k in range(d.num_probe_out):
            s = d.probe_out_slots[k]
            t = chunk.types[chunk.slots.index(s)]
            slots.append(s)
            outs.append(_alloc_col(t, n))
            nulls.append(np.zeros(n, dtype=np.uint8))
        if not semi:
            for k in range(d.num_build_out):
                s = d.build_out_slots[k]
                slots.append(s)
                outs.append(_alloc_col(self.build_types[s], n))
                nulls.append(np.zeros(n, dtype=np.uint8))
        pd = (C.c_void_p * max(1, len(outs)))(*[o.ctypes.data for o in outs])
        pn = (C.c_void_p * max(1, len(outs)))(*[x.ctypes.data for x in nulls])
        pi = np.ascontiguousarray(pi, dtype=np.uint32)
        bi = np.ascontiguousarray(bi, dtype=np.uint32)
        _check(lib().orc_join_output(self.h, chunk.ref(), n, _np_ptr(pi), _np_ptr(bi), pd, pn))
        return list(zip(slots, outs, nulls))


    def probe_remain(self, probe_types=()):
        """POST_PROBE rows of a RIGHT / FULL join -> list of (slot, data, nulls): NULL probe_out columns (types given), then build_out"""
        d = self.desc
        cap = self.build_rows + 1
        bi = np.zeros(cap, dtype=np.uint32)
        n = _check(lib().orc_join_probe_remain(self.h, _np_ptr(bi), cap))
        bi = bi[:n]
        with_probe = d.join_type in (abi.JOIN_RIGHT_OUTER, abi.JOIN_FULL_OUTER)
        slots, outs, nulls = [], [], []
        if with_probe:
            for k in range(d.num_probe_out):
                slots.append(d.probe_out_slots[k])
                outs.append(_alloc_col(probe_types[k], n))
                nulls.append(np.zeros(n, dtype=np.uint8))
        for k in range(d.num_build_out):
            s = d.build_out_slots[k]
            slots.append(s)
            outs.append(_alloc_col(self.build_types[s], n))
            nulls.append(np.zeros(n, dtype=np.uint8))
        pd = (C.c_void_p * max(1, len(outs)))(*[o.ctypes.data for o in outs])
        pn = (C.c_void_p * max(1, len(outs)))(*[x.ctypes.data for x in nulls])
        pt = (C.c_int32 * max(1, len(probe_types)))(*probe_types)
        _check(lib().orc_join_output_remain(self.h, n, _np_ptr(bi), pt, pd, pn))
        return list(zip(slots, outs, nulls))


class Agg:
    def __init__(self, desc):
        self.desc = desc
        self.h = lib().orc_agg_create(C.byref(desc))
        if not self.h:
            raise OracleError(lib().orc_last_error().decode())

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_agg_destroy(self.h)
            self.h = None

    def push(self, chunk):
        _check(lib().orc_agg_push(self.h, chunk.ref()))

    def merge(self, other):
        _check(lib().orc_agg_merge(self.h, other.h))

    def streaming_selection(self, chunk):
        """build_hash_map_with_selection: uint8 per row, 1 = the row's group is not in the table"""
        sel = np.zeros(chunk.num_rows, dtype=np.uint8)
        L = lib()
        L.orc_agg_streaming_selection.restype = C.c_int32
        L.orc_agg_streaming_selection.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _check(L.orc_agg_streaming_selection(self.h, chunk.ref(), sel.ctypes.data if sel.size else None))
        return sel

    @property
    def num_groups(self):
        return lib().orc_agg_num_groups(self.h)

    def output(self):
        """-> list of (type, data ndarray, nulls ndarray) keys first then results"""
        n = self.num_groups
        nc = lib().orc_agg_num_out_cols(self.h)
        types = [lib().orc_agg_out_type(self.h, k) for k in range(nc)]
        outs = [_alloc_col(t, n) for t in types]
        nulls = [np.zeros(n, dtype=np.uint8) for _ in types]
        pd = (C.c_void_p * max(1, nc))(*[o.ctypes.data for o in outs])
        pn = (C.c_void_p * max(1, nc))(*[x.ctypes.data for x in nulls])
        _check(lib().orc_agg_output(self.h, pd, pn))
        return list(zip(types, outs, nulls))


def convert_to_states(first_phase_desc, chunk):
    """pass-through leg of the streaming aggregate -> list of (slot, type, data, nulls_or_None): group-by columns of the
    chunk as they are, then one state column per function (orc_agg_convert_to_states)"""
    d = first_phase_desc
    n = chunk.num_rows
    cols = {s: (t, a, nl) for (s, a, nl), t in zip(chunk.columns(), chunk.types)}
    res = [(d.group_slots[k],) + cols[d.group_slots[k]] for k in range(d.num_group_keys)]
    types = [abi.agg_result_type(d.fns[f].kind, d.fns[f].input_type) for f in range(d.num_fns)]
    outs = [_alloc_col(t, n) for t in types]
    nulls = [np.zeros(n, dtype=np.uint8) for _ in types]
    pd = (C.c_void_p * max(1, d.num_fns))(*[o.ctypes.data for o in outs])
    pn = (C.c_void_p * max(1, d.num_fns))(*[x.ctypes.data for x in nulls])
    _check(lib().orc_agg_convert_to_states(C.byref(d), chunk.ref(), pd, pn))
    for f in range(d.num_fns):
        countlike = d.fns[f].kind in (abi.AGG_COUNT, abi.AGG_COUNT_STAR)
        res.append((d.fns[f].out_slot, types[f], outs[f], None if countlike else nulls[f]))
    return res


def hash_partition(part_desc, chunk):
    n = chunk.num_rows
    hv = np.zeros(n, dtype=np.uint32)
    ch = np.zeros(n, dtype=np.uint32)
    ri = np.zeros(n, dtype=np.uint32)
    st = np.zeros(part_desc.num_channels + 1, dtype=np.int64)
    _check(lib().orc_hash_partition(C.byref(part_desc), chunk.ref(), _np_ptr(hv), _np_ptr(ch), _np_ptr(ri),
                                    _np_ptr(st)))
    return hv, ch, ri, st


class RuntimeFilter:
    """MinMaxRuntimeFilter + SimdBlockFilter restatement (orc_rf_*)"""

    def __init__(self, key_type, expected_rows, with_bloom=True):
        self.h = lib().orc_rf_create(key_type, expected_rows, 1 if with_bloom else 0)
        if not self.h:
            raise OracleError(lib().orc_last_error().decode())

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_rf_destroy(self.h)
            self.h = None

    def insert_hash(self, h):
        lib().orc_rf_insert_hash(self.h, h)

    def test_hash(self, h):
        return bool(lib().orc_rf_test_hash(self.h, h))

    def insert(self, chunk, slot, insert_nulls=False):
        _check(lib().orc_rf_insert(self.h, chunk.ref(), slot, 1 if insert_nulls else 0))

    def merge(self, other):
        _check(lib().orc_rf_merge(self.h, other.h))

    def evaluate(self, chunk, slot, selection=None):
        sel = np.zeros(chunk.num_rows, dtype=np.uint8) if selection is None else np.ascontiguousarray(selection, dtype=np.uint8).copy()
        _check(lib().orc_rf_evaluate(self.h, chunk.ref(), slot, _np_ptr(sel), 0 if selection is None else 1))
        return sel

    def info(self):
        inf = abi.sr_rf_info()
        _check(lib().orc_rf_get_info(self.h, C.byref(inf)))
        return inf

    def directory(self):
        n = C.c_int64(0)
        p = lib().orc_rf_directory(self.h, C.byref(n))
        if n.value == 0:
            return np.empty(0, dtype=np.uint32)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(n.value // 4,)).copy()


def value_hash(v):
    return int(lib().orc_rf_value_hash(int(v)))


def chunk_serialize(chunk, row_begin=0, row_end=None):
    """ChunkPB.data bytes (encode level 0) of rows [row_begin, row_end) -> uint8 ndarray"""
    row_end = chunk.num_rows if row_end is None else row_end
    n = lib().orc_chunk_serialize(chunk.ref(), row_begin, row_end, None, 0)
    if n < 0:
        _check(int(n))
    buf = np.zeros(n, dtype=np.uint8)
    n2 = lib().orc_chunk_serialize(chunk.ref(), row_begin, row_end, buf.ctypes.data, n)
    assert n2 == n
    return buf


def ref_hash():
    """oracle/_ref/libhash_ref.so (HashUtil::fnv_hash / zlib_crc_hash and crc_hash_32 of the reference), or None when it was never built"""
    path = os.path.join(_HERE, "_ref", "libhash_ref.so")
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    for nm in ("ref_fnv_hash", "ref_zlib_crc_hash", "ref_crc_hash_32"):
        getattr(L, nm).restype = C.c_uint32
        getattr(L, nm).argtypes = [C.c_void_p, C.c_int, C.c_uint32]
    return L


def ref_for():
    """oracle/_ref/libfor_ref.so (the reference's own ForEncoder / ForDecoder), or None when it was never built"""
    path = os.path.join(_HERE, "_ref", "libfor_ref.so")
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    for nm in ("ref_for_encode_i32", "ref_for_encode_i64", "ref_for_decode_i32", "ref_for_decode_i64"):
        getattr(L, nm).restype = C.c_longlong
        getattr(L, nm).argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong]
    return L


def _page_call(fn, elem_size, src, n_or_len, out_dtype, cap_guess):
    cap = cap_guess
    while True:
        out = np.zeros(max(cap, 1), dtype=out_dtype)
        r = fn(elem_size, src.ctypes.data if src.size else None, n_or_len, out.ctypes.data, cap)
        if r >= 0:
            return out[:r]
        if r == -1:
            raise OracleError("corrupt page")
        cap = int(-r)


def for_encode(values):
    """frame-of-reference page bytes of an int32 / int64 array (FrameOfReferencePageBuilder::finish)"""
    v = np.ascontiguousarray(values)
    L = lib()
    L.orc_for_encode.restype = C.c_int64
    L.orc_for_encode.argtypes = [C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
    return _page_call(L.orc_for_encode, v.dtype.itemsize, v, len(v), np.uint8, len(v) * (v.dtype.itemsize + 1) + 64)


def for_decode(page, dtype):
    p = np.ascontiguousarray(page, dtype=np.uint8)
    L = lib()
    L.orc_for_decode.restype = C.c_int64
    L.orc_for_decode.argtypes = [C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
    return _page_call(L.orc_for_decode, np.dtype(dtype).itemsize, p, len(p), dtype, 1024)


def plain_encode(values):
    v = np.ascontiguousarray(values)
    L = lib()
    L.orc_plain_encode.restype = C.c_int64
    L.orc_plain_encode.argtypes = [C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
    return _page_call(L.orc_plain_encode, v.dtype.itemsize, v, len(v), np.uint8, 4 + v.nbytes)


def plain_decode(page, dtype):
    p = np.ascontiguousarray(page, dtype=np.uint8)
    L = lib()
    L.orc_plain_decode.restype = C.c_int64
    L.orc_plain_decode.argtypes = [C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
    return _page_call(L.orc_plain_decode, np.dtype(dtype).itemsize, p, len(p), dtype, 1024)


def fragment_run(scan_desc, joins, agg_desc, fact_chunk, num_threads=1):
    """joins: list of (Join, probe_key_slot, [payload build slots]).  -> (Agg result, rows_passed)"""
    d = orc_fragment_desc()
    d.scan = scan_desc.desc
    d.num_joins = len(joins)
    for k, (j, slot, payload) in enumerate(joins):
        d.joins[k].join = j.h
        d.joins[k].probe_key_slot = slot
        d.joins[k].num_payload = len(payload)
        for q, s in enumerate(payload):
            d.joins[k].payload_build_slots[q] = s
    d.agg = agg_desc
    result = Agg(agg_desc)
    passed = C.c_int64(0)
    _check(lib().orc_fragment_run(C.byref(d), fact_chunk.ref(), num_threads, result.h, C.byref(passed)))
    return result, passed.value
