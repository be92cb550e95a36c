"""Host-side plumbing of the multi-GPU plan, backend agnostic (NCCL on GPUs, gloo in the CPU tests).

Two exchanges exist in the reference plans of the configs:
  * UNPARTITIONED gather of partial aggregate states to the final fragment instance (SSB Q4.1: <= 175 rows per rank);
  * HASH_PARTITIONED shuffle of rows between fragment instances (TPC-H Q3 / TPC-DS Q95): channel = ReduceOp(fnv_hash(key), n)
    (exchange_sink_operator.cpp:586-637, shuffler.h:72-89).  The device side (hash + stable partition) is sr_xchg_partition;
    this module moves the partitioned column buffers with all_to_all_single using the per-channel counts it produced.
Only torch.distributed calls live here -- no kernels.
"""
import torch
import torch.distributed as dist


def _to_i64_bits(c):
    """a state column as int64 WORDS: integers by value, floating-point states bit for bit (a DOUBLE sum must not be
    rounded to an integer on its way through the packed int64 exchange buffer)"""
    if c.dtype == torch.float64:
        return c.contiguous().view(torch.int64)
    if c.dtype == torch.float32:
        return c.to(torch.float64).view(torch.int64)      # exact; _from_i64_bits narrows it back
    if c.dtype.is_floating_point:
        raise TypeError(f"state column of dtype {c.dtype}")
    return c.to(torch.int64)


def _from_i64_bits(t, dtype):
    if dtype == torch.float64:
        return t.contiguous().view(torch.float64)
    if dtype == torch.float32:
        return t.contiguous().view(torch.float64).to(torch.float32)
    return t.to(dtype)


def gather_partial_states(cols, max_rows, dst=0):
    """cols: list of 1-D tensors (same length g <= max_rows) holding a rank's partial aggregate rows (group keys and
    states; integer or floating point, every rank passes the same dtypes).  Returns on `dst` a list (one entry per rank)
    of lists of tensors of the same dtypes trimmed to each rank's row count; None elsewhere."""
    world, rank = dist.get_world_size(), dist.get_rank()
    g = int(cols[0].numel())
    if g > max_rows:
        raise ValueError(f"{g} partial rows exceed the gather capacity {max_rows}")
    dev = cols[0].device
    part = torch.zeros((len(cols), max_rows), dtype=torch.int64, device=dev)
    for k, c in enumerate(cols):
        part[k, :g] = _to_i64_bits(c)
    cnt = torch.tensor([g], dtype=torch.int64, device=dev)
    parts = [torch.empty_like(part) for _ in range(world)] if rank == dst else None
    cnts = [torch.empty_like(cnt) for _ in range(world)] if rank == dst else None
    dist.gather(part, parts, dst=dst)
    dist.gather(cnt, cnts, dst=dst)
    if rank != dst:
        return None
    out = []
    for p, c in zip(parts, cnts):
        gg = int(c.item())
        out.append([_from_i64_bits(p[k, :gg], cols[k].dtype) for k in range(len(cols))])
    return out


class _DevicePtr:
    """zero-copy view of a device buffer handed back by the C-ABI (sr_chunk_out column) for torch.as_tensor"""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def device_view(ptr, n, width, device):
    """1-D integer tensor aliasing `n` values of `width` bytes at device address `ptr` (no copy)."""
    if n == 0:
        return torch.empty(0, dtype=torch.int32 if width == 4 else torch.int64, device=device)
    return torch.as_tensor(_DevicePtr(ptr, n, "<i4" if width == 4 else "<i8"), device=device)


def all_gather_partial_states(cols, max_rows, dst=0):
    """Same exchange as gather_partial_states in ONE collective and one host sync: every rank contributes a packed
    [len(cols) * max_rows + 1] int64 buffer (last word = its row count; floating-point states travel bit for bit); `dst`
    returns one concatenated tensor per column in the column's own dtype (rank 0's rows first), the others return None
    without waiting."""
    world, rank = dist.get_world_size(), dist.get_rank()
    g = int(cols[0].numel())
    if g > max_rows:
        raise ValueError(f"{g} partial rows exceed the gather capacity {max_rows}")
    dev = cols[0].device
    nc = len(cols)
    packed = torch.zeros(nc * max_rows + 1, dtype=torch.int64, device=dev)
    for k, c in enumerate(cols):
        packed[k * max_rows:k * max_rows + g] = _to_i64_bits(c)
    packed[-1] = g
    flat = torch.empty(world * (nc * max_rows + 1), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(flat, packed)
    if rank != dst:
        return None
    out = flat.view(world, nc * max_rows + 1)
    counts = [int(x) for x in out[:, -1].tolist()]
    return [_from_i64_bits(torch.cat([out[r, k * max_rows:k * max_rows + counts[r]] for r in range(world)]), cols[k].dtype) for k in range(nc)]


_REDUCE_OPS = {0: "SUM", 1: "MIN", 2: "MAX"}


def dense_state_views(arrays, device):
    """[(tensor aliasing the state array, reduce code)] for what sr_agg_dense_state returned"""
    out = []
    for ptr, count, elem_type, reduce in arrays:
        typestr = "<f8" if elem_type == 8 else "<i8"   # SR_TYPE_DOUBLE = 8, else int64 states
        out.append((torch.as_tensor(_DevicePtr(ptr, count, typestr), device=device), reduce))
    return out


def all_reduce_dense_state(arrays, device):
    """In-place merge of dense aggregate tables across ranks (all-reduce on the enumerable slot array).
    arrays: what sr_agg_dense_state returned on this rank -- [(device_ptr, count, elem_type, reduce)]; every rank
    passes the same list shape.  Afterwards every rank's table holds the final states.
    The library keeps the arrays of a dense table back to back in one slab, so neighbouring arrays with the same operator
    and element type are ONE contiguous range: they travel in one in-place collective, no cat / copy around it (Q4.1:
    COUNT(*) + two int64 sums = one all-reduce over 3 x 175 int64).  The ranks first agree on the list shape -- a mismatch
    (a column that turned nullable on one rank only) would otherwise hang the collective: MIN / MAX all-reduces of the
    list's description, once per distinct list."""
    runs = []   # [ptr, count, elem_type, reduce]
    for ptr, count, elem_type, reduce in arrays:
        if runs and runs[-1][2] == elem_type and runs[-1][3] == reduce and runs[-1][0] + 8 * runs[-1][1] == ptr:
            runs[-1][1] += count
        else:
            runs.append([ptr, count, elem_type, reduce])
    key = tuple([len(runs)] + [x for r in runs for x in (r[1], r[2], r[3])])
    if dist.get_world_size() > 1 and key not in _CHECKED_SHAPES:   # a property of the plan: checked the first time it is seen
        shape = torch.tensor(key, dtype=torch.int64, device=device)
        n = torch.tensor([len(key)], dtype=torch.int64, device=device)
        nlo, nhi = n.clone(), n.clone()
        dist.all_reduce(nlo, op=dist.ReduceOp.MIN)
        dist.all_reduce(nhi, op=dist.ReduceOp.MAX)
        ok = int(nlo) == int(nhi)
        if ok:
            lo, hi = shape.clone(), shape.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            ok = torch.equal(lo, hi)
        if not ok:
            raise RuntimeError("all_reduce_dense_state: the ranks expose different state array lists (a column nullable on some ranks only?)")
        _CHECKED_SHAPES.add(key)
    for ptr, count, elem_type, reduce in runs:
        typestr = "<f8" if elem_type == 8 else "<i8"
        t = torch.as_tensor(_DevicePtr(ptr, count, typestr), device=device)
        dist.all_reduce(t, op=getattr(dist.ReduceOp, _REDUCE_OPS[reduce]))


_CHECKED_SHAPES = set()


def all_reduce_runtime_filter(directory, min_value, max_value, num_inserted, has_null):
    """Global runtime filter (the RuntimeFilterMerger of be/src/runtime/runtime_filter_worker.cpp, every fragment instance
    ships its PARTIAL filter of a partitioned join's build side and the merged filter goes to the probe-side scans): the
    partial SimdBlockFilter directories -- all sized for the GLOBAL build row count, so they have the same shape -- are
    gathered in one collective and OR-ed (SimdBlockFilter::merge is a bitwise OR), min / max / has_null / counts are
    reduced alongside.  directory: 1-D int32 tensor aliasing or holding the 32-byte buckets (device tensor under NCCL,
    host tensor under gloo), modified in place.  Returns (min, max, num_inserted, has_null) of the merged filter; feed
    them with the directory to sr_rf_merge_directory of an EMPTY filter (or overwrite this rank's own)."""
    dev = directory.device
    if directory.numel():
        # NCCL has no bitwise-OR reduction: gather the partial directories (a filter is at most a few MB) and OR them here;
        # a BE would hand each gathered slice to sr_rf_merge_directory (k_or_u32) instead of the torch op
        world = dist.get_world_size()
        flat = torch.empty(world * directory.numel(), dtype=directory.dtype, device=dev)
        dist.all_gather_into_tensor(flat, directory.contiguous())
        parts = flat.view(world, directory.numel())
        for r in range(world):
            directory.bitwise_or_(parts[r])
    lo = torch.tensor([min_value], dtype=torch.int64, device=dev)
    hi = torch.tensor([max_value, 1 if has_null else 0], dtype=torch.int64, device=dev)
    cnt = torch.tensor([num_inserted], dtype=torch.int64, device=dev)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    return int(lo[0]), int(hi[0]), int(cnt[0]), bool(int(hi[1]))


def all_gather_in_values(values, row_limit=1024):
    """the IN part of a global runtime filter (PartialRuntimeFilterMerger: the total filter keeps its IN list only when every
    partial filter has one and the union stays within the row limit).  values: sorted int64 numpy array of this rank's IN
    part or None.  Returns the merged array (the same on every rank) or None."""
    import numpy as np
    world = dist.get_world_size()
    parts = [None] * world
    dist.all_gather_object(parts, None if values is None else np.asarray(values, dtype=np.int64).tolist())
    if any(p is None for p in parts):
        return None
    merged = np.unique(np.concatenate([np.asarray(p, dtype=np.int64) for p in parts])) if parts else np.zeros(0, dtype=np.int64)
    return None if len(merged) > row_limit else merged


def exchange_partitions(cols, channel_offsets):
    """HASH_PARTITIONED exchange.  cols: list of 1-D tensors already reordered so that the rows of channel c occupy
    [channel_offsets[c], channel_offsets[c+1]) (what sr_xchg_partition / the reference's counting sort produce);
    channel c is rank c.  Returns the list of received columns (rows from rank 0 first, then rank 1, ... -- each
    sender's rows keep their order, like the reference's per-sender queues).
    One small all-to-all carries the per-channel row counts (the receive sizes must be known to the host: one sync, the
    only one), then ALL columns travel in ONE grouped NCCL operation (ncclGroupStart .. ncclSend/ncclRecv per column and
    peer .. ncclGroupEnd via batch_isend_irecv); a rank's own channel is a device-to-device copy."""
    world, rank = dist.get_world_size(), dist.get_rank()
    if len(channel_offsets) != world + 1:
        raise ValueError("one channel per rank expected")
    dev = cols[0].device
    send_list = [int(channel_offsets[c + 1] - channel_offsets[c]) for c in range(world)]
    send_counts = torch.tensor(send_list, dtype=torch.int64, device=dev)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts)
    recv_list = [int(x) for x in recv_counts.tolist()]
    send_off = [int(x) for x in channel_offsets]
    recv_off = [0]
    for x in recv_list:
        recv_off.append(recv_off[-1] + x)
    out = [torch.empty(recv_off[-1], dtype=c.dtype, device=dev) for c in cols]
    ops = []
    for c, r in zip(cols, out):
        c = c.contiguous()
        for p in range(world):
            if p == rank:
                if send_list[p]:
                    r[recv_off[p]:recv_off[p + 1]].copy_(c[send_off[p]:send_off[p + 1]])
                continue
            if send_list[p]:
                ops.append(dist.P2POp(dist.isend, c[send_off[p]:send_off[p + 1]], p))
            if recv_list[p]:
                ops.append(dist.P2POp(dist.irecv, r[recv_off[p]:recv_off[p + 1]], p))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return out
