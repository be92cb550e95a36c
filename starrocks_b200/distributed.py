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


def gather_partial_states(cols, max_rows, dst=0):
    """cols: list of 1-D integer tensors (same length g <= max_rows) holding a rank's partial aggregate rows
    (group keys and states).  Returns on `dst` a list (one entry per rank) of lists of tensors trimmed to each
    rank's row count; None elsewhere."""
    world, rank = dist.get_world_size(), dist.get_rank()
    g = int(cols[0].numel())
    if g > max_rows:
        raise ValueError(f"{g} partial rows exceed the gather capacity {max_rows}")
    dev = cols[0].device
    part = torch.zeros((len(cols), max_rows), dtype=torch.int64, device=dev)
    for k, c in enumerate(cols):
        part[k, :g] = c.to(torch.int64)
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
        out.append([p[k, :gg].contiguous() for k in range(len(cols))])
    return out


def exchange_partitions(cols, channel_offsets):
    """HASH_PARTITIONED exchange.  cols: list of 1-D tensors already reordered so that the rows of channel c occupy
    [channel_offsets[c], channel_offsets[c+1]) (what sr_xchg_partition / the reference's counting sort produce);
    channel c is rank c.  Returns the list of received columns (rows from rank 0 first, then rank 1, ... -- each
    sender's rows keep their order, like the reference's per-sender queues)."""
    world = dist.get_world_size()
    if len(channel_offsets) != world + 1:
        raise ValueError("one channel per rank expected")
    dev = cols[0].device
    send_counts = torch.tensor([int(channel_offsets[c + 1] - channel_offsets[c]) for c in range(world)], dtype=torch.int64, device=dev)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts)
    send_list = [int(x) for x in send_counts.tolist()]
    recv_list = [int(x) for x in recv_counts.tolist()]
    out = []
    for c in cols:
        r = torch.empty(sum(recv_list), dtype=c.dtype, device=dev)
        dist.all_to_all_single(r, c.contiguous(), output_split_sizes=recv_list, input_split_sizes=send_list)
        out.append(r)
    return out
