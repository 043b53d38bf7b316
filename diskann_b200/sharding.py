"""Multi-GPU plumbing for the search path: replicate the (immutable) index on every GPU, shard the
query batch into contiguous ranges, no collective on the search path.

Mirrors the reference's only parallelism on this path: `PartitionIter` over the queries with one
task per partition (diskann-benchmark-core/src/search/api.rs:400-434).  `torch.distributed` is
used as plumbing only (NCCL on GPUs, gloo in the CPU tests); compute stays in libdiskann_b200.so.
"""
import numpy as np


def partition(n, parts):
    """Contiguous ranges covering [0, n): the first n % parts ranges are one longer
    (PartitionIter, diskann-benchmark-core/src/search/api.rs:410-419)."""
    if parts <= 0:
        raise ValueError("parts must be positive")
    base, extra = divmod(n, parts)
    out, lo = [], 0
    for r in range(parts):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def shard_of(n, rank, world):
    return partition(n, world)[rank]


def broadcast_arrays(arrays, src=0, device=None):
    """Broadcast a dict of numpy arrays from `src` to every rank (one collective per buffer at
    index load).  Shapes/dtypes must be known on every rank (pass zero-filled arrays elsewhere).
    With `device` set the transfer goes through device tensors (NCCL over NVLink); returns the
    dict of torch tensors (on `device`, or CPU)."""
    import torch
    import torch.distributed as dist
    out = {}
    for name in sorted(arrays):
        a = arrays[name]
        view = a.view(np.int32) if a.dtype == np.uint32 else a  # torch has no uint32 collectives
        t = torch.from_numpy(np.ascontiguousarray(view))
        if device is not None:
            t = t.to(device)
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.broadcast(t, src=src)
        out[name] = t
    return out


def gather_results(local_ids, local_dists, n_total, k, dst=0):
    """Collect per-rank result rows (contiguous shards in rank order) on `dst`."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_ids, local_dists
    world, rank = dist.get_world_size(), dist.get_rank()
    ranges = partition(n_total, world)
    ids = [None] * world
    dists = [None] * world
    dist.all_gather_object(ids, np.ascontiguousarray(local_ids))
    dist.all_gather_object(dists, np.ascontiguousarray(local_dists))
    if rank != dst:
        return None, None
    for r, (lo, hi) in enumerate(ranges):
        assert ids[r].shape == (hi - lo, k), (r, ids[r].shape, (hi - lo, k))
    return np.concatenate(ids), np.concatenate(dists)


def max_over_ranks(value, device=None):
    """Device-side timing is reported as the max over ranks."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
