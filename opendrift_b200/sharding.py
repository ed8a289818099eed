"""Multi-GPU decomposition of the advection path: particles are independent given the
forcing, so they are sharded by contiguous index ranges over the ranks (one process per GPU) and the forcing
slabs are replicated; the only exchange is the broadcast of each new reader time slab from the rank that read
it (NCCL over NVLink on the GPU box, gloo in the CPU tests) and a small all-reduce of run statistics.
"""
import numpy as np


def shard_range(n_total, rank, world):
    """Contiguous index range [lo, hi) of `rank`: sizes differ by at most one, order preserved."""
    base, extra = divmod(n_total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def broadcast_slab(tensor, src=0):
    """Broadcast one forcing slab (in place) from the rank that read it."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(tensor, src)
    return tensor


def gather_by_id(local_ids, local_values, n_total, dtype=np.float64):
    """Assemble a full-length array keyed by element ID from per-rank pieces (all ranks get the result)."""
    import torch
    import torch.distributed as dist
    out = torch.zeros(n_total, dtype=torch.float64)
    out[torch.as_tensor(np.asarray(local_ids, dtype=np.int64))] = torch.as_tensor(np.asarray(local_values, dtype=np.float64))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(out)
    return out.numpy().astype(dtype)


def allreduce_stats(count_active, lon_min, lon_max, lat_min, lat_max):
    """Global element count and bounding box (one small all-reduce per call)."""
    import torch
    import torch.distributed as dist
    mx = torch.tensor([lon_max, lat_max, -lon_min, -lat_min], dtype=torch.float64)
    cnt = torch.tensor([float(count_active)], dtype=torch.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt)
    return int(cnt.item()), -float(mx[2]), float(mx[0]), -float(mx[3]), float(mx[1])
