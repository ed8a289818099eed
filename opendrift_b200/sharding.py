"""Multi-GPU decomposition of the advection path: particles are independent given the
forcing, so they are sharded by contiguous index ranges over the ranks (one process per GPU) and the forcing
slabs are replicated; the only exchange is the broadcast of each new reader time slab from the rank that read
it (NCCL over NVLink on the GPU box, gloo in the CPU tests) and a small all-reduce of run statistics.

Optional spatial-tile mode (BASELINE configs[2] wording: "field broadcast + particle all-to-all"; needed only when the
forcing is too large to replicate): the domain is cut into one longitude strip per rank (`strip_owner`), and after a step
the particles that left their strip travel to the new owner as packed SoA records in ONE all-to-all (`exchange_particles`:
counts first, then a single `all_to_all_single` of [n, record_bytes] rows).  Element identity travels in the ID column.
The field side: `strip_columns` gives every rank the columns of its strip plus a halo wide enough for the Runge-Kutta
excursions of one step, `scatter_field_tiles` / `receive_field_tile` move each new time slab's tiles from the rank that read it.
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


def _collective_device():
    """Where the tensors of a collective must live: the GPU for NCCL, the host for gloo."""
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_backend() == 'nccl':
        return torch.device('cuda', torch.cuda.current_device())
    return torch.device('cpu')


def gather_by_id(local_ids, local_values, n_total, dtype=np.float64):
    """Assemble a full-length array keyed by element ID from per-rank pieces (all ranks get the result)."""
    import torch
    import torch.distributed as dist
    out = torch.zeros(n_total, dtype=torch.float64)
    out[torch.as_tensor(np.asarray(local_ids, dtype=np.int64))] = torch.as_tensor(np.asarray(local_values, dtype=np.float64))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        out = out.to(_collective_device())
        dist.all_reduce(out)
        out = out.cpu()
    return out.numpy().astype(dtype)


def allreduce_stats(count_active, lon_min, lon_max, lat_min, lat_max):
    """Global element count and bounding box (one small all-reduce per call)."""
    import torch
    import torch.distributed as dist
    mx = torch.tensor([lon_max, lat_max, -lon_min, -lat_min], dtype=torch.float64)
    cnt = torch.tensor([float(count_active)], dtype=torch.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        mx, cnt = mx.to(_collective_device()), cnt.to(_collective_device())
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt)
        mx, cnt = mx.cpu(), cnt.cpu()
    return int(cnt.item()), -float(mx[2]), float(mx[0]), -float(mx[3]), float(mx[1])


def strip_bounds(lon_min, lon_max, world):
    """Edges of `world` equal-width longitude strips covering [lon_min, lon_max]."""
    return np.linspace(float(lon_min), float(lon_max), world + 1)


def strip_owner(lon, bounds):
    """Owner rank of each particle: the strip its longitude falls in (particles outside the domain go to the edge strips).
    lon: 1-D tensor (any device); returns int64 tensor."""
    import torch
    world = len(bounds) - 1
    inner = torch.as_tensor(np.asarray(bounds[1:-1], dtype=np.float64), device=lon.device)
    owner = torch.bucketize(lon.to(torch.float64), inner, right=True)
    return owner.clamp_(0, world - 1)


def _pack(columns, order):
    """SoA columns -> [n, record_bytes] uint8 rows in `order` (one record per particle)."""
    import torch
    parts, layout = [], []
    for name, t in columns.items():
        t = t[order].contiguous()
        width = t.element_size() * int(np.prod(t.shape[1:], dtype=np.int64))
        parts.append(t.view(torch.uint8).reshape(t.shape[0], width))
        layout.append((name, t.dtype, tuple(t.shape[1:]), width))
    rec = torch.cat(parts, dim=1) if parts else torch.empty((len(order), 0), dtype=torch.uint8)
    return rec.contiguous(), layout


def _unpack(rec, layout):
    out, off = {}, 0
    for name, dtype, tail, width in layout:
        col = rec[:, off:off + width].contiguous().view(dtype)
        out[name] = col.reshape((rec.shape[0],) + tail)
        off += width
    return out


def exchange_particles(columns, owner, group=None):
    """Send every particle to its owner rank; returns the columns of the particles this rank now owns (those it kept
    first, in their previous relative order, then the arrivals by source rank).

    columns: dict name -> tensor with leading dimension n (same device, any dtypes); owner: int64 tensor [n].
    Two collectives: the per-destination counts (world int64 each way) and ONE all_to_all_single of the packed records.
    With world size 1 (or no process group) the input is returned unchanged."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(columns)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    owner = owner.to(torch.int64)
    order = torch.argsort(owner, stable=True)
    send_counts = torch.bincount(owner, minlength=world).to(torch.int64)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    rec, layout = _pack(columns, order)
    n_recv = int(recv_counts.sum().item())
    out = torch.empty((n_recv, rec.shape[1]), dtype=torch.uint8, device=rec.device)
    dist.all_to_all_single(out, rec, output_split_sizes=recv_counts.tolist(), input_split_sizes=send_counts.tolist(), group=group)
    # own particles first (they did not travel), then the arrivals in rank order
    offs = np.concatenate([[0], np.cumsum(recv_counts.tolist())])
    mine = out[offs[rank]:offs[rank + 1]]
    others = [out[offs[r]:offs[r + 1]] for r in range(world) if r != rank]
    return _unpack(torch.cat([mine] + others, dim=0), layout)


def exchange_particles_device(eng, columns, bounds, group=None):
    """exchange_particles with the packing done by the library on the device: od_pack_by_owner groups the elements by the strip
    their 'lon' column falls in and packs them as records, ONE all_to_all_single (after the counts) moves them over NCCL,
    od_unpack_records restores the SoA columns.  columns: dict name -> 1-D device tensor incl. 'lon' (float64).  Returns the
    columns of the elements this rank owns now (segments in source-rank order; identity is in the ID column)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(columns)
    rec, counts, layout, _ = eng.pack_by_owner(columns['lon'], bounds, columns)
    send_counts = torch.tensor(counts, dtype=torch.int64, device=rec.device)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    recv = recv_counts.tolist()
    out = torch.empty((int(sum(recv)), rec.shape[1]), dtype=torch.uint8, device=rec.device)
    dist.all_to_all_single(out, rec, output_split_sizes=recv, input_split_sizes=counts, group=group)
    return eng.unpack_records(out, layout)


def strip_columns(lon, bounds, halo_cells):
    """Column range [i0, i1) of the global grid each rank holds in spatial-tile mode: the cells of its longitude strip plus
    `halo_cells` on either side -- the reference's own block rule, buffer = ceil(max_speed * dt / pixel size) + 2 cells
    around the particles (basereader/variables.py:616-617), so that every Runge-Kutta stage of a particle that starts the
    step inside the strip finds its corners in the tile."""
    lon = np.asarray(lon, dtype=np.float64)
    nx = len(lon)
    out = []
    for r in range(len(bounds) - 1):
        inside = np.where((lon >= bounds[r]) & (lon <= bounds[r + 1]))[0]
        lo = (inside[0] if len(inside) else int(np.searchsorted(lon, bounds[r]))) - 1 - int(halo_cells)
        hi = (inside[-1] if len(inside) else int(np.searchsorted(lon, bounds[r + 1]))) + 2 + int(halo_cells)
        out.append((max(0, lo), min(nx, hi)))
    return out


def scatter_field_tiles(slab, columns, src=0, group=None):
    """One time slab [..., ny, nx] on `src` -> every rank's tile slab[..., i0:i1] (point-to-point sends of the tiles; the
    other ranks pass slab=None and the leading shape / dtype of the slab as `like`).  Returns this rank's tile."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        i0, i1 = columns[0]
        return slab[..., i0:i1].contiguous()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if rank == src:
        mine = None
        for r in range(world):
            i0, i1 = columns[r]
            tile = slab[..., i0:i1].contiguous()
            if r == src:
                mine = tile
            else:
                dist.send(tile, r, group=group)
        return mine
    raise RuntimeError('receivers call receive_field_tile')


def receive_field_tile(shape_prefix, columns, dtype, device='cpu', src=0, group=None):
    """The receiving side of scatter_field_tiles: allocates [*shape_prefix, i1 - i0] and receives this rank's tile."""
    import torch
    import torch.distributed as dist
    i0, i1 = columns[dist.get_rank(group)]
    tile = torch.empty(tuple(shape_prefix) + (i1 - i0,), dtype=dtype, device=device)
    dist.recv(tile, src, group=group)
    return tile
