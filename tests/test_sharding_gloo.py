"""World-size-2 test (gloo, CPU) of the multi-GPU host logic: index sharding, slab broadcast, gather by ID,
statistics all-reduce; and that sharded advection (oracle arithmetic per shard) equals the unsharded run."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import common
from opendrift_b200 import sharding


def test_shard_ranges_partition():
    for n in (0, 1, 7, 10, 1000003):
        for w in (1, 2, 3, 8):
            r = [sharding.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sys.path.insert(0, common.ROOT)
    from oracle import advect_port as ap
    fx = common.Fixture('rk4_3d')
    # rank 0 "reads" the forcing; the other rank receives it by broadcast
    shape_u = fx.u.shape
    u = torch.from_numpy(fx.u.copy()) if rank == 0 else torch.zeros(shape_u, dtype=torch.float32)
    v = torch.from_numpy(fx.v.copy()) if rank == 0 else torch.zeros(shape_u, dtype=torch.float32)
    sharding.broadcast_slab(u, 0)
    sharding.broadcast_slab(v, 0)
    assert np.array_equal(u.numpy(), fx.u) and np.array_equal(v.numpy(), fx.v)
    lo, hi = sharding.shard_range(fx.n, rank, world)
    reader = ap.GridReader(fx.grid_lon, fx.grid_lat, fx.grid_z, fx.times, {common.CUR[0]: u.numpy(), common.CUR[1]: v.numpy()})
    lon, lat, z = ap.run_oceandrift([reader], fx.lon0[lo:hi], fx.lat0[lo:hi], fx.z0[lo:hi], fx.start, fx.dt, fx.steps,
                                    scheme=fx.meta['scheme'])
    full_lon = sharding.gather_by_id(np.arange(lo, hi), lon, fx.n)
    full_lat = sharding.gather_by_id(np.arange(lo, hi), lat, fx.n)
    cnt, lon_min, lon_max, lat_min, lat_max = sharding.allreduce_stats(hi - lo, lon.min(), lon.max(), lat.min(), lat.max())
    ok = (np.array_equal(full_lon, fx.lon) and np.array_equal(full_lat, fx.lat) and cnt == fx.n
          and lon_min == fx.lon.min() and lat_max == fx.lat.max())
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_rank_sharded_run_equals_reference():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)], res
