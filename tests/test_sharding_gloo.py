"""World-size-2 tests (gloo, CPU) of the multi-GPU host logic: index sharding, slab broadcast, gather by ID,
statistics all-reduce; that sharded advection (oracle arithmetic per shard) equals the unsharded run; and the optional
spatial-tile mode: longitude strips + one all-to-all of packed particle records after every step."""
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


def _tile_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sys.path.insert(0, common.ROOT)
    from datetime import timedelta
    from oracle import advect_port as ap
    fx = common.Fixture('rk4_2d')
    reader = ap.GridReader(fx.grid_lon, fx.grid_lat, fx.grid_z, fx.times, {common.CUR[0]: fx.u, common.CUR[1]: fx.v})
    lo, hi = sharding.shard_range(fx.n, rank, world)
    bounds = sharding.strip_bounds(fx.grid_lon.min(), fx.grid_lon.max(), world)
    # seeding casts to float32 (elements.py:156-158); afterwards positions are float64
    cols = {'ID': torch.arange(lo, hi, dtype=torch.int32),
            'lon': torch.from_numpy(fx.lon0[lo:hi].astype(np.float32).astype(np.float64)),
            'lat': torch.from_numpy(fx.lat0[lo:hi].astype(np.float32).astype(np.float64)),
            'z': torch.from_numpy(np.asarray(fx.z0[lo:hi], dtype=np.float32)),
            'first': torch.ones(hi - lo, dtype=torch.uint8)}
    ok = True
    moved_total = 0
    t = fx.start
    for k in range(fx.steps):
        owner = sharding.strip_owner(cols['lon'], bounds)
        moved_total += int((owner != rank).sum())
        cols = sharding.exchange_particles(cols, owner)
        own = sharding.strip_owner(cols['lon'], bounds)
        ok &= bool((own == rank).all())                              # every particle sits with its owner
        n = len(cols['ID'])
        lon, lat = cols['lon'].numpy(), cols['lat'].numpy()
        if n:
            # one step of the oracle arithmetic on this rank's particles (float32 positions on the very first step only)
            first = cols['first'].numpy().astype(bool)
            assert first.all() or not first.any()
            if first.all():
                lon, lat = lon.astype(np.float32), lat.astype(np.float32)
            env = ap.get_environment([reader], common.CUR, t, lon, lat, cols['z'].numpy())
            lon, lat = ap.advect_ocean_current([reader], fx.meta['scheme'], t, fx.dt, lon, lat, cols['z'].numpy(),
                                               np.ones(n), np.ones(n, dtype=np.int32), env)
            cols['lon'], cols['lat'] = torch.from_numpy(np.asarray(lon, dtype=np.float64)), torch.from_numpy(np.asarray(lat, dtype=np.float64))
            cols['first'] = torch.zeros(n, dtype=torch.uint8)
        t = t + timedelta(seconds=fx.dt)
    ids = cols['ID'].numpy().astype(np.int64)
    full_lon = sharding.gather_by_id(ids, cols['lon'].numpy(), fx.n)
    full_lat = sharding.gather_by_id(ids, cols['lat'].numpy(), fx.n)
    cnt = torch.tensor([float(len(ids)), float(moved_total)], dtype=torch.float64)
    dist.all_reduce(cnt)
    ok &= int(cnt[0]) == fx.n and int(cnt[1]) > 0                    # nobody lost or duplicated; some did cross a strip edge
    ok &= bool(np.array_equal(full_lon, fx.lon) and np.array_equal(full_lat, fx.lat))   # == the unsharded reference run
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_rank_spatial_tiles_with_particle_all_to_all():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_tile_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)], res


def test_exchange_is_identity_without_process_group():
    cols = {'ID': torch.arange(7, dtype=torch.int32), 'lon': torch.rand(7, dtype=torch.float64), 'k': torch.rand(7, 3)}
    out = sharding.exchange_particles(cols, torch.zeros(7, dtype=torch.int64))
    assert all(torch.equal(out[k], cols[k]) for k in cols)
    rec, layout = sharding._pack(cols, torch.arange(6, -1, -1))
    assert rec.shape == (7, 4 + 8 + 12)
    back = sharding._unpack(rec, layout)
    assert all(torch.equal(back[k], cols[k].flip(0)) for k in cols)
    b = sharding.strip_bounds(0.0, 8.0, 4)
    assert sharding.strip_owner(torch.tensor([-3.0, 0.0, 1.99, 2.0, 7.9, 8.0, 11.0]), b).tolist() == [0, 0, 0, 1, 3, 3, 3]
