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


def _tiled_field_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sys.path.insert(0, common.ROOT)
    from datetime import timedelta
    from oracle import advect_port as ap
    fx = common.Fixture('rk4_2d')
    bounds = sharding.strip_bounds(fx.grid_lon.min(), fx.grid_lon.max(), world)
    # halo: the reference's block buffer, ceil(max_speed * dt / pixel) + 2 cells (variables.py:616-617)
    dx_m = float(fx.grid_lon[1] - fx.grid_lon[0]) * 111e3 * np.cos(np.radians(float(fx.grid_lat.max())))
    vmax = float(max(np.nanmax(np.abs(fx.u)), np.nanmax(np.abs(fx.v))))
    halo = int(np.ceil(vmax * abs(fx.dt) / dx_m)) + 2
    cols = sharding.strip_columns(fx.grid_lon, bounds, halo)
    # rank 0 read the forcing; every rank receives only its tile of every slab
    tiles = {}
    for name, full in ((common.CUR[0], fx.u), (common.CUR[1], fx.v)):
        if rank == 0:
            tiles[name] = sharding.scatter_field_tiles(torch.from_numpy(full.copy()), cols, 0).numpy()
        else:
            tiles[name] = sharding.receive_field_tile(full.shape[:-1], cols, torch.float32, src=0).numpy()
    i0, i1 = cols[rank]
    ok = np.array_equal(tiles[common.CUR[0]], fx.u[..., i0:i1]) and (i1 - i0) < len(fx.grid_lon)       # a real sub-grid
    reader = ap.GridReader(fx.grid_lon[i0:i1], fx.grid_lat, None, fx.times, tiles)
    lo, hi = sharding.shard_range(fx.n, rank, world)
    cols_p = {'ID': torch.arange(lo, hi, dtype=torch.int32),
              'lon': torch.from_numpy(fx.lon0[lo:hi].astype(np.float32).astype(np.float64)),
              'lat': torch.from_numpy(fx.lat0[lo:hi].astype(np.float32).astype(np.float64)),
              'first': torch.ones(hi - lo, dtype=torch.uint8)}
    t = fx.start
    for k in range(fx.steps):
        cols_p = sharding.exchange_particles(cols_p, sharding.strip_owner(cols_p['lon'], bounds))
        n = len(cols_p['ID'])
        if n:
            lon, lat = cols_p['lon'].numpy(), cols_p['lat'].numpy()
            first = cols_p['first'].numpy().astype(bool)
            if first.all():
                lon, lat = lon.astype(np.float32), lat.astype(np.float32)
            z = np.zeros(n, dtype=np.float32)
            env = ap.get_environment([reader], common.CUR, t, lon, lat, z, fallback={})        # no fallback: a miss must show
            ok &= bool(np.isfinite(env[common.CUR[0]]).all())
            lon, lat = ap.advect_ocean_current([reader], fx.meta['scheme'], t, fx.dt, lon, lat, z, np.ones(n),
                                               np.ones(n, dtype=np.int32), env)
            cols_p['lon'], cols_p['lat'] = torch.from_numpy(np.asarray(lon, dtype=np.float64)), torch.from_numpy(np.asarray(lat, dtype=np.float64))
            cols_p['first'] = torch.zeros(n, dtype=torch.uint8)
        t = t + timedelta(seconds=fx.dt)
    ids = cols_p['ID'].numpy().astype(np.int64)
    full_lon = sharding.gather_by_id(ids, cols_p['lon'].numpy(), fx.n)
    full_lat = sharding.gather_by_id(ids, cols_p['lat'].numpy(), fx.n)
    # the tiles have their own float32 end points, like the reference's own sub-blocks: index arithmetic differs in the last bits
    e = common.max_err_deg(full_lon, full_lat, fx.lon, fx.lat)
    ok &= max(e) < 5e-8
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_rank_field_tiles_with_halo_and_particle_exchange():
    """BASELINE configs[2] wording end to end on gloo: every rank holds only its longitude strip of the field (+ halo), new
    slabs arrive as tiles from the rank that read them, particles move to their owner in one all-to-all per step; the result is
    the unsharded reference run's."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_tiled_field_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)], res


def test_strip_columns_cover_their_strips():
    lon = np.linspace(2.0, 3.95, 40)
    b = sharding.strip_bounds(lon.min(), lon.max(), 4)
    cols = sharding.strip_columns(lon, b, 2)
    for r, (i0, i1) in enumerate(cols):
        assert 0 <= i0 < i1 <= 40
        inside = np.where((lon >= b[r]) & (lon <= b[r + 1]))[0]
        assert i0 <= max(0, inside[0] - 3) and i1 >= min(40, inside[-1] + 4)
    assert cols[0][0] == 0 and cols[-1][1] == 40
    assert sharding.scatter_field_tiles(torch.arange(12.0).reshape(3, 4), [(1, 3)]).tolist() == [[1.0, 2.0], [5.0, 6.0], [9.0, 10.0]]


# ---- OceanDrift.run() under a torch.distributed job: index shards, slabs read by rank 0 and broadcast into the ring -------------
def _run_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sys.path.insert(0, common.ROOT)
    sys.path.insert(0, os.path.join(common.ROOT, 'tests'))
    from hostengine import HostEngine
    import test_gpu_dropin as T
    import opendrift_b200.engine as E
    import opendrift_b200.models.basemodel as B
    eng = HostEngine()
    E.default_engine = B.default_engine = lambda device=None: eng
    fx = common.Fixture('rk4_3d_full')
    fx.meta['diffusivity'] = 0.0            # (the legacy generator's draws are per process: parity mode is single-process)
    reads = {'n': 0}
    o = T._model(fx, **{'drift:max_age_seconds': 4000})
    if rank != 0:                            # only rank 0 may touch the data: the other ranks' suppliers must never be called
        for r in o.env.readers.values():
            orig = r.get_variables

            def guarded(requested_variables, time=None, x=None, y=None, z=None, orig=orig, r=r):
                if time != r.times[0]:      # (the geometry probe at bind() reads the first block)
                    reads['n'] += 1
                return orig(requested_variables, time=time, x=x, y=y, z=z)
            r.get_variables = guarded
    o.run(steps=fx.steps, time_step=fx.dt, time_step_output=2 * fx.dt)
    lo, hi, n_all = o.shard
    ids = np.concatenate([np.asarray(o.elements.ID, dtype=np.int64), np.asarray(o.elements_deactivated.ID, dtype=np.int64)])
    lon = np.concatenate([np.asarray(o.elements.lon), np.asarray(o.elements_deactivated.lon)])
    lat = np.concatenate([np.asarray(o.elements.lat), np.asarray(o.elements_deactivated.lat)])
    full_lon = sharding.gather_by_id(ids, lon, n_all)
    full_lat = sharding.gather_by_id(ids, lat, n_all)
    hist_rows = o.history.lon.values.shape
    q.put((rank, (lo, hi, n_all), full_lon, full_lat, reads['n'], eng.dist.slabs_broadcast, hist_rows, sorted(ids.tolist()) == list(range(lo, hi))))
    dist.destroy_process_group()


def test_two_rank_model_run_shards_elements_and_broadcasts_slabs():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_run_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
    # the unsharded run of the same model in this process
    from hostengine import HostEngine
    import test_gpu_dropin as T
    fx = common.Fixture('rk4_3d_full')
    fx.meta['diffusivity'] = 0.0
    o = T._model(fx, **{'drift:max_age_seconds': 4000})
    o._engine = HostEngine()
    o.run(steps=fx.steps, time_step=fx.dt, time_step_output=2 * fx.dt)
    ids = np.concatenate([np.asarray(o.elements.ID, dtype=np.int64), np.asarray(o.elements_deactivated.ID, dtype=np.int64)])
    ref_lon, ref_lat = np.zeros(fx.n), np.zeros(fx.n)
    ref_lon[ids] = np.concatenate([np.asarray(o.elements.lon), np.asarray(o.elements_deactivated.lon)])
    ref_lat[ids] = np.concatenate([np.asarray(o.elements.lat), np.asarray(o.elements_deactivated.lat)])
    n_cols = len(o.history['time'])
    for rank, (lo, hi, n_all), lon, lat, foreign_reads, n_bcast, hist_shape, ids_ok in res:
        assert n_all == fx.n and (lo, hi) == sharding.shard_range(fx.n, rank, 2) and ids_ok
        assert np.array_equal(lon, ref_lon) and np.array_equal(lat, ref_lat)          # bit-identical to the unsharded run
        # every rank buffers its own trajectories only; a distributed run keeps stepping (idle) to the requested end so that the
        # ranks stay in lockstep, where the single process stops when its last element has retired
        assert hist_shape[0] == hi - lo and hist_shape[1] >= n_cols
        assert n_bcast > 0
        if rank != 0:
            assert foreign_reads == 0                                                  # forcing arrived by broadcast only


def _coast_worker(rank, world, port, q, case):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sys.path.insert(0, common.ROOT)
    sys.path.insert(0, os.path.join(common.ROOT, 'tests'))
    from hostengine import HostEngine
    import coastcases as cc
    import opendrift_b200.engine as E
    import opendrift_b200.models.basemodel as B
    eng = HostEngine()
    E.default_engine = B.default_engine = lambda device=None: eng
    o = cc.run_product(case)
    s = cc.summary(o)
    q.put((rank, o.shard, s['id'], s['lon'], s['lat'], s['d_id'], s['d_lon'], s['d_lat'], s['d_status'], [str(c) for c in s['cats']]))
    dist.destroy_process_group()


def test_two_rank_run_with_coastline_interaction_equals_the_reference():
    """Index shards under gloo with a land mask reader (read by rank 0, broadcast), 'stranding' and 'previous': the union of the
    two ranks' elements is the reference's result -- survivors, deactivated elements and their status."""
    import coastcases as cc
    for case in ('stranding_rk4_3d_release', 'previous_rk4_3d'):
        ctx = mp.get_context('spawn')
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_coast_worker, args=(r, 2, port, q, case)) for r in range(2)]
        for p in procs:
            p.start()
        res = sorted([q.get(timeout=600) for _ in procs], key=lambda r: r[0])
        for p in procs:
            p.join(timeout=60)
        ref = np.load(cc.GOLDEN)
        g = lambda k: ref['%s__%s' % (case, k)]                      # noqa: E731
        ids = np.concatenate([r[2] for r in res])
        lon, lat = np.concatenate([r[3] for r in res]), np.concatenate([r[4] for r in res])
        o = np.argsort(ids)
        ro = np.argsort(g('id'))
        assert np.array_equal(ids[o], g('id')[ro])
        assert max(common.max_err_deg(lon[o], lat[o], g('lon')[ro], g('lat')[ro])) < 5e-7
        d_ids = np.concatenate([r[5] for r in res])
        do, dro = np.argsort(d_ids), np.argsort(g('d_id'))
        assert np.array_equal(d_ids[do], g('d_id')[dro]) and len(d_ids) > 0
        names = np.concatenate([np.array(r[9])[r[8]] for r in res])          # status NAMES (every rank numbers its own categories)
        assert list(names[do]) == list(np.array([str(c) for c in g('cats')])[g('d_status')][dro])
        assert all(r[1][:2] == sharding.shard_range(cc.N, r[0], 2) for r in res)
