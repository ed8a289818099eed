"""Multi-GPU runs over NCCL (torchrun --nproc-per-node N tools/nccl_tile_run.py), one JSON line per leg from rank 0:

  shard   OceanDrift.run() of a reference fixture under the distributed job: index shards, forcing read by rank 0 and broadcast
          into the other ranks' device ring -- against the unmodified reference's result (tests/golden)
  tiles   BASELINE configs[2] wording: every rank holds only its longitude strip of the field (+ halo) as a window of its field
          group, elements travel to their owner in ONE all-to-all per step packed by od_pack_by_owner -- against the same fixture
  scale   the same exchange at 10 M elements per rank on the 512 x 512 x 50 field: time of pack + all-to-all + unpack per step

Only used on the GPU box (tools/gpu_multi.sh); the gloo twins of the first two legs run in the CPU suite (tests/test_sharding_gloo.py)."""
import json
import os
import sys
import time
from datetime import timedelta

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import common                                       # noqa: E402
from opendrift_b200 import sharding, synthetic as syn  # noqa: E402
from opendrift_b200.engine import Engine, bind_process_to_gpu_numa  # noqa: E402


def leg_shard(eng, rank, world):
    import test_gpu_dropin as T
    import opendrift_b200.engine as E
    import opendrift_b200.models.basemodel as B
    E.default_engine = B.default_engine = lambda device=None: eng
    fx = common.Fixture('rk4_3d_full')
    fx.meta['diffusivity'] = 0.0
    ref = common.run_port(fx)                       # (the fixture itself was generated with diffusion: the port gives the no-diffusion run)
    o = T._model(fx)
    o.run(steps=fx.steps, time_step=fx.dt, time_step_output=fx.dt)
    lo, hi, n_all = o.shard
    ids = np.asarray(o.elements.ID, dtype=np.int64)
    lon = sharding.gather_by_id(ids, np.asarray(o.elements.lon), n_all)
    lat = sharding.gather_by_id(ids, np.asarray(o.elements.lat), n_all)
    e = max(common.max_err_deg(lon, lat, ref[0], ref[1]))
    return {'leg': 'shard', 'world': world, 'elements': int(n_all), 'per_rank': int(hi - lo), 'steps': fx.steps, 'max_err_deg_vs_oracle': float(e),
            'slabs_broadcast': eng.dist.slabs_broadcast, 'ok': bool(e < 5e-8)}


def leg_tiles(eng, rank, world):
    fx = common.Fixture('rk4_2d')
    dev = eng.device
    bounds = sharding.strip_bounds(fx.grid_lon.min(), fx.grid_lon.max(), world)
    dx_m = float(fx.grid_lon[1] - fx.grid_lon[0]) * 111e3 * np.cos(np.radians(float(fx.grid_lat.max())))
    vmax = float(max(np.nanmax(np.abs(fx.u)), np.nanmax(np.abs(fx.v))))
    halo = int(np.ceil(vmax * abs(fx.dt) / dx_m)) + 2
    cols = sharding.strip_columns(fx.grid_lon, bounds, halo)
    i0, i1 = cols[rank]
    # every rank receives only its tile of every slab (NCCL point-to-point from the rank that read the data)
    tiles = {}
    for name, full in ((common.CUR[0], fx.u), (common.CUR[1], fx.v)):
        if rank == 0:
            tiles[name] = sharding.scatter_field_tiles(torch.from_numpy(full.copy()).to(dev), cols, 0)
        else:
            tiles[name] = sharding.receive_field_tile(full.shape[:-1], cols, torch.float32, device=dev, src=0)
    saved, eng.dist = eng.dist, None                # the tiles are this rank's own data: no slab broadcast in this mode
    g = eng.add_group(fx.grid_lon, fx.grid_lat, None, 2, fx.times, lambda ti, c: (tiles[common.CUR[0]], tiles[common.CUR[1]])[c][ti].contiguous(),
                      (0.0, 0.0))
    g.set_window(fx.grid_lon[i0:i1], fx.grid_lat)   # the group's blocks are the tile: index geometry of the tile's own axes
    lo, hi = sharding.shard_range(fx.n, rank, world)
    c = {'ID': torch.arange(lo, hi, dtype=torch.int32, device=dev),
         'lon': torch.from_numpy(fx.lon0[lo:hi].astype(np.float32).astype(np.float64)).to(dev),
         'lat': torch.from_numpy(fx.lat0[lo:hi].astype(np.float32).astype(np.float64)).to(dev)}
    t, dt = fx.start, timedelta(seconds=fx.dt)
    moved = 0
    for k in range(fx.steps):
        before = c['ID'].numel()
        owner = sharding.strip_owner(c['lon'], bounds)
        moved += int((owner != rank).sum())
        c = sharding.exchange_particles_device(eng, c, bounds)
        assert bool((sharding.strip_owner(c['lon'], bounds) == rank).all())
        if c['ID'].numel():
            eng.advect_current(g, fx.meta['scheme'], t, dt, c['lon'], c['lat'], None, pos_f32=(k == 0))
        t = t + dt
    eng.dist = saved
    ids = c['ID'].cpu().numpy().astype(np.int64)
    lon = sharding.gather_by_id(ids, c['lon'].cpu().numpy(), fx.n)
    lat = sharding.gather_by_id(ids, c['lat'].cpu().numpy(), fx.n)
    cnt = torch.tensor([float(len(ids)), float(moved)], dtype=torch.float64, device=dev)
    dist.all_reduce(cnt)
    e = max(common.max_err_deg(lon, lat, fx.lon, fx.lat))
    return {'leg': 'tiles', 'world': world, 'elements': fx.n, 'tile_columns': [int(i0), int(i1)], 'grid_columns': len(fx.grid_lon), 'halo_cells': halo,
            'elements_after': int(cnt[0]), 'crossings': int(cnt[1]), 'max_err_deg_vs_reference': float(e),
            'ok': bool(e < 5e-8 and int(cnt[0]) == fx.n and int(cnt[1]) > 0)}


def leg_scale(eng, rank, world, n=10_000_000, steps=12):
    dev = eng.device
    grid = syn.GridSpec()
    bounds = sharding.strip_bounds(float(grid.lon.min()), float(grid.lon.max()), world)
    halo = 4
    cols = sharding.strip_columns(grid.lon, bounds, halo)
    i0, i1 = cols[rank]
    times = syn.slab_times(4)
    slabs = [tuple(torch.from_numpy(np.ascontiguousarray(a[..., i0:i1])).to(dev) for a in syn.double_gyre_uv(grid, (tt - syn.T0).total_seconds()))
             for tt in times]
    saved, eng.dist = eng.dist, None
    g = eng.add_group(grid.lon, grid.lat, grid.z, 2, times, lambda ti, c: slabs[ti][c], (0.0, 0.0))
    g.set_window(grid.lon[i0:i1], grid.lat)
    rng = np.random.default_rng(100 + rank)
    lon = rng.uniform(bounds[rank] + 0.01, bounds[rank + 1] - 0.01, n)
    c = {'ID': torch.arange(rank * n, (rank + 1) * n, dtype=torch.int32, device=dev),
         'lon': torch.from_numpy(lon).to(dev), 'lat': torch.from_numpy(rng.uniform(55.5, 59.6, n)).to(dev),
         'z': torch.from_numpy(rng.uniform(-90, 0, n).astype(np.float32)).to(dev)}
    t, dt = times[0], timedelta(seconds=600)
    ex_ms, st_ms, crossed = [], [], 0
    for k in range(steps):
        a, b, d = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        n_before = c['ID'].numel()
        a.record()
        c = sharding.exchange_particles_device(eng, c, bounds)
        b.record()
        eng.advect_current(g, 'runge-kutta4', t, dt, c['lon'], c['lat'], c['z'])
        d.record()
        torch.cuda.synchronize()
        if k >= 2:
            ex_ms.append(a.elapsed_time(b))
            st_ms.append(b.elapsed_time(d))
        t = t + dt
    eng.dist = saved
    tt = torch.tensor([float(np.mean(ex_ms)), float(np.mean(st_ms)), float(c['ID'].numel())], dtype=torch.float64, device=dev)
    mx = tt.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(tt)
    rec_bytes = 4 + 8 + 8 + 4
    return {'leg': 'scale', 'world': world, 'elements_per_rank': n, 'record_bytes': rec_bytes, 'exchange_ms_max_over_ranks': float(mx[0]),
            'step_kernel_ms_max_over_ranks': float(mx[1]), 'elements_total_after': int(tt[2]), 'pack_GBps': n * rec_bytes * 2 / (float(mx[0]) * 1e-3) / 1e9,
            'note': 'exchange = od_pack_by_owner (owner search, stable grouping, record packing) + counts all-to-all + ONE all_to_all_single over NCCL + '
                    'od_unpack_records, every step, all elements'}


def main():
    local = int(os.environ.get('LOCAL_RANK', '0'))
    bind_process_to_gpu_numa(local)
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank, world = dist.get_rank(), dist.get_world_size()
    eng = Engine(local)
    eng.enable_distributed()
    for leg in (leg_shard, leg_tiles, leg_scale):
        try:
            r = leg(eng, rank, world)
        except Exception as ex:
            import traceback
            r = {'leg': leg.__name__, 'error': repr(ex)[:300], 'where': traceback.format_exc()[-600:]}
        if rank == 0:
            print(json.dumps(r), flush=True)
        dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
