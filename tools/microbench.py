"""Component timings on the GPU box: od_geod_fwd, od_interp (one RK stage worth of sampling), od_update_positions,
step_kernel<RK4>, sort/permute -- CUDA events, 10 M particles.  Also the target of `ncu --set full -k regex:...`."""
import sys, os, json
from datetime import timedelta
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opendrift_b200 import synthetic as syn
from opendrift_b200.engine import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
eng = Engine(0)
g = syn.GridSpec()
times = syn.slab_times(3)
slabs = [tuple(torch.from_numpy(a).cuda() for a in syn.double_gyre_uv(g, (t - syn.T0).total_seconds())) for t in times]
grp = eng.add_group(g.lon, g.lat, g.z, 2, times, lambda ti, c: slabs[ti][c], (0.0, 0.0))
lon0, lat0, z0 = syn.particle_cloud(n, seed=5)
lon, lat, z = eng.to_device(lon0.astype(np.float64)), eng.to_device(lat0.astype(np.float64)), eng.to_device(z0)
perm = eng.sort_by_cell(grp, lon, lat, z)
lon, lat, z = eng.permute(perm, lon), eng.permute(perm, lat), eng.permute(perm, z)
az = torch.rand(n, device='cuda', dtype=torch.float64) * 360 - 180
dist = torch.rand(n, device='cuda', dtype=torch.float64) * 600
xv = torch.randn(n, device='cuda', dtype=torch.float32)
yv = torch.randn(n, device='cuda', dtype=torch.float32)
t = times[0] + timedelta(seconds=300)
dt = timedelta(seconds=600)


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        out.append(a.elapsed_time(b))
    return float(np.median(out))


res = {}
l2, a2 = lon.clone(), lat.clone()
res['geod_fwd_ms'] = timeit(lambda: eng.geod_fwd(l2, a2, az, dist))
res['interp_uv_ms'] = timeit(lambda: eng.interp(grp, t, lon, lat, z))
l2, a2 = lon.clone(), lat.clone()
res['update_positions_f32_ms'] = timeit(lambda: eng.update_positions(l2, a2, xv, yv, None, 600.0))
l2, a2 = lon.clone(), lat.clone()
res['step_rk4_exact_ms'] = timeit(lambda: eng.advect_current(grp, 'runge-kutta4', t, dt, l2, a2, z, fast=0))
l2, a2 = lon.clone(), lat.clone()
res['step_rk4_series_ms'] = timeit(lambda: eng.advect_current(grp, 'runge-kutta4', t, dt, l2, a2, z, fast=2))
l2, a2 = lon.clone(), lat.clone()
res['step_euler_series_ms'] = timeit(lambda: eng.advect_current(grp, 'euler', t, dt, l2, a2, z, fast=2))
l2, a2 = lon.clone(), lat.clone()
res['step_rk4_fast_ms'] = timeit(lambda: eng.advect_current(grp, 'runge-kutta4', t, dt, l2, a2, z, fast=True))
eng.set_tile(True)
l2, a2 = lon.clone(), lat.clone()
res['step_rk4_tile_ms'] = timeit(lambda: eng.advect_current(grp, 'runge-kutta4', t, dt, l2, a2, z))
l2, a2 = lon.clone(), lat.clone()
res['step_rk4_fast_tile_ms'] = timeit(lambda: eng.advect_current(grp, 'runge-kutta4', t, dt, l2, a2, z, fast=True))
eng.set_tile(False)
l2, a2 = lon.clone(), lat.clone()
res['step_euler_exact_ms'] = timeit(lambda: eng.advect_current(grp, 'euler', t, dt, l2, a2, z, fast=0))
# cfg 4: vertical mixing, 50-level diffusivity column, dt/dt_mix = 10 inner iterations, Philox draws
kslabs = [torch.from_numpy(syn.vertical_diffusivity(g, (tt - syn.T0).total_seconds())).cuda() for tt in times]
kgrp = eng.add_group(g.lon, g.lat, g.z, 1, times, lambda ti, c: kslabs[ti], (0.0,))
ids = torch.arange(n, device='cuda', dtype=torch.int32)
res['vertical_mixing_10it_ms'] = timeit(lambda: eng.vertical_mixing(kgrp, t, lon, lat, z, 60.0, 10, ids=ids, seed=1))
res['vertical_mixing_60it_ms'] = timeit(lambda: eng.vertical_mixing(kgrp, t, lon, lat, z, 60.0, 60, ids=ids, seed=1))
# cfg 5: Leeway step (2-D wind + current, Euler)
g2 = syn.GridSpec(nz=1)
wsl = [tuple(torch.from_numpy(a).cuda() for a in syn.wind_xy(g2, (tt - syn.T0).total_seconds())) for tt in times]
csl = [tuple(torch.from_numpy(a).cuda() for a in syn.double_gyre_uv(g2, (tt - syn.T0).total_seconds(), three_d=False)) for tt in times]
wg = eng.add_group(g2.lon, g2.lat, None, 2, times, lambda ti, c: wsl[ti][c], (float('nan'),) * 2)
cg = eng.add_group(g2.lon, g2.lat, None, 2, times, lambda ti, c: csl[ti][c], (float('nan'),) * 2)
el = {k: torch.rand(n, device='cuda', dtype=torch.float32) for k in ('dw_slope', 'dw_offset', 'dw_eps', 'cw_slope', 'cw_offset', 'cw_eps')}
el['orientation'] = (torch.arange(n, device='cuda') % 2).to(torch.uint8)
el['capsized'] = None
el['jibe_probability'] = torch.full((n,), 0.04, device='cuda', dtype=torch.float64)
l2, a2 = lon.clone(), lat.clone()
res['leeway_step_ms'] = timeit(lambda: eng.leeway_step(wg, cg, t, dt, l2, a2, el, ids=ids, seed=3))
res['sort_by_cell_ms'] = timeit(lambda: eng.sort_by_cell(grp, lon, lat, z))
res['permute_f64_ms'] = timeit(lambda: eng.permute(perm, lon))
res['n'] = n
print(json.dumps(res))
