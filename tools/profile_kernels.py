"""Targets of `ncu --set full -k regex:<kernel> -s 1 -c 1 python tools/profile_kernels.py <which>`: the dominant launches of the
other BASELINE configurations at their scale -- mix (vertical mixing, 5 M elements, 10 inner iterations: configs[3]), leeway (Leeway
step, 20 M elements: configs[4]), analytic (the double gyre on its stereographic plane, RK4, 10 M elements: configs[0]), fast (the fused
OceanDrift step in the FAST arithmetic, 10 M cell-sorted elements)."""
import os
import sys
from datetime import datetime, timedelta

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opendrift_b200 import synthetic as syn          # noqa: E402
from opendrift_b200.engine import Engine             # noqa: E402

which = sys.argv[1]
eng = Engine(0)
dev = eng.device
dt = timedelta(seconds=600)
if which == 'mix':
    n = 5_000_000
    g = syn.GridSpec()
    times = syn.slab_times(3)
    K = [torch.from_numpy(syn.vertical_diffusivity(g, (t - syn.T0).total_seconds())).to(dev) for t in times]
    gk = eng.add_group(g.lon, g.lat, g.z, 1, times, lambda ti, c: K[ti], (0.0,))
    lon0, lat0, z0 = syn.particle_cloud(n, seed=9)
    lon, lat, z = eng.to_device(lon0.astype(np.float64)), eng.to_device(lat0.astype(np.float64)), eng.to_device(z0.astype(np.float64))
    ids = torch.arange(n, dtype=torch.int32, device=dev)
    mv = torch.ones(n, dtype=torch.int32, device=dev)
    for k in range(3):
        eng.vertical_mixing(gk, times[0] + timedelta(seconds=900), lon, lat, z, 60.0, 10, moving=mv, ids=ids, rand=None, seed=1, step_index=k)
elif which == 'leeway':
    n = 20_000_000
    g = syn.GridSpec(nz=1)
    times = syn.slab_times(3)
    cur = [tuple(torch.from_numpy(a).to(dev) for a in syn.double_gyre_uv(g, (t - syn.T0).total_seconds(), three_d=False)) for t in times]
    wnd = [tuple(torch.from_numpy(a).to(dev) for a in syn.wind_xy(g, (t - syn.T0).total_seconds())) for t in times]
    nan = float('nan')
    gc = eng.add_group(g.lon, g.lat, None, 2, times, lambda ti, c: cur[ti][c], (nan, nan))
    gw = eng.add_group(g.lon, g.lat, None, 2, times, lambda ti, c: wnd[ti][c], (nan, nan))
    lon0, lat0, _ = syn.particle_cloud(n, seed=9, three_d=False)
    lon, lat = eng.to_device(lon0.astype(np.float64)), eng.to_device(lat0.astype(np.float64))
    f32 = lambda v: torch.full((n,), v, dtype=torch.float32, device=dev)     # noqa: E731
    el = {'dw_slope': f32(0.96), 'dw_offset': f32(0.0), 'dw_eps': f32(1.0), 'cw_slope': f32(0.54), 'cw_offset': f32(0.0), 'cw_eps': f32(0.5),
          'orientation': (torch.arange(n, device=dev) % 2).to(torch.uint8), 'capsized': torch.zeros(n, dtype=torch.uint8, device=dev),
          'jibe_probability': torch.full((n,), 0.04, dtype=torch.float64, device=dev)}
    mv = torch.ones(n, dtype=torch.int32, device=dev)
    st = torch.zeros(n, dtype=torch.int32, device=dev)
    ids = torch.arange(n, dtype=torch.int32, device=dev)
    for k in range(3):
        eng.leeway_step(gw, gc, times[0] + timedelta(seconds=900), dt, lon, lat, el, moving=mv, status=st, ids=ids, rand=None, seed=1,
                        step_index=k, missing_code=1)
elif which == 'analytic':
    from opendrift_b200.readers import reader_double_gyre
    n = 10_000_000
    rd = reader_double_gyre.Reader(initial_time=datetime(2000, 1, 1), epsilon=0.25, omega=0.628, A=0.25)
    rd.bind(eng, {v: 0.0 for v in rd.variables})
    d = rd.analytic_desc()
    rng = np.random.default_rng(0)
    lon, lat = rd.xy2lonlat(rng.uniform(0.0, 2.0, n), rng.uniform(0.0, 1.0, n))
    dl, da = eng.to_device(lon), eng.to_device(lat)
    for k in range(3):
        eng.analytic_advect(d, 'runge-kutta4', (0.0, 0.05, 0.1), 0.1, dl.clone(), da.clone())
else:
    raise SystemExit('mix | leeway | analytic')
torch.cuda.synchronize()
print('done')
