"""Target of `ncu --set full -k regex:^step_kernel$ -s 2 -c 1`: the bench's dominant launch -- the fused OceanDrift step
(RK4 current advection + vertical advection, default arithmetic) on 10 M cell-sorted particles in the 512x512x50 u/v/w field."""
import os
import sys
from datetime import timedelta

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opendrift_b200 import synthetic as syn          # noqa: E402
from opendrift_b200.engine import Engine             # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 2
eng = Engine(0)
g = syn.GridSpec()
times = syn.slab_times(3)
slabs = [tuple(torch.from_numpy(a).cuda() for a in syn.double_gyre_uv(g, (t - syn.T0).total_seconds())) for t in times]
w = torch.from_numpy(syn.upward_w(g)).cuda()
grp = eng.add_group(g.lon, g.lat, g.z, 2, times, lambda ti, c: slabs[ti][c], (0.0, 0.0))
wgrp = eng.add_group(g.lon, g.lat, g.z, 1, times, lambda ti, c: w, (0.0,))
lon0, lat0, z0 = syn.particle_cloud(n, seed=5)
lon, lat, z = eng.to_device(lon0.astype(np.float64)), eng.to_device(lat0.astype(np.float64)), eng.to_device(z0)
perm = eng.sort_by_cell(grp, lon, lat, z)
lon, lat, z = eng.permute(perm, lon), eng.permute(perm, lat), eng.permute(perm, z)
t, dt = times[0] + timedelta(seconds=300), timedelta(seconds=600)
for _ in range(4):
    eng.step_oceandrift(grp, 'runge-kutta4', t, dt, lon.clone(), lat.clone(), z.clone(), w_group=wgrp, fast=mode)
torch.cuda.synchronize()
print('done')
