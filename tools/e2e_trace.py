"""Per-chunk timeline of od_advect_current_host (OD_HOST_TRACE=1) at the bench workload; prints to stderr."""
import os, sys
os.environ['OD_HOST_TRACE'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from datetime import timedelta
import numpy as np
import torch
from opendrift_b200 import synthetic as syn
from opendrift_b200.engine import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 0
eng = Engine(0)
g = syn.GridSpec()
times = syn.slab_times(3)
slabs = [tuple(torch.from_numpy(a).cuda() for a in syn.double_gyre_uv(g, (t - syn.T0).total_seconds())) for t in times]
grp = eng.add_group(g.lon, g.lat, g.z, 2, times, lambda ti, c: slabs[ti][c], (0.0, 0.0))
lon0, lat0, z0 = syn.particle_cloud(n, seed=5)
h_lon = torch.from_numpy(lon0.astype(np.float64)).pin_memory()
h_lat = torch.from_numpy(lat0.astype(np.float64)).pin_memory()
h_z = torch.from_numpy(z0).pin_memory()
o_lon, o_lat = torch.empty_like(h_lon).pin_memory(), torch.empty_like(h_lat).pin_memory()
t, dt = times[0] + timedelta(seconds=300), timedelta(seconds=600)
for i in range(3):
    eng.advect_current_host(grp, 'runge-kutta4', t, dt, h_lon, h_lat, h_z, o_lon, o_lat, chunks=chunks)
    print('---', file=sys.stderr)
