"""Per-step timeline of Leeway.run() at 20 M particles (BASELINE configs[4] leg of the bench): a CUDA event before every step launch
and the host time of it; prints the step periods on the device and on the host, to see what a step costs beyond its kernel."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_legs as bl                                   # noqa: E402
from opendrift_b200 import synthetic as syn              # noqa: E402
from opendrift_b200.engine import Engine                 # noqa: E402
from opendrift_b200.models.leeway import Leeway          # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
steps = 40
eng = Engine(0)
grid, host = bl.host_fields('cfg5', three_d=False)
dev = bl.to_device(host, eng, torch)
n_times = syn.n_slabs_for(steps + 2, bl.DT) + 1
o = Leeway(loglevel=50, seed=0, engine=eng)
for r in bl.product_readers(grid, dev, n_times):
    o.add_reader(r)
o.set_config('general:use_auto_landmask', False)
o.set_config('gpu:rng', 'philox')
lon0, lat0, _ = syn.particle_cloud(n, seed=31, three_d=False)
o.seed_elements(lon=lon0, lat=lat0, time=syn.T0, object_type=1)
ev, stamps, names = [], [], []
for meth in ('leeway_step', 'bookkeeping', 'partition_active', 'sort_by_cell', 'permute', 'upload', 'fill_nan'):
    orig = getattr(eng, meth)

    def wrapped(*a, _orig=orig, _m=meth, **k):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        ev.append(e)
        stamps.append(time.perf_counter())
        names.append(_m)
        r = _orig(*a, **k)
        e2 = torch.cuda.Event(enable_timing=True)
        e2.record()
        ev.append(e2)
        stamps.append(time.perf_counter())
        names.append(_m + ':end')
        return r
    setattr(eng, meth, wrapped)
t0 = time.perf_counter()
o.run(steps=steps, time_step=bl.DT, time_step_output=steps * bl.DT)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
rows = []
for k in range(1, len(ev)):
    rows.append((names[k - 1] + ' -> ' + names[k], ev[k - 1].elapsed_time(ev[k]), (stamps[k] - stamps[k - 1]) * 1e3))
agg = {}
for nm, g, h in rows[len(rows) // 3:]:
    agg.setdefault(nm, []).append((g, h))
out = {nm: {'n': len(v), 'gpu_ms_median': float(np.median([x[0] for x in v])), 'host_ms_median': float(np.median([x[1] for x in v])),
            'gpu_ms_mean': float(np.mean([x[0] for x in v]))} for nm, v in agg.items()}
big = [(k, nm, round(g, 2), round(h, 2)) for k, (nm, g, h) in enumerate(rows) if g > 1.0 and 'leeway_step -> leeway_step:end' not in nm]
print(json.dumps({'n': n, 'steps': steps, 'wall_s_run': wall, 'intervals': out, 'gaps_over_1ms': big}, indent=None))
