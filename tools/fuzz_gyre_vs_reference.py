"""Double-gyre fuzzer (build container only: needs /root/reference): random stereographic plane (all four aspects, across the
dateline), field parameters, scheme, forward / backward, release interval, 1 / 2 / 150 elements, drift-factor variants -- the UNMODIFIED
reference (with oracle/proj_stere.py + oracle/geod_karney.py as pyproj) and the drop-in classes on the host build of
csrc/od_analytic.cuh side by side; positions compared on the reader's plane (1e-4 m).

    python tools/fuzz_gyre_vs_reference.py FIRST_SEED LAST_SEED      (30 configurations were run in round 1: all equal)"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, ROOT)
import numpy as np, common, gyre_common as gc, traceback
from datetime import timedelta
from oracle import refrun
refrun.setup()
from oracle.proj_stere import Stere
from hostengine import HostEngine
bad = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    r = np.random.default_rng(5000 + seed)
    proj4 = r.choice(gc.ASPECTS + ['+proj=stere +lat_0=0 +lon_0=0 +lat_ts=0 +units=m +a=6.371e+06 +e=0 +no_defs'] * 3)
    par = dict(epsilon=float(r.uniform(0.05, 0.3)), omega=float(r.uniform(0.3, 1.0)), A=float(r.uniform(0.1, 0.3)))
    scheme = r.choice(['euler', 'runge-kutta', 'runge-kutta4']); dt = float(r.choice([0.1, 0.05, -0.1])); steps = int(r.integers(5, 40)); n = int(r.choice([1, 2, 150]))
    release = bool(r.integers(3) == 0) and n > 2
    cdf = None if r.integers(2) else (0.7 if r.integers(2) else np.linspace(0.5, 1, n).astype(np.float32))
    out = []
    try:
        for kind in ('ref', 'prod'):
            if kind == 'ref':
                from opendrift.models.oceandrift import OceanDrift as M
                from opendrift.readers import reader_double_gyre as dgm
                o = M(loglevel=50, logfile='/tmp/x.log', seed=0); o.set_config('environment:fallback:land_binary_mask', 0)
            else:
                from opendrift_b200.models.oceandrift import OceanDrift as M
                from opendrift_b200.readers import reader_double_gyre as dgm
                o = M(loglevel=50, seed=0, engine=HostEngine())
            o.set_config('general:use_auto_landmask', False); o.set_config('drift:advection_scheme', scheme)
            dg = dgm.Reader(proj4=proj4, **par)
            o.add_reader(dg)
            rr = np.random.default_rng(seed)
            lon, lat = dg.xy2lonlat(rr.uniform(0.05, 1.95, n), rr.uniform(0.05, 0.95, n))
            t0 = dg.initial_time + timedelta(seconds=20)
            t = [t0, t0 + timedelta(seconds=abs(dt) * 3)] if release else t0
            kw = {} if cdf is None else {'current_drift_factor': cdf}
            o.seed_elements(lon, lat, time=t, **kw)
            o.run(steps=steps, time_step=dt)
            out.append(o)
        ro, po = out
        P = Stere(proj4)
        ok = np.array_equal(np.asarray(ro.elements.ID), np.asarray(po.elements.ID))
        e = -1
        if ok and len(ro.elements.ID):
            rx, ry = P.forward(np.asarray(ro.elements.lon), np.asarray(ro.elements.lat)); px, py = P.forward(np.asarray(po.elements.lon), np.asarray(po.elements.lat))
            e = float(np.max(np.hypot(px - rx, py - ry))); ok = e < 1e-4
        bad += not ok
        print(seed, 'OK ' if ok else 'BAD', 'err m %.1e' % e, '' if ok else (proj4[:40], par, scheme, dt, steps, n, release))
    except BaseException as ex:
        bad += 1; print(seed, 'EXC', repr(ex)[:200], (proj4[:40], scheme, dt, steps, n, release)); traceback.print_exc(limit=3)
print('bad', bad)
