"""Random stereographic / Mercator / Lambert-conformal planes (random parameters and ellipsoids) under four model configurations: the drop-in
classes on the host build of the device sources beside the UNMODIFIED reference (its pyproj = oracle/proj_stere.py, oracle/proj_conformal.py).
Run in the build container: python tools/fuzz_proj_vs_reference.py"""
import sys
import os; R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np, projcases as pc, common
from hostengine import HostEngine
import opendrift_b200.engine as E, opendrift_b200.models.basemodel as B
eng = HostEngine()
E.default_engine = lambda device=None: eng
B.default_engine = lambda device=None: eng
from oracle import refrun
refrun.setup()
from opendrift.models.oceandrift import OceanDrift as RefOD
from opendrift.models.leeway import Leeway as RefLW
from opendrift_b200.models.oceandrift import OceanDrift
from opendrift_b200.models.leeway import Leeway
from opendrift_b200.readers import reader_regular_grid
rng = np.random.default_rng(11)
ell = ['+ellps=WGS84', '+ellps=GRS80', '+R=6371000', '+a=6378137 +rf=298.3', '+a=6378388 +es=0.0067']
def rand_proj():
    k = rng.integers(0, 4)
    e = ell[rng.integers(0, len(ell))]
    off = ' +x_0=%d +y_0=%d' % (rng.integers(-5e5, 5e5), rng.integers(-5e5, 5e5))
    if k == 0:
        return '+proj=merc +lon_0=%.2f +lat_ts=%.2f %s%s +units=m +no_defs' % (rng.uniform(-30, 30), rng.uniform(0, 75), e, off)
    if k == 1:
        l1, l2 = rng.uniform(40, 75), rng.uniform(40, 75)
        return '+proj=lcc +lat_1=%.2f +lat_2=%.2f +lat_0=%.2f +lon_0=%.2f +k_0=%.5f %s%s +units=m +no_defs' % (l1, l2, rng.uniform(45, 70), rng.uniform(-20, 30), rng.uniform(0.99, 1.0), e, off)
    if k == 2:
        return '+proj=stere +lat_0=90 +lon_0=%.2f +lat_ts=%.2f %s%s +units=m +no_defs' % (rng.uniform(-60, 60), rng.uniform(50, 90), e, off)
    return '+proj=stere +lat_0=%.2f +lon_0=%.2f +k_0=%.5f %s%s +units=m +no_defs' % (rng.uniform(30, 75), rng.uniform(-20, 30), rng.uniform(0.99, 1.0), e, off)
models = [('OceanDrift', ('cur3d',), {'drift:advection_scheme': 'runge-kutta4'}, {}), ('OceanDrift', ('cur2d', 'wind'), {'drift:advection_scheme': 'runge-kutta', 'drift:vertical_advection': False}, {'z': 0.0}),
          ('Leeway', ('cur2d', 'wind'), {}, {}), ('OceanDrift', ('cur3d_k',), {'drift:vertical_mixing': True, 'drift:vertical_advection': False, 'vertical_mixing:timestep': 60.0}, {})]
bad = 0
for it in range(12):
    p4 = rand_proj()
    m = models[it % 4]
    pc.CASES['fuzz'] = dict(proj4=p4, model=m[0], readers=m[1], steps=5, dt=600, cfg=m[2], seed=m[3])
    try:
        r = pc.run_case('fuzz', {'OceanDrift': RefOD, 'Leeway': RefLW}, lambda x, y, z, t, f, name, proj4: refrun.make_grid_reader(x, y, z, t, f, name=name, proj4=proj4), logfile='/tmp/fz.log')
        p = pc.run_case('fuzz', {'OceanDrift': OceanDrift, 'Leeway': Leeway}, lambda x, y, z, t, f, name, proj4: reader_regular_grid.Reader(x, y, z, t, f, name=name, proj4=proj4))
        e = max(common.max_err_deg(np.asarray(p.elements.lon), np.asarray(p.elements.lat), np.asarray(r.elements.lon), np.asarray(r.elements.lat)))
        dz = float(np.abs(np.asarray(p.elements.z, dtype=np.float64) - np.asarray(r.elements.z, dtype=np.float64)).max())
        flag = '' if e < 5e-8 and dz < 1e-5 else '   <<<<<<'
        bad += flag != ''
        print(it, m[0], m[1], 'err %.2e dz %.1e' % (e, dz), p4[:70], flag)
    except Exception as ex:
        bad += 1
        print(it, 'EXC', repr(ex)[:200], p4)
print('bad', bad)
