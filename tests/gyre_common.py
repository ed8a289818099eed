"""Helpers of the double-gyre (BASELINE configs[0]) tests: fixture loading and the ways a fixture is replayed --
oracle port on CPU, host-compiled device math on CPU, CUDA library / drop-in model on the GPU."""
import ctypes as C
import glob
import json
import os
from datetime import datetime, timedelta

import numpy as np

import common
from opendrift_b200 import _lib
from opendrift_b200.readers import reader_double_gyre

SCHEMES = {'euler': 0, 'runge-kutta': 1, 'runge-kutta4': 2}
R_SPHERE = 6.371e6


def gyre_fixtures(example=None):
    names = sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(common.GOLDEN, 'ref_gyre_*.npz')))
    if example is None:
        return names
    return [n for n in names if ('example' in n) == example]


class GyreFixture:
    def __init__(self, name):
        d = np.load(os.path.join(common.GOLDEN, 'ref_%s.npz' % name))
        self.name, self.meta = name, json.loads(str(d['meta']))
        self.seed_lon, self.seed_lat, self.lon, self.lat = d['seed_lon'], d['seed_lat'], d['lon'], d['lat']
        self.cdf = d['cdf'] if 'cdf' in d.files else None
        self.n = len(self.seed_lon)
        m = self.meta
        self.scheme, self.dt, self.steps = m['scheme'], m['dt'], m['steps']
        self.t0 = datetime.fromisoformat(m['initial_time'])

    def product_reader(self):
        m = self.meta
        return reader_double_gyre.Reader(initial_time=self.t0, epsilon=m['epsilon'], omega=m['omega'], A=m['A'], proj4=m['proj4'])

    def port_reader(self):
        from oracle import gyre_port
        m = self.meta
        return gyre_port.DoubleGyreReader(self.t0, epsilon=m['epsilon'], omega=m['omega'], A=m['A'], proj4=m['proj4'])


def run_port(fx):
    from oracle import advect_port as ap
    kw = {} if fx.cdf is None else {'cdf': fx.cdf}
    lon, lat, _ = ap.run_oceandrift([fx.port_reader()], fx.seed_lon, fx.seed_lat, np.zeros(fx.n), fx.t0, fx.dt, fx.steps,
                                    scheme=fx.scheme, **kw)
    return lon, lat


def plane_error_m(fx_or_reader, lon, lat, rlon, rlat):
    """max distance (m) between two position sets: at the reader's origin one degree is R*pi/180 metres"""
    k = R_SPHERE * np.pi / 180.0
    return float(np.max(np.hypot((np.asarray(lon) - rlon) * k, (np.asarray(lat) - rlat) * k)))


def _factor(fx):
    """factor * current_drift_factor as the reference forms it: float64 when the property was seeded as a scalar,
    float32 when seeded as an array (elements.py:213-216)."""
    if fx.cdf is None:
        return np.ones(fx.n, dtype=np.float32 if fx.n == 1 else np.float64)     # no promotion for a single element (:213)
    return np.asarray(fx.cdf, dtype=np.float32)


def stage_times(fx, step):
    ts = timedelta(seconds=fx.dt)
    t = fx.t0 + step * ts if fx.dt > 0 else fx.t0 + step * ts
    return tuple((x - fx.t0).total_seconds() for x in (t, t + ts / 2, t + ts))


def run_hostshim(fx, mode=_lib.OD_MATH_SERIES):
    lib = common.hostshim()
    desc = fx.product_reader()
    desc.bind(None, {'x_sea_water_velocity': 0.0, 'y_sea_water_velocity': 0.0})
    d = desc.analytic_desc()
    lon = np.asarray(fx.seed_lon, dtype=np.float32).astype(np.float64)       # seed_elements casts to float32
    lat = np.asarray(fx.seed_lat, dtype=np.float32).astype(np.float64)
    fac = _factor(fx)
    for k in range(fx.steps):
        a = _lib.AnalyticAdvectArgs()
        a.scheme, a.math = SCHEMES[fx.scheme], mode
        a.factor_f64 = 1 if fac.dtype == np.float64 else 0
        a.pos_f32 = 1 if k == 0 else 0
        a.t_start, a.t_mid, a.t_end = stage_times(fx, k)
        a.dt, a.n = fx.dt, fx.n
        a.d_lon, a.d_lat = lon.ctypes.data, lat.ctypes.data
        a.d_factor = fac.ctypes.data
        rc = lib.hs_analytic_advect(C.byref(d), C.byref(a))
        assert rc == 0
    return lon, lat


def run_engine(fx, mode=None):
    """The C-ABI on the GPU: od_analytic_advect step by step."""
    from opendrift_b200.engine import default_engine
    eng = default_engine()
    rd = fx.product_reader()
    rd.bind(eng, {'x_sea_water_velocity': 0.0, 'y_sea_water_velocity': 0.0})
    d = rd.analytic_desc()
    lon = eng.to_device(np.asarray(fx.seed_lon, dtype=np.float32).astype(np.float64))
    lat = eng.to_device(np.asarray(fx.seed_lat, dtype=np.float32).astype(np.float64))
    fac = eng.to_device(_factor(fx))
    for k in range(fx.steps):
        eng.analytic_advect(d, fx.scheme, stage_times(fx, k), fx.dt, lon, lat, factor=fac, pos_f32=(k == 0), fast=mode)
    eng.sync()
    return lon.cpu().numpy(), lat.cpu().numpy()


def run_model(fx, arithmetic=None):
    """The drop-in classes, as the example script uses them."""
    from opendrift_b200.models.oceandrift import OceanDrift
    o = OceanDrift(loglevel=50)
    o.set_config('environment:fallback:land_binary_mask', 0)
    o.set_config('general:use_auto_landmask', False)
    o.set_config('drift:advection_scheme', fx.scheme)
    if arithmetic is not None:
        o.set_config('gpu:arithmetic', arithmetic)
    rd = fx.product_reader()
    o.add_reader(rd)
    kw = {} if fx.cdf is None else {'current_drift_factor': fx.cdf}
    o.seed_elements(fx.seed_lon, fx.seed_lat, time=rd.initial_time, **kw)
    o.run(steps=fx.steps, time_step=fx.dt)
    return np.asarray(o.elements.lon), np.asarray(o.elements.lat)


ASPECTS = ['+proj=stere +lat_0=60 +lon_0=10 +R=6371000 +x_0=0.3 +y_0=-0.2 +units=m +no_defs',
           '+proj=stere +lat_0=-35 +lon_0=170 +R=6371000 +x_0=-1 +y_0=-0.5 +units=m +no_defs',
           '+proj=stere +lat_0=90 +lon_0=70 +lat_ts=60 +R=6371000 +x_0=-1500000 +y_0=-1000000 +units=m +no_defs',
           '+proj=stere +lat_0=-90 +lon_0=-30 +R=6371000 +x_0=800000 +y_0=-1200000 +units=m +no_defs',
           '+proj=stere +lat_0=0 +lon_0=-179.99999 +R=6371000 +x_0=-1 +units=m +no_defs']


class AspectCase:
    """The double gyre on another stereographic plane (the reader takes any proj4, reader_double_gyre.py:29-31): a
    fixture-like object whose reference result is the port's (checked against the live reference in tests/test_gyre.py)."""

    def __init__(self, proj4, n=200, steps=30, dt=0.1, scheme='runge-kutta4'):
        from oracle import advect_port as ap, gyre_port
        from oracle.proj_stere import Stere
        self.proj4, self.n, self.steps, self.dt, self.scheme, self.cdf = proj4, n, steps, dt, scheme, None
        self.t0 = datetime(2000, 1, 1)
        self.par = dict(epsilon=.25, omega=.628, A=.25)
        rng = np.random.default_rng(len(proj4))
        self.seed_lon, self.seed_lat = self.product_reader().xy2lonlat(rng.uniform(0.05, 1.95, n), rng.uniform(0.05, 0.95, n))
        self.plane = Stere(proj4)
        pr = gyre_port.DoubleGyreReader(self.t0, proj4=proj4, **self.par)
        self.lon, self.lat, _ = ap.run_oceandrift([pr], self.seed_lon, self.seed_lat, np.zeros(n), self.t0, dt, steps, scheme=scheme)

    def product_reader(self):
        return reader_double_gyre.Reader(initial_time=self.t0, proj4=self.proj4, **self.par)

    def error_m(self, lon, lat):
        x, y = self.plane.forward(lon, lat)
        rx, ry = self.plane.forward(self.lon, self.lat)
        return float(np.max(np.hypot(x - rx, y - ry)))
