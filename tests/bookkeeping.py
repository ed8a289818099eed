"""Run-loop bookkeeping cases (release, age / retirement, deactivate_outside, compaction order) shared by
the CPU (host engine) and GPU tests.  The expected results come from the UNMODIFIED reference: tests/golden/bookkeeping_ref.npz,
written by `python tests/bookkeeping.py` in the build container (oracle/refrun.py)."""
import os
from datetime import timedelta

import numpy as np

import common

GOLDEN = os.path.join(common.GOLDEN, 'bookkeeping_ref.npz')
N, STEPS = 300, 8
CASES = ['linear_release', 'release_max_age', 'release_deactivate_north', 'per_element_times']


def case_setup(fx, case):
    """(release time spec, extra config) of a case; fx = common.Fixture('rk4_3d')"""
    extra = {}
    t = [fx.start, fx.start + timedelta(seconds=4 * fx.dt)]           # linear release over the first four steps
    if case == 'release_max_age':
        extra = {'drift:max_age_seconds': 3000}
    elif case == 'release_deactivate_north':
        extra = {'drift:deactivate_north_of': float(np.median(fx.lat0[:N])) + 0.01}
    elif case == 'per_element_times':
        rng = np.random.default_rng(1)
        t = [fx.start + timedelta(seconds=float(fx.dt * k)) for k in rng.integers(0, 5, N)]
    cfg = {'drift:advection_scheme': 'runge-kutta4', 'drift:vertical_advection': False}
    cfg.update(extra)
    return t, cfg


def run_product(fx, case, **model_kw):
    from opendrift_b200.models.oceandrift import OceanDrift
    from opendrift_b200.readers import reader_regular_grid
    t, cfg = case_setup(fx, case)
    o = OceanDrift(loglevel=50, **model_kw)
    o.add_reader(reader_regular_grid.Reader(fx.grid_lon, fx.grid_lat, fx.grid_z, fx.times, {common.CUR[0]: fx.u, common.CUR[1]: fx.v}))
    o.set_config('general:use_auto_landmask', False)
    o.set_config('general:coastline_action', 'none')
    for k, v in cfg.items():
        o.set_config(k, v)
    o.seed_elements(lon=fx.lon0[:N], lat=fx.lat0[:N], z=fx.z0[:N], time=t)
    o.run(steps=STEPS, time_step=fx.dt, time_step_output=fx.dt)
    return o


def summary(o, deactivated_len):
    el, de = o.elements, o.elements_deactivated
    out = {'id': np.asarray(el.ID, dtype=np.int64), 'lon': np.asarray(el.lon, dtype=np.float64), 'lat': np.asarray(el.lat, dtype=np.float64),
           'age': np.asarray(el.age_seconds, dtype=np.float64)}
    if deactivated_len:
        out.update({'d_id': np.asarray(de.ID, dtype=np.int64), 'd_lon': np.asarray(de.lon, dtype=np.float64),
                    'd_lat': np.asarray(de.lat, dtype=np.float64), 'd_status': np.asarray(de.status, dtype=np.int64)})
    else:
        out.update({'d_id': np.zeros(0, np.int64), 'd_lon': np.zeros(0), 'd_lat': np.zeros(0), 'd_status': np.zeros(0, np.int64)})
    return out


def check(o, case):
    """Compare a finished product run with the reference's result of the same case."""
    ref = np.load(GOLDEN)
    got = summary(o, o.num_elements_deactivated())
    g = lambda k: ref['%s__%s' % (case, k)]                      # noqa: E731
    assert np.array_equal(got['id'], g('id')) and np.array_equal(got['age'], g('age'))
    assert np.array_equal(got['d_id'], g('d_id'))                # same elements, same (concatenation) order
    assert np.array_equal(got['d_status'], g('d_status'))
    if len(got['id']):
        assert max(common.max_err_deg(got['lon'], got['lat'], g('lon'), g('lat'))) < 5e-8
    if len(got['d_id']):
        assert max(common.max_err_deg(got['d_lon'], got['d_lat'], g('d_lon'), g('d_lat'))) < 5e-8
    return len(got['id']), len(got['d_id'])


if __name__ == '__main__':
    from oracle import refrun
    fx = common.Fixture('rk4_3d')
    out = {}
    for case in CASES:
        t, cfg = case_setup(fx, case)
        rd = refrun.make_grid_reader(fx.grid_lon, fx.grid_lat, fx.grid_z, fx.times, {common.CUR[0]: fx.u, common.CUR[1]: fx.v})
        ro = refrun.run_oceandrift([rd], fx.lon0[:N], fx.lat0[:N], fx.z0[:N], t, fx.dt, STEPS, config=cfg)
        s = summary(ro, len(ro.elements_deactivated))
        for k, v in s.items():
            out['%s__%s' % (case, k)] = v
        print(case, 'active', len(s['id']), 'deactivated', len(s['d_id']))
    np.savez_compressed(GOLDEN, **out)
    print('wrote', GOLDEN)
