"""Variables without a fallback value (`environment:fallback:<variable>` = None): elements that leave the readers' coverage are taken out
as 'missing_data' at the top of the loop (report_missing_variables, basemodel/__init__.py:2249, 2501-2515); an element whose
Runge-Kutta mid-point lies outside gets an undefined position and leaves one step later.  Cases shared by the CPU (host engine) and
GPU tests; expected results from the UNMODIFIED reference: tests/golden/missing_ref.npz, written by `python tests/missingcases.py`
in the build container."""
import os
from datetime import timedelta

import numpy as np

import common

GOLDEN = os.path.join(common.GOLDEN, 'missing_ref.npz')
N, STEPS = 300, 14
NONE_CUR = {'environment:fallback:x_sea_water_velocity': None, 'environment:fallback:y_sea_water_velocity': None}
NONE_WIND = {'environment:fallback:x_wind': None, 'environment:fallback:y_wind': None}

# name -> (fixture, config, release over steps, fraction of the columns the current reader covers, the same for a wind reader or None)
CASES = {
    'euler_2d': ('rk4_2d', dict(NONE_CUR, **{'drift:advection_scheme': 'euler'}), 0, 0.6, None),
    'rk4_2d': ('rk4_2d', dict(NONE_CUR, **{'drift:advection_scheme': 'runge-kutta4'}), 0, 0.6, None),
    'rk2_diffusion_release': ('rk4_2d', dict(NONE_CUR, **{'drift:advection_scheme': 'runge-kutta', 'environment:constant:horizontal_diffusivity': 10.0}), 4, 0.6, None),
    'rk4_3d_w': ('rk4_3d', dict(NONE_CUR, **{'drift:advection_scheme': 'runge-kutta4', 'drift:vertical_advection': True}), 3, 0.62, None),
    # a drift:deactivate_east_of limit just inside the edge of the coverage: elements can be beyond it and without data on the same step
    'east_limit_rk2': ('rk4_2d', dict(NONE_CUR, **{'drift:advection_scheme': 'runge-kutta', 'drift:deactivate_east_of': 3.147}), 0, 0.6, None),
    'wind_leaves_euler': ('rk4_2d', dict(NONE_WIND, **{'drift:advection_scheme': 'euler'}), 0, 1.0, 0.55),
    'wind_and_current_rk4': ('rk4_2d', dict(NONE_WIND, **dict(NONE_CUR, **{'drift:advection_scheme': 'runge-kutta4', 'drift:current_uncertainty': 0.05})), 3, 0.7, 0.6),
}
SPEED = 6.0


def run_case(case, Model, make_reader, **model_kw):
    """The same script on the reference's OceanDrift (generator) and on the product's."""
    fxname, cfg, release, cut, wcut = CASES[case]
    fx = common.Fixture(fxname)
    k = int(len(fx.grid_lon) * cut)
    c = np.ascontiguousarray
    o = Model(loglevel=50, **model_kw)
    fields = {common.CUR[0]: c((SPEED * fx.u[..., :k]).astype(np.float32)), common.CUR[1]: c((SPEED * fx.v[..., :k]).astype(np.float32))}
    if fx.grid_z is not None and cfg.get('drift:vertical_advection'):
        w = 0.002 * np.sin(np.arange(fx.u.size, dtype=np.float64).reshape(fx.u.shape) * 0.37)
        fields['upward_sea_water_velocity'] = c(w[..., :k].astype(np.float32))
    o.add_reader(make_reader(fx.grid_lon[:k], fx.grid_lat, fx.grid_z, fx.times, fields, 'current'))
    if wcut is not None:
        kw_ = int(len(fx.grid_lon) * wcut)
        X, Y = np.meshgrid(np.linspace(0, 1, kw_), np.linspace(0, 1, len(fx.grid_lat)))
        wx = np.stack([12.0 * np.cos(0.4 * i) * (1 + 0.3 * np.sin(np.pi * X)) for i in range(len(fx.times))]).astype(np.float32)
        wy = np.stack([-8.0 * np.sin(0.4 * i) * (1 + 0.3 * np.cos(np.pi * Y)) for i in range(len(fx.times))]).astype(np.float32)
        o.add_reader(make_reader(fx.grid_lon[:kw_], fx.grid_lat, None, fx.times, {'x_wind': wx, 'y_wind': wy}, 'wind'))
    for key, val in {'general:use_auto_landmask': False, 'environment:constant:land_binary_mask': 0, 'general:coastline_action': 'none',
                     'drift:vertical_advection': False}.items():
        o.set_config(key, val)
    for key, val in cfg.items():
        o.set_config(key, val)
    t = fx.start if not release else [fx.start, fx.start + timedelta(seconds=release * fx.dt)]
    # the N elements of the fixture closest to the edge of the coverage, inside it: they leave during the run, not at its start
    edge = min(float(fx.grid_lon[k - 1]), float(fx.grid_lon[int(len(fx.grid_lon) * wcut) - 1]) if wcut is not None else 1e9)
    inside = np.flatnonzero(fx.lon0 < edge - 0.002)
    sel = np.sort(inside[np.argsort(-fx.lon0[inside], kind='stable')[:N]])
    z = fx.z0[sel] if fx.grid_z is not None else (np.zeros(N, dtype=np.float32) if wcut is not None else -np.linspace(0, 5, N).astype(np.float32))
    np.random.seed(5)
    o.seed_elements(lon=fx.lon0[sel], lat=fx.lat0[sel], z=z, time=t, wind_drift_factor=0.03)
    o.run(steps=STEPS, time_step=fx.dt, time_step_output=fx.dt)
    return o


def run_product(case, **model_kw):
    from opendrift_b200.models.oceandrift import OceanDrift
    from opendrift_b200.readers import reader_regular_grid
    return run_case(case, OceanDrift, lambda lon, lat, z, t, f, name: reader_regular_grid.Reader(lon, lat, z, t, f, name=name), **model_kw)


def summary(o):
    el, de = o.elements, o.elements_deactivated
    out = {'id': np.asarray(el.ID, dtype=np.int64), 'lon': np.asarray(el.lon, dtype=np.float64), 'lat': np.asarray(el.lat, dtype=np.float64),
           'z': np.asarray(el.z, dtype=np.float64), 'cats': np.array(list(o.status_categories))}
    if o.num_elements_deactivated():
        out.update({'d_id': np.asarray(de.ID, dtype=np.int64), 'd_lon': np.asarray(de.lon, dtype=np.float64),
                    'd_lat': np.asarray(de.lat, dtype=np.float64), 'd_status': np.asarray(de.status, dtype=np.int64)})
    else:
        out.update({'d_id': np.zeros(0, np.int64), 'd_lon': np.zeros(0), 'd_lat': np.zeros(0), 'd_status': np.zeros(0, np.int64)})
    return out


def _err(lon, lat, rlon, rlat):
    """Largest difference in degrees; undefined positions must be undefined in both."""
    nan = np.isnan(rlon)
    assert np.array_equal(nan, np.isnan(lon)) and np.array_equal(np.isnan(rlat), np.isnan(lat))
    if nan.all():
        return 0.0
    return max(common.max_err_deg(lon[~nan], lat[~nan], rlon[~nan], rlat[~nan]))


def check(o, case):
    ref = np.load(GOLDEN)
    got = summary(o)
    g = lambda k: ref['%s__%s' % (case, k)]                      # noqa: E731
    assert list(got['cats']) == list(g('cats')), (list(got['cats']), list(g('cats')))
    assert np.array_equal(got['id'], g('id'))
    assert np.array_equal(got['d_id'], g('d_id'))                # the same elements left, in the same order
    assert np.array_equal(got['d_status'], g('d_status'))
    if len(got['id']):
        assert _err(got['lon'], got['lat'], g('lon'), g('lat')) < 5e-8
        assert np.allclose(got['z'], g('z'), rtol=0, atol=1e-5, equal_nan=True)
    if len(got['d_id']):
        assert _err(got['d_lon'], got['d_lat'], g('d_lon'), g('d_lat')) < 5e-8
    return len(got['id']), len(got['d_id']), int(np.isnan(got['d_lon']).sum()), list(got['cats'])


if __name__ == '__main__':
    from oracle import refrun
    refrun.setup()
    from opendrift.models.oceandrift import OceanDrift as RefOD
    out = {}
    for case in CASES:
        ro = run_case(case, RefOD, lambda lon, lat, z, t, f, name: refrun.make_grid_reader(lon, lat, z, t, f, name=name), logfile='/tmp/od_miss.log')
        s = summary(ro)
        for k, v in s.items():
            out['%s__%s' % (case, k)] = v
        print(case, 'active', len(s['id']), 'of them undefined', int(np.isnan(s['lon']).sum()), 'deactivated', len(s['d_id']),
              'of them undefined', int(np.isnan(s['d_lon']).sum()), list(s['cats']))
    np.savez_compressed(GOLDEN, **out)
    print('wrote', GOLDEN)
