"""reader_constant_2d (static arrays on a grid) and ContinuousReader subclasses (the reference's extension point for analytical readers, basereader/continuous.py: `get_variables`
returns exact values at the positions it is given) -- reader_oscillating and a user-written reader, the same scripts on the
reference's classes and on the product's.  Expected results from the UNMODIFIED reference: tests/golden/cont_ref.npz, written by
`python tests/contcases.py` in the build container."""
import os
from datetime import timedelta

import numpy as np

import common

GOLDEN = os.path.join(common.GOLDEN, 'cont_ref.npz')
N, STEPS = 200, 10


def user_reader_class(Base):
    """A reader as a user of the reference would write one: current depending on position and depth, valid in a box."""
    class Reader(Base):
        def __init__(self, lon0, lon1, lat0, lat1):
            self.variables = ['x_sea_water_velocity', 'y_sea_water_velocity']
            self.proj4 = '+proj=latlong'
            self.xmin, self.xmax, self.ymin, self.ymax = lon0, lon1, lat0, lat1
            self.start_time = self.end_time = self.time_step = None
            self.name = 'user_reader'
            super().__init__()

        def get_variables(self, variables, time=None, x=None, y=None, z=None):
            variables, time, x, y, z, outside = self.check_arguments(variables, time, x, y, z)     # (as the reference's readers do)
            assert len(outside) == 0
            s = (time - common.syn.T0).total_seconds()
            x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
            depth = np.exp(np.asarray(z, dtype=np.float64) / 40.0)
            return {'time': time, 'x': x, 'y': y, 'z': z,
                    'x_sea_water_velocity': (0.9 * np.sin(3.0 * x) * np.cos(2.0 * y) + 0.2 * np.sin(s / 2000.0)) * depth,
                    'y_sea_water_velocity': 0.7 * np.cos(2.5 * x + s / 3000.0) * np.sin(1.5 * y) * depth}
    return Reader


# name -> (config, depths?)
CASES = {
    'oscillating_current_rk4': {'drift:advection_scheme': 'runge-kutta4'},
    'oscillating_wind_grid_current_rk2': {'drift:advection_scheme': 'runge-kutta'},
    'user_reader_rk4_depths': {'drift:advection_scheme': 'runge-kutta4'},
    'user_reader_euler_uncertainty': {'drift:advection_scheme': 'euler', 'drift:current_uncertainty': 0.05},
}


def run_case(case, Model, make_grid_reader, oscillating, ContinuousBase, **model_kw):
    fx = common.Fixture('rk4_2d')
    o = Model(loglevel=50, **model_kw)
    z = np.zeros(N, dtype=np.float32)
    kw = {}
    if case == 'oscillating_current_rk4':
        o.add_reader(oscillating.Reader('x_sea_water_velocity', amplitude=1.2, period=timedelta(hours=3), zero_time=fx.start))
        o.add_reader(oscillating.Reader('y_sea_water_velocity', amplitude=-0.8, period=timedelta(hours=5), zero_time=fx.start - timedelta(hours=1)))
    elif case == 'oscillating_wind_grid_current_rk2':
        o.add_reader(make_grid_reader(fx.grid_lon, fx.grid_lat, None, fx.times, {common.CUR[0]: fx.u, common.CUR[1]: fx.v}, 'current'))
        o.add_reader(oscillating.Reader('x_wind', amplitude=14.0, period=timedelta(hours=2), zero_time=fx.start - timedelta(minutes=20)))
        kw['wind_drift_factor'] = 0.03
    else:
        R = user_reader_class(ContinuousBase)
        lon, lat = fx.lon0[:N], fx.lat0[:N]
        # (the box leaves out the eastern fifth of the cloud: those elements get the fallback value, 0)
        o.add_reader(R(float(lon.min()) - 0.5, float(np.percentile(lon, 80)), float(lat.min()) - 0.5, float(lat.max()) + 0.5))
        if 'depths' in case:
            z = -np.linspace(0, 60, N).astype(np.float32)
    for key, val in {'general:use_auto_landmask': False, 'environment:constant:land_binary_mask': 0, 'general:coastline_action': 'none',
                     'drift:vertical_advection': False}.items():
        o.set_config(key, val)
    for key, val in CASES[case].items():
        o.set_config(key, val)
    np.random.seed(7)
    o.seed_elements(lon=fx.lon0[:N], lat=fx.lat0[:N], z=z, time=[fx.start, fx.start + timedelta(seconds=3 * fx.dt)], **kw)
    o.run(steps=STEPS, time_step=fx.dt, time_step_output=fx.dt)
    return o


def run_product(case, **model_kw):
    from opendrift_b200.models.oceandrift import OceanDrift
    from opendrift_b200.readers import reader_regular_grid, reader_oscillating
    from opendrift_b200.readers.continuous import ContinuousReader
    return run_case(case, OceanDrift, lambda lon, lat, z, t, f, name: reader_regular_grid.Reader(lon, lat, z, t, f, name=name),
                    reader_oscillating, ContinuousReader, **model_kw)


# reader_constant_2d (static 2-D arrays): the current AND the land mask of a run -> (scheme, coastline action)
C2D_CASES = {'constant_2d_stranding_euler': ('euler', 'stranding'), 'constant_2d_previous_rk4': ('runge-kutta4', 'previous')}


def run_c2d(case, Model, constant_2d, **model_kw):
    import coastcases
    scheme, action = C2D_CASES[case]
    fx = common.Fixture('rk4_2d')
    mlon, mlat, mask = coastcases.mask_grid(fx, True)
    o = Model(loglevel=50, **model_kw)
    o.add_reader(constant_2d.Reader(fx.grid_lon.astype(np.float64), fx.grid_lat.astype(np.float64),
                                    {common.CUR[0]: 5 * fx.u[1], common.CUR[1]: 5 * fx.v[1]}))
    o.add_reader(constant_2d.Reader(mlon, mlat, {'land_binary_mask': mask}))
    for key, val in {'general:use_auto_landmask': False, 'general:coastline_action': action, 'general:coastline_approximation_precision': None,
                     'drift:advection_scheme': scheme, 'seed:ocean_only': False, 'drift:vertical_advection': False}.items():
        o.set_config(key, val)
    np.random.seed(3)
    o.seed_elements(lon=fx.lon0[:300], lat=fx.lat0[:300], time=[fx.start, fx.start + timedelta(seconds=1800)])
    o.run(steps=8, time_step=fx.dt, time_step_output=fx.dt)
    return o


def run_c2d_product(case, **model_kw):
    from opendrift_b200.models.oceandrift import OceanDrift
    from opendrift_b200.readers import reader_constant_2d
    return run_c2d(case, OceanDrift, reader_constant_2d, **model_kw)


def summary_c2d(o):
    out = summary(o)
    de = o.elements_deactivated
    out.update({'d_id': np.asarray(de.ID, dtype=np.int64), 'd_status': np.asarray(de.status, dtype=np.int64),
                'd_lon': np.asarray(de.lon, dtype=np.float64), 'd_lat': np.asarray(de.lat, dtype=np.float64),
                'cats': np.array(list(o.status_categories))})
    return out


def check_c2d(o, case):
    ref = np.load(GOLDEN)
    got = summary_c2d(o)
    g = lambda k: ref['%s__%s' % (case, k)]                      # noqa: E731
    assert list(got['cats']) == list(g('cats'))
    assert np.array_equal(got['id'], g('id')) and np.array_equal(got['d_id'], g('d_id')) and np.array_equal(got['d_status'], g('d_status'))
    assert max(common.max_err_deg(got['lon'], got['lat'], g('lon'), g('lat'))) < 5e-8
    assert max(common.max_err_deg(got['d_lon'], got['d_lat'], g('d_lon'), g('d_lat'))) < 5e-7      # ('previous': float32 positions, coastcases.py)
    return len(got['id']), len(got['d_id'])


def reader_queries(oscillating, ContinuousBase):
    """get_variables_interpolated of the two reader kinds at a handful of positions and times."""
    fx = common.Fixture('rk4_2d')
    out = {}
    r = oscillating.Reader('sea_surface_height', amplitude=1.0, period=timedelta(hours=6), zero_time=fx.start)
    lon, lat = np.array([3.0, 4.0, -170.0]), np.array([60.0, 61.0, -20.0])
    for i, dt in enumerate((0, 1800, 12345)):
        env, _ = r.get_variables_interpolated(['sea_surface_height'], time=fx.start + timedelta(seconds=dt), lon=lon, lat=lat, z=np.zeros(3))
        out['osc_%d' % i] = np.ma.filled(np.ma.masked_invalid(np.asarray(env['sea_surface_height'], dtype=np.float64)), np.nan)
    R = user_reader_class(ContinuousBase)(2.0, 3.2, 56.0, 57.0)
    lon, lat, z = fx.lon0[:50].astype(np.float64), fx.lat0[:50].astype(np.float64), -np.linspace(0, 30, 50)
    env, prof = R.get_variables_interpolated(['x_sea_water_velocity', 'y_sea_water_velocity'], profiles=['x_sea_water_velocity'], profiles_depth=20,
                                             time=fx.start + timedelta(seconds=700), lon=lon, lat=lat, z=z)
    for k in ('x_sea_water_velocity', 'y_sea_water_velocity'):
        out['user_' + k] = np.ma.filled(np.ma.masked_invalid(np.asarray(env[k], dtype=np.float64)), np.nan)
    out['user_profile'] = np.ma.filled(np.ma.masked_invalid(np.asarray(prof['x_sea_water_velocity'], dtype=np.float64)), np.nan)
    out['user_profile_z'] = np.asarray(prof['z'], dtype=np.float64)
    return out


def summary(o):
    el = o.elements
    return {'id': np.asarray(el.ID, dtype=np.int64), 'lon': np.asarray(el.lon, dtype=np.float64), 'lat': np.asarray(el.lat, dtype=np.float64),
            'z': np.asarray(el.z, dtype=np.float64)}


def check(o, case):
    ref = np.load(GOLDEN)
    got = summary(o)
    g = lambda k: ref['%s__%s' % (case, k)]                      # noqa: E731
    assert np.array_equal(got['id'], g('id'))
    err = max(common.max_err_deg(got['lon'], got['lat'], g('lon'), g('lat')))
    assert err < 5e-8, err
    assert np.array_equal(got['z'], g('z'))
    return err


if __name__ == '__main__':
    from oracle import refrun
    refrun.setup()
    from opendrift.models.oceandrift import OceanDrift as RefOD
    from opendrift.readers import reader_oscillating as ref_osc
    from opendrift.readers.basereader.continuous import ContinuousReader as RefCont
    out = {}
    for case in CASES:
        ro = run_case(case, RefOD, lambda lon, lat, z, t, f, name: refrun.make_grid_reader(lon, lat, z, t, f, name=name), ref_osc, RefCont,
                      logfile='/tmp/od_cont.log')
        for k, v in summary(ro).items():
            out['%s__%s' % (case, k)] = v
        print(case, 'active', ro.num_elements_active(), 'lon', float(np.min(ro.elements.lon)), float(np.max(ro.elements.lon)))
    from opendrift.readers import reader_constant_2d as ref_c2d
    for case in C2D_CASES:
        ro = run_c2d(case, RefOD, ref_c2d, logfile='/tmp/od_cont.log')
        for k, v in summary_c2d(ro).items():
            out['%s__%s' % (case, k)] = v
        print(case, 'active', ro.num_elements_active(), 'deactivated', ro.num_elements_deactivated(), list(ro.status_categories))
    for k, v in reader_queries(ref_osc, RefCont).items():
        out['query__' + k] = v
    np.savez_compressed(GOLDEN, **out)
    print('wrote', GOLDEN)
