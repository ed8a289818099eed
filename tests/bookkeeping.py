"""Run-loop bookkeeping cases (release, age / retirement, deactivate_outside, compaction order) shared by
the CPU (host engine) and GPU tests.  The expected results come from the UNMODIFIED reference: tests/golden/bookkeeping_ref.npz,
written by `python tests/bookkeeping.py` in the build container (oracle/refrun.py)."""
import os
from datetime import timedelta

import numpy as np

import common

GOLDEN = os.path.join(common.GOLDEN, 'bookkeeping_ref.npz')
N, STEPS = 300, 8
CASES = ['linear_release', 'release_max_age', 'release_deactivate_north', 'per_element_times', 'backward_release', 'backward_release_past_reader_end']


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
    elif case == 'backward_release':          # a backward run: released from the last time backwards, IDs flipped (:2056-2062)
        t = [fx.times[-1] - timedelta(seconds=4 * fx.dt), fx.times[-1] - timedelta(seconds=fx.dt)]
    elif case == 'backward_release_past_reader_end':
        # the interpolated release times end 33 microseconds after the reader's last slab: the reference discards the reader
        # for good at the first step (environment.py:430-432) and the elements never move
        t = [fx.times[-1] - timedelta(seconds=3 * fx.dt), fx.times[-1]]
    cfg = {'drift:advection_scheme': 'runge-kutta4', 'drift:vertical_advection': False}
    cfg.update(extra)
    return t, cfg


def case_dt(fx, case):
    return -fx.dt if case.startswith('backward_release') else fx.dt


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
    o.run(steps=STEPS, time_step=case_dt(fx, case), time_step_output=case_dt(fx, case))
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


# ---- several readers for one variable (a nested model inside a coarser one): environment.py:613-780 -----------------------
MULTI_SCHEMES = ['euler', 'runge-kutta', 'runge-kutta4']
MULTI_STEPS = 6


def multireader_fields(fx):
    """Reader A (first in priority): the western half of the fixture's grid; reader B: the whole grid, different values;
    a band in the north is NaN in both (fallback 0)."""
    h = len(fx.grid_lon) // 2
    a = dict(lon=fx.grid_lon[:h], lat=fx.grid_lat, u=fx.u[..., :h].copy(), v=fx.v[..., :h].copy())
    b = dict(lon=fx.grid_lon, lat=fx.grid_lat, u=(0.5 * fx.u).astype(np.float32), v=(-0.7 * fx.v).astype(np.float32))
    return a, b


def run_product_multireader(fx, scheme, cls=None, **model_kw):
    from opendrift_b200.models.oceandrift import OceanDrift
    from opendrift_b200.readers import reader_regular_grid
    a, b = multireader_fields(fx)
    o = (cls or OceanDrift)(loglevel=50, **model_kw)
    o.add_reader([reader_regular_grid.Reader(r['lon'], r['lat'], None, fx.times, {common.CUR[0]: r['u'], common.CUR[1]: r['v']}, name=nm)
                  for nm, r in (('A', a), ('B', b))])
    o.set_config('general:use_auto_landmask', False)
    o.set_config('drift:advection_scheme', scheme)
    o.set_config('drift:vertical_advection', False)
    o.seed_elements(lon=fx.lon0, lat=fx.lat0, z=fx.z0, time=fx.start)
    o.run(steps=MULTI_STEPS, time_step=fx.dt, time_step_output=fx.dt)
    return o


def check_multireader(o, scheme):
    ref = np.load(GOLDEN)
    rl, ra = ref['multireader_%s__lon' % scheme], ref['multireader_%s__lat' % scheme]
    assert max(common.max_err_deg(np.asarray(o.elements.lon), np.asarray(o.elements.lat), rl, ra)) < 5e-8
    # the two readers really disagree: using only the first one (+ fallback) is off by ~1e-2 deg
    return rl, ra



# ---- run() argument variants (basemodel/__init__.py:1829-2060: duration / end_time / steps, output step, config time steps) -----
RUN_N = 200


def run_cases():
    fx = common.Fixture('rk4_2d')
    n = RUN_N
    s = dict(lon=fx.lon0[:n], lat=fx.lat0[:n], time=fx.start)
    m = timedelta(minutes=1)
    return fx, {
        'duration_50min': (s, dict(duration=50 * m, time_step=600), {}),
        'duration_not_a_multiple': (s, dict(duration=55 * m, time_step=600), {}),
        'end_time': (s, dict(end_time=fx.start + 70 * m, time_step=600), {}),
        'output_every_third_step': (s, dict(steps=9, time_step=600, time_step_output=1800), {}),
        'until_reader_end': (s, dict(time_step=900), {}),
        'timedelta_time_step': (s, dict(steps=5, time_step=7 * m), {}),
        'backward_duration': ({**s, 'time': fx.times[-1]}, dict(duration=40 * m, time_step=-600), {}),
        'config_time_steps': (s, dict(steps=4), {'general:time_step_minutes': 12, 'general:time_step_output_minutes': 24}),
        'number_per_point_radius': ({'lon': [3.0, 3.5], 'lat': [57.0, 57.2], 'number_per_point': 30, 'radius': 500, 'time': fx.start},
                                    dict(steps=4, time_step=600), {}),
        'deactivate_west_and_south': (s, dict(steps=8, time_step=600),
                                      {'drift:deactivate_west_of': float(np.median(fx.lon0[:n])),
                                       'drift:deactivate_south_of': float(np.percentile(fx.lat0[:n], 20))}),
    }


def run_product_runcase(case, **model_kw):
    from opendrift_b200.models.oceandrift import OceanDrift
    from opendrift_b200.readers import reader_regular_grid
    fx, cases = run_cases()
    seedkw, runkw, cfg = cases[case]
    o = OceanDrift(loglevel=50, seed=0, **model_kw)
    o.add_reader(reader_regular_grid.Reader(fx.grid_lon, fx.grid_lat, None, fx.times, {common.CUR[0]: fx.u, common.CUR[1]: fx.v}))
    o.set_config('general:use_auto_landmask', False)
    o.set_config('drift:advection_scheme', 'runge-kutta')
    for k, v in cfg.items():
        o.set_config(k, v)
    o.seed_elements(**seedkw)
    o.run(**runkw)
    return o


def check_runcase(o, case):
    ref = np.load(GOLDEN)
    g = lambda k: ref['run_%s__%s' % (case, k)]                  # noqa: E731
    assert o.steps_calculation == int(g('steps')) and np.array_equal(np.asarray(o.elements.ID, dtype=np.int64), g('id'))
    assert (o.time - run_cases()[0].start).total_seconds() == float(g('elapsed'))
    assert max(common.max_err_deg(np.asarray(o.elements.lon), np.asarray(o.elements.lat), g('lon'), g('lat'))) < 5e-8



# ---- three current readers with different extents and level tables + w, wind, noise, diffusion (reader chain, 3-D) ------------
CHAIN3D_N, CHAIN3D_STEPS = 500, 6
CHAIN3D_CFG = {'drift:advection_scheme': 'runge-kutta4', 'environment:constant:horizontal_diffusivity': 5.0,
               'drift:current_uncertainty': 0.05}


def chain3d_readers(fx, make):
    """A: western third on every second level; B: southern half on all levels, other values; C: everywhere, surface only;
    plus the vertical velocity and the wind on readers of their own."""
    nx, ny = len(fx.grid_lon), len(fx.grid_lat)
    cur = common.CUR
    specs = [('A', fx.grid_lon[:nx // 3], fx.grid_lat, fx.grid_z[::2], fx.u[:, ::2, :, :nx // 3].copy(), fx.v[:, ::2, :, :nx // 3].copy()),
             ('B', fx.grid_lon, fx.grid_lat[:ny // 2], fx.grid_z, (0.6 * fx.u[:, :, :ny // 2, :]).astype(np.float32),
              (-0.5 * fx.v[:, :, :ny // 2, :]).astype(np.float32)),
             ('C', fx.grid_lon, fx.grid_lat, None, (0.3 * fx.u[:, 0]).astype(np.float32), (0.9 * fx.v[:, 0]).astype(np.float32))]
    out = [make(lon, lat, z, fx.times, {cur[0]: u, cur[1]: v}, nm) for nm, lon, lat, z, u, v in specs]
    out.append(make(fx.grid_lon, fx.grid_lat, fx.grid_z, fx.times, {'upward_sea_water_velocity': fx.w}, 'w'))
    out.append(make(fx.wind_lon, fx.wind_lat, None, fx.times, {'x_wind': fx.x_wind, 'y_wind': fx.y_wind}, 'wind'))
    return out


def run_product_chain3d(cls=None, **model_kw):
    from opendrift_b200.models.oceandrift import OceanDrift
    from opendrift_b200.readers import reader_regular_grid
    fx = common.Fixture('rk4_3d_full')
    o = (cls or OceanDrift)(loglevel=50, seed=0, **model_kw)
    o.add_reader(chain3d_readers(fx, lambda lon, lat, z, t, f, name: reader_regular_grid.Reader(lon, lat, z, t, f, name=name)))
    o.set_config('general:use_auto_landmask', False)
    for k, v in CHAIN3D_CFG.items():
        o.set_config(k, v)
    n = CHAIN3D_N
    o.seed_elements(lon=fx.lon0[:n], lat=fx.lat0[:n], z=fx.z0[:n], time=fx.start)
    o.run(steps=CHAIN3D_STEPS, time_step=fx.dt, time_step_output=fx.dt)
    return o


def check_chain3d(o):
    ref = np.load(GOLDEN)
    assert max(common.max_err_deg(np.asarray(o.elements.lon), np.asarray(o.elements.lat), ref['chain3d__lon'], ref['chain3d__lat'])) < 5e-8
    assert np.abs(np.asarray(o.elements.z, dtype=np.float64) - ref['chain3d__z']).max() <= 1e-5



# ---- readers that follow each other in time: a step that straddles the hand-over samples both (one get_environment per stage) -----
HANDOVER_CASES = [('two_readers', 'runge-kutta'), ('two_readers', 'runge-kutta4'), ('late_reader', 'runge-kutta'), ('late_reader', 'runge-kutta4')]
HANDOVER_N, HANDOVER_STEPS, HANDOVER_DT = 300, 8, 840          # 14-minute steps over hourly slabs


def handover_readers(fx, kind, make):
    t, cur = fx.times, common.CUR
    out = []
    if kind == 'two_readers':           # A: the first hour; B: from the first hour on, other values
        out.append(make(fx.grid_lon, fx.grid_lat, None, t[:2], {cur[0]: fx.u[:2], cur[1]: fx.v[:2]}, 'A'))
        out.append(make(fx.grid_lon, fx.grid_lat, None, t[1:], {cur[0]: (0.5 * fx.u[1:]).astype(np.float32),
                                                                 cur[1]: (-0.5 * fx.v[1:]).astype(np.float32)}, 'B'))
    else:                               # nothing before the first hour: the stages before it get the fallback
        out.append(make(fx.grid_lon, fx.grid_lat, None, t[1:], {cur[0]: fx.u[1:], cur[1]: fx.v[1:]}, 'B'))
    return out


def run_product_handover(kind, scheme, **model_kw):
    from opendrift_b200.models.oceandrift import OceanDrift
    from opendrift_b200.readers import reader_regular_grid
    fx = common.Fixture('rk4_2d')
    o = OceanDrift(loglevel=50, seed=0, **model_kw)
    o.add_reader(handover_readers(fx, kind, lambda lon, lat, z, t, f, name: reader_regular_grid.Reader(lon, lat, z, t, f, name=name)))
    o.set_config('general:use_auto_landmask', False)
    o.set_config('drift:advection_scheme', scheme)
    o.set_config('drift:vertical_advection', False)
    n = HANDOVER_N
    o.seed_elements(lon=fx.lon0[:n], lat=fx.lat0[:n], z=0.0, time=fx.start)
    o.run(steps=HANDOVER_STEPS, time_step=HANDOVER_DT, time_step_output=HANDOVER_DT)
    return o


def check_handover(o, kind, scheme):
    ref = np.load(GOLDEN)
    k = 'handover_%s_%s' % (kind, scheme)
    assert max(common.max_err_deg(np.asarray(o.elements.lon), np.asarray(o.elements.lat), ref[k + '__lon'], ref[k + '__lat'])) < 5e-8


# ---- Leeway: staggered release, backward runs with capsizing (leeway.py:430-494) -------------------------------------------------
LEEWAY_CASES = {
    'leeway_staggered': ({}, 1, 'interval', 7, 600),
    'leeway_backward': ({}, 2, 'end', 6, -600),
    'leeway_backward_capsizing': ({'processes:capsizing': True, 'capsizing:wind_threshold': 8.0, 'capsizing:wind_threshold_sigma': 5.0},
                                  1, 'end', 6, -600),
    # uncertainty of current and wind: one set of draws per step, made before capsizing / jibing draw theirs
    'leeway_uncertainty': ({'drift:current_uncertainty': 0.1, 'drift:current_uncertainty_uniform': 0.05, 'drift:wind_uncertainty': 1.5},
                           1, 'interval', 7, 600),
    'leeway_uncertainty_capsizing': ({'drift:wind_uncertainty': 2.0, 'processes:capsizing': True, 'capsizing:wind_threshold': 6.0,
                                      'capsizing:wind_threshold_sigma': 3.0}, 3, 'interval', 6, 600),
    'leeway_capsizing_staggered': ({'processes:capsizing': True, 'capsizing:wind_threshold': 6.0, 'capsizing:wind_threshold_sigma': 3.0},
                                   3, 'interval', 7, 600),
}
LEEWAY_N = 400


def leeway_seed(fx, case):
    cfg, object_type, when, steps, dt = LEEWAY_CASES[case]
    t = [fx.start, fx.start + timedelta(seconds=2400)] if when == 'interval' else fx.times[-1] - timedelta(seconds=600)
    kw = dict(lon=fx.lon0[:LEEWAY_N], lat=fx.lat0[:LEEWAY_N], time=t, object_type=object_type)
    if case == 'leeway_backward_capsizing':
        kw['capsized'] = 1                      # a backward run 'un-capsizes' (leeway.py:443-454)
    return cfg, kw, steps, dt


def run_product_leeway(fx, case, **model_kw):
    from opendrift_b200.models.leeway import Leeway
    from opendrift_b200.readers import reader_regular_grid
    cfg, seedkw, steps, dt = leeway_seed(fx, case)
    o = Leeway(loglevel=50, seed=0, **model_kw)
    o.add_reader([reader_regular_grid.Reader(fx.grid_lon, fx.grid_lat, None, fx.times, {common.CUR[0]: fx.u, common.CUR[1]: fx.v}, name='current'),
                  reader_regular_grid.Reader(fx.grid_lon, fx.grid_lat, None, fx.times, {'x_wind': fx.x_wind, 'y_wind': fx.y_wind}, name='wind')])
    o.set_config('general:use_auto_landmask', False)
    for k, v in cfg.items():
        o.set_config(k, v)
    o.seed_elements(**seedkw)
    o.run(steps=steps, time_step=dt, time_step_output=dt)
    return o


def check_leeway(o, case):
    ref = np.load(GOLDEN)
    g = lambda k: ref['%s__%s' % (case, k)]                      # noqa: E731
    assert np.array_equal(np.asarray(o.elements.ID, dtype=np.int64), g('id'))
    assert max(common.max_err_deg(np.asarray(o.elements.lon), np.asarray(o.elements.lat), g('lon'), g('lat'))) < 5e-8
    assert np.array_equal(np.asarray(o.elements.orientation, dtype=np.int64), g('orientation'))
    assert np.array_equal(np.asarray(o.elements.capsized, dtype=np.int64), g('capsized'))
    assert np.array_equal(np.asarray(o.elements.crosswind_slope, dtype=np.float64), g('crosswind_slope'))



# ---- OceanDrift option combinations the fixtures do not have (compared with the live reference through the model classes) ------
OD_N = 400


def od_cases():
    """name -> (fixture, readers, config, seed overrides, steps, dt).  readers: 'cur' (u, v, w), 'cur_k' (u, v, K), 'cur_short'
    (u, v on the first two slabs only), 'wind'."""
    full = common.Fixture('rk4_3d_full')
    mix = common.Fixture('rk4_3d_mixing')
    n = OD_N
    idx = np.arange(n)
    rk4, rk2 = {'drift:advection_scheme': 'runge-kutta4'}, {'drift:advection_scheme': 'runge-kutta'}
    mixing = {'drift:vertical_mixing': True, 'drift:vertical_advection': False}
    return {
        'reader_ends_mid_run': (full, ['cur_short'], {**rk4, 'drift:vertical_advection': False}, {}, 9, 600),
        'constant_wind_gridded_current': (full, ['cur'], {**rk2, 'environment:constant:x_wind': 7.0, 'environment:constant:y_wind': -3.0}, {'z': 0.0}, 5, 600),
        'gridded_wind_constant_current': (full, ['wind'], {**rk4, 'environment:constant:x_sea_water_velocity': 0.3,
                                                            'environment:constant:y_sea_water_velocity': 0.1}, {'z': 0.0}, 5, 600),
        'wind_drift_depth_2_and_z_profile': (full, ['cur', 'wind'], {'drift:wind_drift_depth': 2.0}, {'z': np.linspace(-3, 0, n)}, 5, 600),
        'diffusion_and_uniform_current_uncertainty': (full, ['cur', 'wind'], {**rk2, 'environment:constant:horizontal_diffusivity': 10.0,
                                                                              'drift:current_uncertainty_uniform': 0.1}, {}, 4, 600),
        'wdf_array_cdf_scalar': (full, ['cur', 'wind'], rk4, {'z': 0.0, 'wind_drift_factor': np.linspace(0, 0.05, n).astype(np.float32),
                                                              'current_drift_factor': 0.8}, 5, 600),
        'relative_wind': (full, ['cur', 'wind'], {**rk2, 'drift:relative_wind': True}, {'z': 0.0}, 4, 600),
        'relative_wind_with_wind_uncertainty': (full, ['cur', 'wind'], {'drift:relative_wind': True, 'drift:wind_uncertainty': 2.0}, {'z': 0.0}, 4, 600),
        'vertical_advection_at_surface': (full, ['cur'], {'drift:vertical_advection_at_surface': True}, {'z': 0.0}, 5, 600),
        'all_noise_and_diffusion_rk4': (full, ['cur', 'wind'], {**rk4, 'drift:current_uncertainty': 0.2, 'drift:wind_uncertainty': 1.0,
                                                                'environment:constant:horizontal_diffusivity': 3.0}, {}, 4, 600),
        'positive_z_seeds': (full, ['cur', 'wind'], {}, {'z': np.linspace(-1, 0.5, n)}, 3, 600),
        'mixing_terminal_velocity': (mix, ['cur_k'], mixing, {'terminal_velocity': 0.002}, 4, 600),
        'mixing_sinking_array': (mix, ['cur_k'], mixing, {'terminal_velocity': np.linspace(-0.003, 0.0, n).astype(np.float32)}, 4, 600),
        'mixing_at_surface': (mix, ['cur_k'], {**mixing, 'drift:vertical_mixing_at_surface': True}, {'z': np.where(idx % 3 == 0, 0.0, mix.z0[:n])}, 4, 600),
        'mixing_shallow_sea_floor': (mix, ['cur_k'], {**mixing, 'environment:constant:sea_floor_depth_below_sea_level': 30.0},
                                     {'z': np.clip(mix.z0[:n], -29, 0)}, 4, 600),
        'mixing_dt_45': (mix, ['cur_k'], {**mixing, 'vertical_mixing:timestep': 45.0}, {}, 4, 600),
        'mixing_constant_model': (mix, ['cur'], {**mixing, 'vertical_mixing:diffusivitymodel': 'constant',
                                                 'environment:fallback:ocean_vertical_diffusivity': 0.01}, {}, 4, 600),
        'mixing_backward': (mix, ['cur_k'], mixing, {'time': mix.times[-1]}, 4, -600),
        # the step's uncertainty is drawn for the elements active at the top of the loop, before this step's deactivations
        # (basemodel/__init__.py:2238-2262): the generator stays in step with the reference while elements leave
        'uncertainty_with_deactivation': (full, ['cur', 'wind'], {**rk4, 'drift:current_uncertainty': 0.1, 'drift:wind_uncertainty': 1.0,
                                                                  'drift:deactivate_east_of': float(np.percentile(full.lon0[:n], 70)),
                                                                  }, {'z': 0.0}, 6, 600),
        # the wind uncertainty is added to the fallback wind (0) too, and the elements drift with it
        'wind_uncertainty_without_wind_reader': (full, ['cur'], {'drift:wind_uncertainty': 0.5, 'drift:vertical_advection': False}, {'z': 0.0}, 4, 600),
        # analytical diffusivity model + uncertainty: the mixing launch reads the environment, whose draws must not be made twice
        # no vertical mixing: update() moves the depth with the terminal velocity instead (vertical_buoyancy, oceandrift.py:201-205,
        # 352-367), before vertical advection; float32 / float64 combinations of z and terminal_velocity
        'buoyancy_rising_scalar_tv': (full, ['cur', 'wind'], rk2, {'terminal_velocity': 0.003}, 5, 600),
        'buoyancy_array_tv_scalar_z': (full, ['cur'], rk4, {'z': -5.0, 'terminal_velocity': np.linspace(-0.004, 0.004, n).astype(np.float32)}, 5, 600),
        'buoyancy_array_tv_no_w': (full, ['cur', 'wind'], {'drift:vertical_advection': False},
                                   {'terminal_velocity': np.linspace(-0.01, 0.01, n).astype(np.float32)}, 4, 600),
        # a reader for the sea floor: interact_with_seafloor at the top of the loop and at the end of vertical_buoyancy
        # (basemodel/__init__.py:748-783)
        'seafloor_reader_lift': (full, ['cur', 'floor'], rk2, {'terminal_velocity': -0.01}, 5, 600),
        'seafloor_reader_deactivate': (full, ['cur', 'floor'], {**rk2, 'general:seafloor_action': 'deactivate'}, {'terminal_velocity': -0.01}, 5, 600),
        'seafloor_reader_with_mixing': (mix, ['cur_k', 'floor'], mixing, {'terminal_velocity': -0.002}, 4, 600),
        'mixing_constant_model_with_uncertainty': (mix, ['cur'], {**mixing, **rk4, 'vertical_mixing:diffusivitymodel': 'constant',
                                                                  'environment:fallback:ocean_vertical_diffusivity': 0.01,
                                                                  'drift:current_uncertainty': 0.1}, {}, 4, 600),
    }


def od_readers(fx, which, make):
    out = []
    cur = {common.CUR[0]: fx.u, common.CUR[1]: fx.v}
    if 'cur' in which:
        f = dict(cur)
        if fx.w is not None:
            f['upward_sea_water_velocity'] = fx.w
        out.append(make(fx.grid_lon, fx.grid_lat, fx.grid_z, fx.times, f, 'cur'))
    if 'cur_k' in which:
        out.append(make(fx.grid_lon, fx.grid_lat, fx.grid_z, fx.times, dict(cur, ocean_vertical_diffusivity=fx.kdiff), 'cur'))
    if 'cur_short' in which:
        out.append(make(fx.grid_lon, fx.grid_lat, fx.grid_z, fx.times[:2], {k: v[:2] for k, v in cur.items()}, 'cur'))
    if 'wind' in which:
        out.append(make(fx.wind_lon, fx.wind_lat, None, fx.times, {'x_wind': fx.x_wind, 'y_wind': fx.y_wind}, 'wind'))
    if 'floor' in which:        # a sloping sea floor, 8 m in the west to 60 m in the east, shallower than the deepest seeds
        depth = np.broadcast_to(np.linspace(8, 60, len(fx.grid_lon), dtype=np.float32), (len(fx.times), len(fx.grid_lat), len(fx.grid_lon)))
        out.append(make(fx.grid_lon, fx.grid_lat, None, fx.times, {'sea_floor_depth_below_sea_level': np.ascontiguousarray(depth)}, 'floor'))
    return out


def od_seed(fx, over):
    kw = dict(lon=fx.lon0[:OD_N], lat=fx.lat0[:OD_N], z=fx.z0[:OD_N], time=fx.start)
    kw.update(over)
    return kw


def run_product_od(case, **model_kw):
    from opendrift_b200.models.oceandrift import OceanDrift
    from opendrift_b200.readers import reader_regular_grid
    fx, which, cfg, over, steps, dt = od_cases()[case]
    o = OceanDrift(loglevel=50, seed=0, **model_kw)
    for r in od_readers(fx, which, lambda lon, lat, z, t, f, name: reader_regular_grid.Reader(lon, lat, z, t, f, name=name)):
        o.add_reader(r)
    o.set_config('general:use_auto_landmask', False)
    for k, v in cfg.items():
        o.set_config(k, v)
    o.seed_elements(**od_seed(fx, over))
    o.run(steps=steps, time_step=dt, time_step_output=dt)
    return o


def check_od(o, case):
    ref = np.load(GOLDEN)
    g = lambda k: ref['od_%s__%s' % (case, k)]                   # noqa: E731
    assert np.array_equal(np.asarray(o.elements.ID, dtype=np.int64), g('id'))
    assert max(common.max_err_deg(np.asarray(o.elements.lon), np.asarray(o.elements.lat), g('lon'), g('lat'))) < 5e-8
    assert np.abs(np.asarray(o.elements.z, dtype=np.float64) - g('z')).max() <= (1e-9 if case.startswith('mixing') else 1e-5)


# ---- subclass recipes on the helpers: drift in sea ice (OpenOil's advect_oil, openoil.py:1179-1216) and the combined swell / wind-sea
#      Stokes profile (physics_methods.py:418-455) -- the same subclass body runs on the reference's OceanDrift and on the product's ----
WAVE_VARS = {'sea_surface_swell_wave_to_direction': 250.0, 'sea_surface_swell_wave_peak_period_from_variance_spectral_density': 11.0,
             'sea_surface_swell_wave_significant_height': 1.4, 'sea_surface_wind_wave_to_direction': 40.0,
             'sea_surface_wind_wave_mean_period': 4.5, 'sea_surface_wind_wave_significant_height': 1.1}


def ice_model_class(Base, ice_velocity):
    extra = {'sea_ice_area_fraction': {'fallback': 0}}
    if ice_velocity:
        extra.update({'sea_ice_x_velocity': {'fallback': 0}, 'sea_ice_y_velocity': {'fallback': 0}})

    class IceDrift(Base):
        required_variables = dict(Base.required_variables, **extra)

        def update(self):
            A = self.environment.sea_ice_area_fraction
            k_ice = (A - 0.3) / (0.8 - 0.3)             # Nordam et al. (2019): drift with the ice above 80 % cover, with the water below 30 %
            k_ice[A < 0.3] = 0
            k_ice[A > 0.8] = 1
            factor_stokes = (0.7 - A) / 0.7             # Arneborg (2017): waves are damped by the ice
            factor_stokes[A > 0.7] = 0
            self.advect_ocean_current(factor=1 - k_ice)
            self.advect_wind(factor=1 - k_ice)
            self.stokes_drift(factor_stokes)
            self.advect_with_sea_ice(factor=k_ice)
            self.vertical_advection()
    return IceDrift


def wave_model_class(Base):
    class WaveDrift(Base):
        required_variables = dict(Base.required_variables, **{v: {'fallback': 0} for v in WAVE_VARS})
    return WaveDrift


SUBCLASS_CASES = {
    'ice_with_ice_velocity': ('ice', True, {'drift:advection_scheme': 'runge-kutta'}, {'z': 0.0}),
    'ice_rule_of_thumb': ('ice', False, {'drift:advection_scheme': 'euler', 'drift:stokes_drift_profile': 'exponential'}, {}),
    'windsea_swell_z_float64': ('wave', None, {'drift:stokes_drift_profile': 'windsea_swell', 'drift:vertical_advection': False}, {'z': -0.5}),
    'windsea_swell_z_float32': ('wave', None, {'drift:stokes_drift_profile': 'windsea_swell', 'drift:advection_scheme': 'runge-kutta4',
                                              'drift:vertical_advection': False}, {}),
}
SUB_N, SUB_STEPS = 400, 4


def subclass_readers(fx, kind, ice_velocity, make, constant):
    st = common.Fixture('rk4_3d_stokes_phillips')
    nt, ny, nx = len(fx.times), len(fx.grid_lat), len(fx.grid_lon)
    out = [make(fx.grid_lon, fx.grid_lat, fx.grid_z, fx.times, {common.CUR[0]: fx.u, common.CUR[1]: fx.v, 'upward_sea_water_velocity': fx.w}, 'cur'),
           make(fx.wind_lon, fx.wind_lat, None, fx.times, {'x_wind': fx.x_wind, 'y_wind': fx.y_wind}, 'wind'),
           make(st.grid_lon, st.grid_lat, None, st.times, dict(st.stokes), 'waves')]
    if kind == 'ice':
        X, Y = np.meshgrid(np.linspace(0, 1, nx), np.linspace(0, 1, ny))
        A = np.clip(0.5 + 0.7 * np.sin(2 * np.pi * X) * np.cos(np.pi * Y), 0, 1).astype(np.float32)
        f = {'sea_ice_area_fraction': np.ascontiguousarray(np.broadcast_to(A, (nt, ny, nx)))}
        if ice_velocity:
            f['sea_ice_x_velocity'] = np.ascontiguousarray(np.broadcast_to((0.2 * np.cos(np.pi * Y)).astype(np.float32), (nt, ny, nx)))
            f['sea_ice_y_velocity'] = np.ascontiguousarray(np.broadcast_to((-0.1 * np.sin(np.pi * X)).astype(np.float32), (nt, ny, nx)))
        out.append(make(fx.grid_lon, fx.grid_lat, None, fx.times, f, 'ice'))
    else:
        out.append(constant(dict(WAVE_VARS)))
    return out


def run_subclass_case(case, Base, make, constant, **model_kw):
    kind, ice_velocity, cfg, over = SUBCLASS_CASES[case]
    fx = common.Fixture('rk4_3d_full')
    Model = ice_model_class(Base, ice_velocity) if kind == 'ice' else wave_model_class(Base)
    o = Model(loglevel=50, seed=0, **model_kw)
    for r in subclass_readers(fx, kind, ice_velocity, make, constant):
        o.add_reader(r)
    for k, v in {'general:use_auto_landmask': False, 'general:coastline_action': 'none', **cfg}.items():
        o.set_config(k, v)
    if 'environment:constant:land_binary_mask' in getattr(o, '_config', {}):
        o.set_config('environment:constant:land_binary_mask', 0)
    kw = dict(lon=fx.lon0[:SUB_N], lat=fx.lat0[:SUB_N], z=np.clip(fx.z0[:SUB_N], -3, 0), time=fx.start)
    kw.update(over)
    o.seed_elements(**kw)
    o.run(steps=SUB_STEPS, time_step=600, time_step_output=600)
    return o


def run_product_subclass(case, **model_kw):
    from opendrift_b200.models.oceandrift import OceanDrift
    from opendrift_b200.readers import reader_regular_grid, reader_constant
    return run_subclass_case(case, OceanDrift, lambda lon, lat, z, t, f, name: reader_regular_grid.Reader(lon, lat, z, t, f, name=name),
                             lambda m: reader_constant.Reader(m), **model_kw)


def check_subclass(o, case):
    ref = np.load(GOLDEN)
    g = lambda k: ref['sub_%s__%s' % (case, k)]                  # noqa: E731
    assert np.array_equal(np.asarray(o.elements.ID, dtype=np.int64), g('id'))
    assert max(common.max_err_deg(np.asarray(o.elements.lon), np.asarray(o.elements.lat), g('lon'), g('lat'))) < 5e-8
    assert np.abs(np.asarray(o.elements.z, dtype=np.float64) - g('z')).max() <= 1e-5
    return float(np.abs(np.asarray(o.elements.lon) - g('lon0')).max())


# ---- readers that hand out sub-blocks around the elements (reader_netCDF_CF_generic.py:404-626; structured.py:243-318) -----------
SUBBLOCK_CASES = {
    # the elements sit in a corner of the grid: an 11 x 11 block of the 36 x 40 grid serves the whole run
    'corner_cluster': dict(spread=0.15, steps=10, dt=600, cfg={'drift:advection_scheme': 'runge-kutta4', 'drift:max_speed': 0.6}),
    # long steps and a small anticipated speed: the elements outrun the buffer and the block is replaced under way
    'outrun_the_buffer': dict(spread=0.08, steps=10, dt=1800, cfg={'drift:advection_scheme': 'runge-kutta', 'drift:max_speed': 0.05}),
}
SUBBLOCK_N = 300


def subblock_setup(case):
    c = SUBBLOCK_CASES[case]
    fx = common.Fixture('rk4_3d_full')
    n = SUBBLOCK_N
    lon0 = (fx.grid_lon[3] + (fx.lon0[:n] - fx.lon0[:n].min()) * c['spread']).astype(np.float32)
    lat0 = (fx.grid_lat[3] + (fx.lat0[:n] - fx.lat0[:n].min()) * c['spread']).astype(np.float32)
    fields = {common.CUR[0]: fx.u, common.CUR[1]: fx.v, 'upward_sea_water_velocity': fx.w}
    return fx, c, lon0, lat0, fx.z0[:n], fields


def run_product_subblock(case, subblocks=True, **model_kw):
    from opendrift_b200.models.oceandrift import OceanDrift
    from opendrift_b200.readers import reader_regular_grid
    fx, c, lon0, lat0, z0, fields = subblock_setup(case)
    o = OceanDrift(loglevel=50, seed=0, **model_kw)
    rd = reader_regular_grid.Reader(fx.grid_lon, fx.grid_lat, fx.grid_z, fx.times, fields, name='cur', subblocks=subblocks)
    o.add_reader(rd)
    o.set_config('general:use_auto_landmask', False)
    for k, v in c['cfg'].items():
        o.set_config(k, v)
    o.seed_elements(lon=lon0, lat=lat0, z=z0, time=fx.start)
    o.run(steps=c['steps'], time_step=c['dt'], time_step_output=c['dt'])
    return o, rd


def run_product_rewindow(subblocks, **model_kw):
    """A day-long drift (the fixture's four slabs repeated): the elements leave the first block and the reader is asked for new
    ones under way."""
    from datetime import timedelta as _td
    from opendrift_b200.models.oceandrift import OceanDrift
    from opendrift_b200.readers import reader_regular_grid
    fx, c, lon0, lat0, z0, fields = subblock_setup('corner_cluster')
    nt = fx.u.shape[0]
    times = [fx.start + _td(hours=k) for k in range(26)]
    per = {v: (lambda ti, a=a: a[ti % nt]) for v, a in fields.items()}
    o = OceanDrift(loglevel=50, seed=0, **model_kw)
    rd = reader_regular_grid.Reader(fx.grid_lon, fx.grid_lat, fx.grid_z, times, per, name='cur', subblocks=subblocks)
    o.add_reader(rd)
    o.set_config('general:use_auto_landmask', False)
    o.set_config('drift:advection_scheme', 'runge-kutta')
    o.set_config('drift:max_speed', 1.0)           # buffer: ceil(1 m/s * 3600 s / 5.5 km) + 2 = 3 cells around the elements
    o.seed_elements(lon=lon0, lat=lat0, z=z0, time=fx.start)
    o.run(steps=24, time_step=3600, time_step_output=3600 * 24)
    return o, rd


def check_subblock(o, rd, case):
    ref = np.load(GOLDEN)
    g = lambda k: ref['subblock_%s__%s' % (case, k)]             # noqa: E731
    e = max(common.max_err_deg(np.asarray(o.elements.lon), np.asarray(o.elements.lat), g('lon'), g('lat')))
    dz = float(np.abs(np.asarray(o.elements.z, dtype=np.float64) - g('z')).max())
    grp = rd.group_of(common.CUR[0])[0]
    return e, dz, (grp.desc.ny, grp.desc.nx), rd.windows_set


# ---- subclasses that override the per-iteration hooks of the mixing loop (oceandrift.py:369-379, 515-564) -----------------------
def hook_model_class(Base, which):
    class Hooked(Base):
        if which in ('same_stick', 'all'):
            def surface_stick(self):                       # the stock behaviour, spelled out by the subclass
                z = self.elements.z
                z[z > 0] = 0
                self.elements.z = z
        if which in ('wave_mixing', 'all'):
            def surface_wave_mixing(self, time_step_seconds):      # surfaced elements are pushed down by breaking waves (draws!)
                z = self.elements.z
                surf = np.where(z == 0)[0]
                if len(surf):
                    z[surf] = -0.5 * np.random.random(len(surf)) * time_step_seconds / 60.0
                    self.elements.z = z
        if which == 'all':
            def update_terminal_velocity(self, Tprofiles=None, Sprofiles=None, z_index=None):
                self.elements.terminal_velocity = 0.002 * np.exp(np.asarray(self.elements.z, dtype=np.float64) / 20.0)

            def prepare_vertical_mixing(self):
                self.prepared = getattr(self, 'prepared', 0) + 1
    return Hooked


HOOK_CASES = ['same_stick', 'wave_mixing', 'all']


def run_hook_case(which, Base, make, **model_kw):
    fx = common.Fixture('rk4_3d_mixing')
    o = hook_model_class(Base, which)(loglevel=50, seed=0, **model_kw)
    o.add_reader(make(fx.grid_lon, fx.grid_lat, fx.grid_z, fx.times, {common.CUR[0]: fx.u, common.CUR[1]: fx.v, 'ocean_vertical_diffusivity': fx.kdiff}, 'cur'))
    for k, v in {'general:use_auto_landmask': False, 'general:coastline_action': 'none', 'drift:vertical_mixing': True,
                 'drift:vertical_advection': False, 'vertical_mixing:timestep': 60.0, 'drift:advection_scheme': 'runge-kutta'}.items():
        o.set_config(k, v)
    if 'environment:constant:land_binary_mask' in getattr(o, '_config', {}):
        o.set_config('environment:constant:land_binary_mask', 0)
    n = 300
    z0 = np.where(np.arange(n) % 4 == 0, 0.0, fx.z0[:n]).astype(np.float32)
    o.seed_elements(lon=fx.lon0[:n], lat=fx.lat0[:n], z=z0, time=fx.start, terminal_velocity=0.001)
    o.run(steps=3, time_step=600, time_step_output=600)
    return o


def run_product_hooks(which, **model_kw):
    from opendrift_b200.models.oceandrift import OceanDrift
    from opendrift_b200.readers import reader_regular_grid
    return run_hook_case(which, OceanDrift, lambda lon, lat, z, t, f, name: reader_regular_grid.Reader(lon, lat, z, t, f, name=name), **model_kw)


def check_hooks(o, which):
    ref = np.load(GOLDEN)
    g = lambda k: ref['hook_%s__%s' % (which, k)]               # noqa: E731
    assert max(common.max_err_deg(np.asarray(o.elements.lon), np.asarray(o.elements.lat), g('lon'), g('lat'))) < 5e-8
    return float(np.abs(np.asarray(o.elements.z, dtype=np.float64) - g('z')).max())


# ---- readers on a projected plane (spherical +proj=stere): lonlat2xy before the index arithmetic, vector pairs rotated to
#      east / north afterwards (readers/basereader/variables.py:129-143, 59-109, 799-837) -----------------------------------------
PROJ_POLAR = '+proj=stere +lat_0=90 +lon_0=10 +lat_ts=60 +R=6371000 +units=m +no_defs'
PROJ_OBLIQUE = '+proj=stere +lat_0=58 +lon_0=3 +k_0=0.9996 +x_0=50000 +y_0=-20000 +R=6370997 +units=m +no_defs'
PROJ_CASES = {
    'stere_polar_rk4_3d_w': dict(proj4=PROJ_POLAR, model='OceanDrift', readers=('cur3d',), steps=8, dt=600,
                                 cfg={'drift:advection_scheme': 'runge-kutta4'}),
    'stere_oblique_rk2_wind': dict(proj4=PROJ_OBLIQUE, model='OceanDrift', readers=('cur2d', 'wind'), steps=6, dt=900,
                                   cfg={'drift:advection_scheme': 'runge-kutta', 'drift:vertical_advection': False}, seed={'z': 0.0}),
    'stere_polar_mixing': dict(proj4=PROJ_POLAR, model='OceanDrift', readers=('cur3d_k',), steps=3, dt=600,
                               cfg={'drift:vertical_mixing': True, 'drift:vertical_advection': False, 'vertical_mixing:timestep': 60.0}),
    'stere_polar_leeway': dict(proj4=PROJ_POLAR, model='Leeway', readers=('cur2d', 'wind'), steps=6, dt=600, cfg={}),
}
PROJ_N = 400


def proj_setup(case):
    from oracle.proj_stere import Stere
    c = PROJ_CASES[case]
    P = Stere(c['proj4'])
    cx, cy = P.forward(np.array([4.0]), np.array([60.0]))
    nx, ny, nz, nt = 44, 38, 6, 4
    xs = (cx[0] + 4000.0 * (np.arange(nx) - nx // 2)).astype(np.float32)
    ys = (cy[0] + 4000.0 * (np.arange(ny) - ny // 2)).astype(np.float32)
    zs = -10.0 * np.arange(nz)
    times = common.syn.slab_times(nt)
    X, Y = np.meshgrid(np.linspace(0, 1, nx), np.linspace(0, 1, ny))
    lay = lambda fn: np.stack([[fn(k, z) for z in zs] for k in range(nt)]).astype(np.float32)      # noqa: E731
    flat = lambda fn: np.stack([fn(k) for k in range(nt)]).astype(np.float32)                      # noqa: E731
    U3 = lay(lambda k, z: 0.5 * np.sin(2 * np.pi * X + 0.3 * k) * np.cos(np.pi * Y) * np.exp(z / 40))
    V3 = lay(lambda k, z: 0.4 * np.cos(2 * np.pi * Y + 0.2 * k) * np.sin(np.pi * X) * np.exp(z / 40))
    fields = {
        'cur3d': {common.CUR[0]: U3, common.CUR[1]: V3,
                  'upward_sea_water_velocity': lay(lambda k, z: 1e-3 * np.sin(np.pi * X) * np.sin(np.pi * Y) * np.sin(np.pi * z / zs.min()))},
        'cur3d_k': {common.CUR[0]: U3, common.CUR[1]: V3,
                    'ocean_vertical_diffusivity': lay(lambda k, z: 0.01 * np.exp(z / 20) * (1 + 0.5 * np.sin(np.pi * X) * np.sin(np.pi * Y)))},
        'cur2d': {common.CUR[0]: U3[:, 0].copy(), common.CUR[1]: V3[:, 0].copy()},
        'wind': {'x_wind': flat(lambda k: 8 * np.cos(0.4 * k) * (1 + 0.3 * np.sin(np.pi * X))),
                 'y_wind': flat(lambda k: 8 * np.sin(0.4 * k) * (1 + 0.3 * np.cos(np.pi * Y)))},
    }
    rng = np.random.default_rng(3)
    px, py = rng.uniform(xs[3], xs[-4], PROJ_N), rng.uniform(ys[3], ys[-4], PROJ_N)
    lon0, lat0 = P.inverse(px, py)
    z0 = rng.uniform(-55, 0, PROJ_N).astype(np.float32)
    return c, xs, ys, zs, times, fields, lon0.astype(np.float32), lat0.astype(np.float32), z0


def run_proj_case(case, classes, make, **model_kw):
    c, xs, ys, zs, times, fields, lon0, lat0, z0 = proj_setup(case)
    o = classes[c['model']](loglevel=50, seed=0, **model_kw)
    for nm in c['readers']:
        three_d = nm.startswith('cur3d')
        o.add_reader(make(xs, ys, zs if three_d else None, times, fields[nm], nm, c['proj4']))
    for k, v in {'general:use_auto_landmask': False, 'general:coastline_action': 'none', **c['cfg']}.items():
        o.set_config(k, v)
    if 'environment:constant:land_binary_mask' in getattr(o, '_config', {}):
        o.set_config('environment:constant:land_binary_mask', 0)
    if c['model'] == 'Leeway':
        o.seed_elements(lon=lon0, lat=lat0, time=times[0], object_type=1)
    else:
        kw = dict(lon=lon0, lat=lat0, z=z0, time=times[0])
        kw.update(c.get('seed', {}))
        o.seed_elements(**kw)
    o.run(steps=c['steps'], time_step=c['dt'], time_step_output=c['dt'])
    return o


def run_product_proj(case, **model_kw):
    from opendrift_b200.models.oceandrift import OceanDrift
    from opendrift_b200.models.leeway import Leeway
    from opendrift_b200.readers import reader_regular_grid
    return run_proj_case(case, {'OceanDrift': OceanDrift, 'Leeway': Leeway},
                         lambda x, y, z, t, f, name, proj4: reader_regular_grid.Reader(x, y, z, t, f, name=name, proj4=proj4), **model_kw)


def check_proj(o, case):
    ref = np.load(GOLDEN)
    g = lambda k: ref['proj_%s__%s' % (case, k)]                 # noqa: E731
    e = max(common.max_err_deg(np.asarray(o.elements.lon), np.asarray(o.elements.lat), g('lon'), g('lat')))
    dz = float(np.abs(np.asarray(o.elements.z, dtype=np.float64) - g('z')).max())
    moved = float(np.abs(g('lon') - g('lon0')).max())
    return e, dz, moved


if __name__ == '__main__':
    from oracle import refrun
    fx = common.Fixture('rk4_3d')
    out = {}
    refrun.setup()
    from opendrift.models.oceandrift import OceanDrift as _RefODp
    from opendrift.models.leeway import Leeway as _RefLWp
    for case in PROJ_CASES:
        ro = run_proj_case(case, {'OceanDrift': _RefODp, 'Leeway': _RefLWp},
                           lambda x, y, z, t, f, name, proj4: refrun.make_grid_reader(x, y, z, t, f, name=name, proj4=proj4), logfile='/tmp/od_bk.log')
        lon0 = proj_setup(case)[6]
        out.update({'proj_%s__lon' % case: np.asarray(ro.elements.lon, dtype=np.float64), 'proj_%s__lat' % case: np.asarray(ro.elements.lat, dtype=np.float64),
                    'proj_%s__z' % case: np.asarray(ro.elements.z, dtype=np.float64), 'proj_%s__lon0' % case: lon0.astype(np.float64)})
        print('proj', case, len(ro.elements.lon), 'moved', float(np.abs(np.asarray(ro.elements.lon) - lon0).max()))
    from opendrift.models.oceandrift import OceanDrift as _RefOD0
    for which in HOOK_CASES:
        ro = run_hook_case(which, _RefOD0, lambda lon, lat, z, t, f, name: refrun.make_grid_reader(lon, lat, z, t, f, name=name), logfile='/tmp/od_bk.log')
        out.update({'hook_%s__lon' % which: np.asarray(ro.elements.lon, dtype=np.float64), 'hook_%s__lat' % which: np.asarray(ro.elements.lat, dtype=np.float64),
                    'hook_%s__z' % which: np.asarray(ro.elements.z, dtype=np.float64)})
        print('hooks', which, 'z range', float(np.min(ro.elements.z)), float(np.max(ro.elements.z)))
    for case in SUBBLOCK_CASES:
        sfx, c, lon0, lat0, z0, fields = subblock_setup(case)
        rd = refrun.make_grid_reader(sfx.grid_lon, sfx.grid_lat, sfx.grid_z, sfx.times, fields, 'cur', subblocks=True)
        ro = refrun.run_oceandrift([rd], lon0, lat0, z0, sfx.start, c['dt'], c['steps'], config=c['cfg'])
        out.update({'subblock_%s__lon' % case: np.asarray(ro.elements.lon, dtype=np.float64), 'subblock_%s__lat' % case: np.asarray(ro.elements.lat, dtype=np.float64),
                    'subblock_%s__z' % case: np.asarray(ro.elements.z, dtype=np.float64)})
        print('subblock', case, 'blocks', rd.blocks_served, sorted(set(rd.block_shapes)), 'buffer', rd.buffer)
    from opendrift.models.oceandrift import OceanDrift as _RefOD
    from opendrift.readers import reader_constant as _ref_constant
    for case in SUBCLASS_CASES:
        ro = run_subclass_case(case, _RefOD, lambda lon, lat, z, t, f, name: refrun.make_grid_reader(lon, lat, z, t, f, name=name),
                               lambda m: _ref_constant.Reader(m), logfile='/tmp/od_bk.log')
        sfx = common.Fixture('rk4_3d_full')
        out.update({'sub_%s__id' % case: np.asarray(ro.elements.ID, dtype=np.int64), 'sub_%s__lon' % case: np.asarray(ro.elements.lon, dtype=np.float64),
                    'sub_%s__lat' % case: np.asarray(ro.elements.lat, dtype=np.float64), 'sub_%s__z' % case: np.asarray(ro.elements.z, dtype=np.float64),
                    'sub_%s__lon0' % case: sfx.lon0[:SUB_N].astype(np.float64)})
        print('subclass', case, len(ro.elements.ID), 'moved', np.abs(np.asarray(ro.elements.lon) - sfx.lon0[:SUB_N]).max())
    for case, (cfx, which, cfg, over, steps, dt) in od_cases().items():
        rds = od_readers(cfx, which, lambda lon, lat, z, t, f, name: refrun.make_grid_reader(lon, lat, z, t, f, name=name))
        kw = od_seed(cfx, over)
        ro = refrun.run_oceandrift(rds, kw['lon'], kw['lat'], kw['z'], kw['time'], dt, steps, config=cfg,
                                   seed_kwargs={k: v for k, v in kw.items() if k not in ('lon', 'lat', 'z', 'time')}, seed=0)
        out.update({'od_%s__id' % case: np.asarray(ro.elements.ID, dtype=np.int64), 'od_%s__lon' % case: np.asarray(ro.elements.lon, dtype=np.float64),
                    'od_%s__lat' % case: np.asarray(ro.elements.lat, dtype=np.float64), 'od_%s__z' % case: np.asarray(ro.elements.z, dtype=np.float64)})
        print('od', case, len(ro.elements.ID))
    cfx = common.Fixture('rk4_3d_full')
    ro = refrun.run_oceandrift(chain3d_readers(cfx, lambda lon, lat, z, t, f, name: refrun.make_grid_reader(lon, lat, z, t, f, name=name)),
                               cfx.lon0[:CHAIN3D_N], cfx.lat0[:CHAIN3D_N], cfx.z0[:CHAIN3D_N], cfx.start, cfx.dt, CHAIN3D_STEPS,
                               config=CHAIN3D_CFG, seed=0)
    out.update({'chain3d__lon': np.asarray(ro.elements.lon, dtype=np.float64), 'chain3d__lat': np.asarray(ro.elements.lat, dtype=np.float64),
                'chain3d__z': np.asarray(ro.elements.z, dtype=np.float64)})
    print('chain3d', len(ro.elements.lon))
    hfx = common.Fixture('rk4_2d')
    for kind, scheme in HANDOVER_CASES:
        ro = refrun.run_oceandrift(handover_readers(hfx, kind, lambda lon, lat, z, t, f, name: refrun.make_grid_reader(lon, lat, z, t, f, name=name)),
                                   hfx.lon0[:HANDOVER_N], hfx.lat0[:HANDOVER_N], 0.0, hfx.start, HANDOVER_DT, HANDOVER_STEPS,
                                   config={'drift:advection_scheme': scheme, 'drift:vertical_advection': False}, seed=0)
        k = 'handover_%s_%s' % (kind, scheme)
        out.update({k + '__lon': np.asarray(ro.elements.lon, dtype=np.float64), k + '__lat': np.asarray(ro.elements.lat, dtype=np.float64)})
        print('handover', kind, scheme)
    rfx, rcases = run_cases()
    for case, (seedkw, runkw, cfg) in rcases.items():
        from opendrift.models.oceandrift import OceanDrift as RefOceanDrift
        ro = RefOceanDrift(loglevel=50, logfile='/tmp/od_bk.log', seed=0)
        ro.add_reader(refrun.make_grid_reader(rfx.grid_lon, rfx.grid_lat, None, rfx.times, {common.CUR[0]: rfx.u, common.CUR[1]: rfx.v}))
        for k, v in {'general:use_auto_landmask': False, 'environment:constant:land_binary_mask': 0, 'general:coastline_action': 'none',
                     'drift:advection_scheme': 'runge-kutta', **cfg}.items():
            ro.set_config(k, v)
        ro.seed_elements(**seedkw)
        ro.run(**runkw)
        out.update({'run_%s__id' % case: np.asarray(ro.elements.ID, dtype=np.int64), 'run_%s__lon' % case: np.asarray(ro.elements.lon, dtype=np.float64),
                    'run_%s__lat' % case: np.asarray(ro.elements.lat, dtype=np.float64), 'run_%s__steps' % case: np.int64(ro.steps_calculation),
                    'run_%s__elapsed' % case: np.float64((ro.time - rfx.start).total_seconds())})
        print('run', case, ro.steps_calculation, len(ro.elements.ID))
    lf = common.LeewayFixture('leeway_piw1')
    for case in LEEWAY_CASES:
        cfg, seedkw, steps, dt = leeway_seed(lf, case)
        rds = [refrun.make_grid_reader(lf.grid_lon, lf.grid_lat, None, lf.times, {common.CUR[0]: lf.u, common.CUR[1]: lf.v}, name='current'),
               refrun.make_grid_reader(lf.grid_lon, lf.grid_lat, None, lf.times, {'x_wind': lf.x_wind, 'y_wind': lf.y_wind}, name='wind')]
        ro = refrun.run_oceandrift(rds, seedkw['lon'], seedkw['lat'], 0, seedkw['time'], dt, steps, config=cfg,
                                   seed_kwargs={k: v for k, v in seedkw.items() if k not in ('lon', 'lat', 'time')}, model='Leeway', seed=0)
        el = ro.elements
        out.update({case + '__id': np.asarray(el.ID, dtype=np.int64), case + '__lon': np.asarray(el.lon, dtype=np.float64),
                    case + '__lat': np.asarray(el.lat, dtype=np.float64), case + '__orientation': np.asarray(el.orientation, dtype=np.int64),
                    case + '__capsized': np.asarray(el.capsized, dtype=np.int64),
                    case + '__crosswind_slope': np.asarray(el.crosswind_slope, dtype=np.float64)})
        print(case, len(el.ID), 'capsized', int(np.sum(el.capsized)))
    f2 = common.Fixture('rk4_2d')
    a, b = multireader_fields(f2)
    for scheme in MULTI_SCHEMES:
        rds = [refrun.make_grid_reader(r['lon'], r['lat'], None, f2.times, {common.CUR[0]: r['u'], common.CUR[1]: r['v']}, name=nm)
               for nm, r in (('A', a), ('B', b))]
        ro = refrun.run_oceandrift(rds, f2.lon0, f2.lat0, f2.z0, f2.start, f2.dt, MULTI_STEPS,
                                   config={'drift:advection_scheme': scheme, 'drift:vertical_advection': False})
        out['multireader_%s__lon' % scheme] = np.asarray(ro.elements.lon, dtype=np.float64)
        out['multireader_%s__lat' % scheme] = np.asarray(ro.elements.lat, dtype=np.float64)
        print('multireader', scheme, len(ro.elements.lon))
    for case in CASES:
        t, cfg = case_setup(fx, case)
        rd = refrun.make_grid_reader(fx.grid_lon, fx.grid_lat, fx.grid_z, fx.times, {common.CUR[0]: fx.u, common.CUR[1]: fx.v})
        ro = refrun.run_oceandrift([rd], fx.lon0[:N], fx.lat0[:N], fx.z0[:N], t, case_dt(fx, case), STEPS, config=cfg)
        s = summary(ro, len(ro.elements_deactivated))
        for k, v in s.items():
            out['%s__%s' % (case, k)] = v
        print(case, 'active', len(s['id']), 'deactivated', len(s['d_id']))
    np.savez_compressed(GOLDEN, **out)
    print('wrote', GOLDEN)
