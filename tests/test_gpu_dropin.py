"""The reference-facing Python surface (OceanDrift / readers / Environment) on the GPU path, driven the
way the reference's own scripts and tests drive it, and compared with the UNMODIFIED reference's results
(tests/golden/ref_*.npz)."""
from datetime import timedelta

import numpy as np
import pytest

import common
from common import Fixture, fixtures

pytestmark = pytest.mark.gpu


def _model(fx, **cfg):
    from opendrift_b200.models.oceandrift import OceanDrift
    from opendrift_b200.readers import reader_regular_grid
    m = fx.meta
    o = OceanDrift(loglevel=50, seed=m['seed'])
    f3 = {common.CUR[0]: fx.u, common.CUR[1]: fx.v}
    w_own_grid = getattr(fx, 'w_lon', None) is not None          # random scenarios: upward velocity from a reader of its own
    if fx.w is not None and not w_own_grid:
        f3['upward_sea_water_velocity'] = fx.w
    if fx.kdiff is not None:
        f3['ocean_vertical_diffusivity'] = fx.kdiff
    o.add_reader(reader_regular_grid.Reader(fx.grid_lon, fx.grid_lat, fx.grid_z, fx.times, f3, name='current'))
    if fx.w is not None and w_own_grid:
        o.add_reader(reader_regular_grid.Reader(fx.w_lon, fx.w_lat, fx.w_z, fx.times, {'upward_sea_water_velocity': fx.w}, name='w'))
    if fx.x_wind is not None:
        o.add_reader(reader_regular_grid.Reader(fx.wind_lon, fx.wind_lat, None, fx.times,
                                                {'x_wind': fx.x_wind, 'y_wind': fx.y_wind}, name='wind'))
    if fx.stokes is not None:
        o.add_reader(reader_regular_grid.Reader(fx.grid_lon, fx.grid_lat, None, fx.times, dict(fx.stokes), name='waves'))
        o.set_config('drift:stokes_drift_profile', m['stokes'])
    o.set_config('general:use_auto_landmask', False)
    o.set_config('general:coastline_action', 'none')
    o.set_config('drift:advection_scheme', m['scheme'])
    o.set_config('drift:vertical_advection', bool(m['with_w']))
    if m['diffusivity']:
        o.set_config('environment:constant:horizontal_diffusivity', m['diffusivity'])
    if m.get('wind_drift_depth') is not None:
        o.set_config('drift:wind_drift_depth', m['wind_drift_depth'])
    for k, v in (m.get('noise') or {}).items():
        o.set_config('drift:current_uncertainty_uniform' if k == 'current_uniform' else 'drift:%s_uncertainty' % k, v)
    if m.get('truncate') is not None:
        o.set_config('drift:truncate_ocean_model_below_m', m['truncate'])
    if m.get('w_at_surface'):
        o.set_config('drift:vertical_advection_at_surface', True)
    if m.get('diffusivity_model') not in (None, 'environment_no_reader'):
        o.set_config('vertical_mixing:diffusivitymodel', m['diffusivity_model'])
    if m.get('background_diffusivity') is not None:
        o.set_config('vertical_mixing:background_diffusivity', m['background_diffusivity'])
    if m.get('mixing'):
        o.set_config('drift:vertical_mixing', True)
        o.set_config('vertical_mixing:timestep', m['dt_mix'])
    for k, v in cfg.items():
        o.set_config(k, v)
    kw = {}
    if fx.cdf is not None:
        kw['current_drift_factor'] = fx.cdf
    if getattr(fx, 'wdf_array', None) is not None:
        kw['wind_drift_factor'] = fx.wdf_array
    elif 'wdf' in m:
        kw['wind_drift_factor'] = m['wdf']
    o.seed_elements(lon=fx.lon0, lat=fx.lat0, z=fx.z0, time=fx.start, **kw)
    return o


@pytest.mark.parametrize('name', fixtures())
def test_oceandrift_run_matches_reference(name):
    fx = Fixture(name)
    o = _model(fx)
    o.run(steps=fx.steps, time_step=fx.dt, time_step_output=fx.dt)
    assert o.num_elements_active() == fx.n
    lon, lat, z = o.elements.lon, o.elements.lat, o.elements.z
    assert lon.dtype == np.float64                       # float64 after the first update, as in the reference
    e = common.max_err_deg(lon, lat, fx.lon, fx.lat)
    assert max(e) < 5e-8, e
    assert np.abs(z - fx.z).max() <= common.z_tolerance(fx.meta, exact=1e-9)
    assert z.dtype == fx.z.dtype
    # (in a backward run the reference flips the IDs: the element scheduled last is trajectory 0, basemodel/__init__.py:2056-2062)
    assert np.array_equal(o.elements.ID, np.arange(fx.n)[::-1] if fx.dt < 0 else np.arange(fx.n))
    assert len(o.history['time']) == fx.steps + 1


class HelperByHelper:
    """A model subclass that overrides update() like the reference's own subclasses do: the helpers then
    run as separate kernels and must give the same trajectories as the fused step."""


def test_overridden_update_uses_helpers_and_matches():
    from opendrift_b200.models.oceandrift import OceanDrift

    class MyDrift(OceanDrift):
        def update(self):
            self.advect_ocean_current()
            self.advect_wind()
            self.vertical_advection()

    class MyStokesDrift(OceanDrift):
        def update(self):
            self.advect_ocean_current()
            self.advect_wind()
            self.stokes_drift()

    class MyMixingDrift(OceanDrift):
        def update(self):
            self.advect_ocean_current()
            self.vertical_mixing()
            self.vertical_advection()

    for name in ('rk4_3d_full', 'euler_2d_wind', 'rk4_3d_cdf32', 'rk4_3d_mixing', 'euler_3d_mixing_w',
                 'rk4_3d_stokes_phillips', 'euler_3d_stokes_mono_nohs', 'rk4_3d_noise', 'rk2_3d_noise', 'rk2_3d_truncate',
                 'rk4_3d_truncate_wsurf'):
        fx = Fixture(name)
        o = _model(fx)
        o.__class__ = MyMixingDrift if fx.meta.get('mixing') else (MyStokesDrift if fx.meta.get('stokes') else MyDrift)
        o.run(steps=fx.steps, time_step=fx.dt)
        e = common.max_err_deg(o.elements.lon, o.elements.lat, fx.lon, fx.lat)
        assert max(e) < 5e-8, (name, e)
        assert np.abs(o.elements.z - fx.z).max() <= 1e-5


def test_subclass_touching_numpy_state_still_works():
    """Drop-in for model code that manipulates self.elements / self.environment as NumPy arrays."""
    from opendrift_b200.models.oceandrift import OceanDrift

    class Halver(OceanDrift):
        def update(self):
            u = self.environment.x_sea_water_velocity          # NumPy float32
            assert isinstance(u, np.ndarray) and u.dtype == np.float32
            self.update_positions(0.5 * u, 0.5 * self.environment.y_sea_water_velocity)
            self.elements.z = self.elements.z - np.float32(0.25)

    fx = Fixture('euler_3d')
    o = _model(fx)
    o.__class__ = Halver
    o.run(steps=3, time_step=fx.dt)
    assert np.allclose(o.elements.z, fx.z0 - 0.75, atol=1e-5)
    assert np.abs(o.elements.lon - fx.lon0).max() > 1e-4


def test_reader_get_variables_interpolated_and_environment():
    from oracle import advect_port as ap
    from opendrift_b200.readers import reader_regular_grid
    from opendrift_b200.errors import OutsideTemporalCoverageError
    fx = Fixture('rk4_3d_offgrid')
    r = reader_regular_grid.Reader(fx.grid_lon, fx.grid_lat, fx.grid_z, fx.times, {common.CUR[0]: fx.u, common.CUR[1]: fx.v})
    ref = ap.GridReader(fx.grid_lon, fx.grid_lat, fx.grid_z, fx.times, {common.CUR[0]: fx.u, common.CUR[1]: fx.v})
    t = fx.times[0] + timedelta(seconds=777)
    lon, lat, z = fx.lon0.astype(np.float64), fx.lat0.astype(np.float64), fx.z0
    env, prof = r.get_variables_interpolated(common.CUR, time=t, lon=lon, lat=lat, z=z)
    want = ap.reader_interpolate(ref, common.CUR, t, lon, lat, z)
    assert prof is None
    for v in common.CUR:
        got = np.ma.filled(env[v].astype(np.float32), np.nan)
        exp = want[v].astype(np.float32)
        assert np.array_equal(np.isnan(got), np.isnan(exp))
        assert np.array_equal(got[~np.isnan(got)], exp[~np.isnan(exp)])
        assert np.isnan(got).sum() > 50                      # uncovered points are masked, not filled
    with pytest.raises(OutsideTemporalCoverageError):
        r.get_variables_interpolated(common.CUR, time=fx.times[-1] + timedelta(days=1), lon=lon, lat=lat, z=z)


@pytest.mark.parametrize('name', common.leeway_fixtures())
def test_leeway_model_matches_reference(name):
    from opendrift_b200.models.leeway import Leeway
    from opendrift_b200.readers import reader_regular_grid
    fx = common.LeewayFixture(name)
    o = Leeway(loglevel=50, seed=fx.meta['seed'])
    o.add_reader([reader_regular_grid.Reader(fx.grid_lon, fx.grid_lat, None, fx.times, {common.CUR[0]: fx.u, common.CUR[1]: fx.v}, name='current'),
                  reader_regular_grid.Reader(fx.grid_lon, fx.grid_lat, None, fx.times, {'x_wind': fx.x_wind, 'y_wind': fx.y_wind}, name='wind')])
    o.set_config('general:use_auto_landmask', False)
    if fx.meta.get('capsizing'):
        o.set_config('processes:capsizing', True)
        o.set_config('capsizing:wind_threshold', fx.meta['capsizing'][0])
        o.set_config('capsizing:wind_threshold_sigma', fx.meta['capsizing'][1])
    o.seed_elements(lon=fx.lon0, lat=fx.lat0, time=fx.start, object_type=fx.meta['object_type'])
    o.run(steps=fx.steps, time_step=fx.dt, time_step_output=fx.dt)
    assert o.num_elements_active() == fx.n
    e = common.max_err_deg(o.elements.lon, o.elements.lat, fx.lon, fx.lat)
    assert max(e) < 5e-8, e
    assert np.array_equal(o.elements.orientation, fx.orientation)
    assert np.array_equal(o.elements.crosswind_slope, fx.crosswind_slope)
    if fx.capsized is not None:
        assert np.array_equal(np.asarray(o.elements.capsized, dtype=np.float64), fx.capsized)


def test_leeway_missing_forcing_deactivates():
    from opendrift_b200.models.leeway import Leeway
    from opendrift_b200.readers import reader_regular_grid
    fx = common.LeewayFixture('leeway_piw1')
    o = Leeway(loglevel=50, seed=1)
    o.add_reader([reader_regular_grid.Reader(fx.grid_lon, fx.grid_lat, None, fx.times, {common.CUR[0]: fx.u, common.CUR[1]: fx.v}),
                  reader_regular_grid.Reader(fx.grid_lon, fx.grid_lat, None, fx.times, {'x_wind': fx.x_wind, 'y_wind': fx.y_wind})])
    lon = fx.lon0.copy()
    lon[:100] += np.float32(5.0)                      # outside the readers' coverage: no fallback in Leeway
    o.seed_elements(lon=lon, lat=fx.lat0, time=fx.start, object_type=1)
    o.run(steps=3, time_step=600)
    assert o.num_elements_active() == fx.n - 100 and o.num_elements_deactivated() == 100
    assert 'missing_data' in o.status_categories
    assert np.array_equal(np.sort(o.elements_deactivated.ID), np.arange(100))


def test_seeding_radius_and_deactivation():
    from opendrift_b200.models.oceandrift import OceanDrift
    from opendrift_b200.readers import reader_regular_grid
    fx = Fixture('rk4_2d')
    o = OceanDrift(loglevel=50, seed=3)
    o.add_reader(reader_regular_grid.Reader(fx.grid_lon, fx.grid_lat, None, fx.times, {common.CUR[0]: fx.u, common.CUR[1]: fx.v}))
    clon, clat = float(fx.grid_lon.mean()), float(fx.grid_lat.mean())
    o.seed_elements(lon=clon, lat=clat, time=fx.times[0], number=5000, radius=2000.0)
    o.set_config('drift:advection_scheme', 'runge-kutta4')
    o.set_config('drift:max_age_seconds', 4 * 600.0)
    with pytest.raises(ValueError):
        o.set_config('drift:advection_scheme', 'leapfrog')
    o.run(steps=6, time_step=600)
    # gaussian radius: sigma of 2 km in each direction
    lon0, lat0 = np.array(o.history['lon'][0]), np.array(o.history['lat'][0])
    sx = np.std((lon0 - clon) * 111e3 * np.cos(np.radians(clat)))
    sy = np.std((lat0 - clat) * 111e3)
    assert 1800 < sx < 2200 and 1800 < sy < 2200
    # everybody retired at age 4 steps
    assert o.num_elements_active() == 0 and o.num_elements_deactivated() == 5000
    assert 'retired' in o.status_categories


def test_reference_known_answers_with_constant_environment():
    """The reference's own model-level known answers for this path, as its tests write them:
    tests/models/test_models.py:44-64 (test_wind_and_current_drift_factor) and
    tests/models/test_environment.py:30-41 (test_previous, coastline_action 'none')."""
    from datetime import datetime
    from opendrift_b200.models.oceandrift import OceanDrift
    lat, lon = 60, 4
    o = OceanDrift(loglevel=50)
    o.set_config('general:use_auto_landmask', False)
    o.set_config('environment:constant:land_binary_mask', 0)
    o.set_config('environment:constant:x_wind', 5)
    o.set_config('environment:constant:y_sea_water_velocity', 1)
    o.seed_elements(lon=lon, lat=lat, time=datetime.now(), wind_drift_factor=0, current_drift_factor=1)
    o.run(duration=timedelta(hours=2))
    o2 = OceanDrift(loglevel=50)
    o2.set_config('general:use_auto_landmask', False)
    o2.set_config('environment:constant:land_binary_mask', 0)
    o2.set_config('environment:constant:x_wind', 5)
    o2.set_config('environment:constant:y_sea_water_velocity', 1)
    o2.seed_elements(lon=lon, lat=lat, time=datetime.now(), wind_drift_factor=0.02, current_drift_factor=.3)
    o2.run(duration=timedelta(hours=2))
    assert abs(o.elements.lat[0] - (lat + 0.0646)) < 5e-4
    assert abs(o.elements.lon[0] - lon) < 5e-8
    assert abs(o2.elements.lat[0] - (lat + 0.0646 * .3)) < 5e-4
    assert abs(o2.elements.lon[0] - (lon + 0.0129)) < 5e-4

    o = OceanDrift(loglevel=50)
    o.set_config('general:coastline_action', 'none')
    o.set_config('drift:vertical_advection', False)
    o.set_config('environment:constant:land_binary_mask', 0)
    o.set_config('environment:constant:x_sea_water_velocity', 1)
    o.seed_elements(lon=3, lat=60, time=datetime.now())
    o.run(steps=1)
    assert o.elements.lon == pytest.approx(3.0645, .001)


@pytest.mark.parametrize('arith,tol', [('exact', 5e-8), ('fast', 1e-7)])
def test_arithmetic_config_selects_the_kernel_policy(arith, tol):
    """gpu:arithmetic = exact / fast run the same model through the other arithmetic policies of the step kernels."""
    fx = Fixture('rk4_3d_full')
    o = _model(fx, **{'gpu:arithmetic': arith})
    o.run(steps=fx.steps, time_step=fx.dt, time_step_output=fx.dt)
    e = common.max_err_deg(o.elements.lon, o.elements.lat, fx.lon, fx.lat)
    assert max(e) < tol, e
    from opendrift_b200 import _lib
    assert o.engine.math_mode == {'exact': _lib.OD_MATH_EXACT, 'fast': _lib.OD_MATH_FAST}[arith]
    o.engine.math_mode = _lib.OD_MATH_SERIES


def test_device_rng_for_diffusion_is_order_independent():
    """gpu:rng = philox: the random-walk draws are made on the device, keyed by element ID -- re-sorting the device arrays
    every step gives the same trajectories as never sorting, and the spread matches sqrt(2 D dt) per step."""
    from opendrift_b200.models.oceandrift import OceanDrift
    from opendrift_b200.readers import reader_regular_grid
    fx = Fixture('rk4_2d')
    n = 120000
    rng = np.random.default_rng(4)
    lon = rng.uniform(fx.grid_lon[4], fx.grid_lon[-5], n)
    lat = rng.uniform(fx.grid_lat[4], fx.grid_lat[-5], n)
    D, dt, steps = 20.0, 600, 3
    out = []
    for sort_every in (0, 1):
        o = OceanDrift(loglevel=50, seed=11)
        zero = {k: np.zeros_like(v) for k, v in {common.CUR[0]: fx.u, common.CUR[1]: fx.v}.items()}
        o.add_reader(reader_regular_grid.Reader(fx.grid_lon, fx.grid_lat, None, fx.times, zero, name='still_water'))
        o.set_config('general:use_auto_landmask', False)
        o.set_config('general:coastline_action', 'none')
        o.set_config('drift:advection_scheme', 'euler')
        o.set_config('drift:vertical_advection', False)
        o.set_config('environment:constant:horizontal_diffusivity', D)
        o.set_config('gpu:rng', 'philox')
        o.set_config('gpu:sort_interval_steps', sort_every)
        o.seed_elements(lon=lon, lat=lat, time=fx.start)
        o.run(steps=steps, time_step=dt, time_step_output=dt)
        assert o.num_elements_active() == n
        ids = np.asarray(o.elements.ID)
        order = np.argsort(ids)
        out.append((np.asarray(o.elements.lon)[order], np.asarray(o.elements.lat)[order]))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    dy = (out[0][1] - lat.astype(np.float32)) * 111194.9                       # metres north (float32 seeding rounding is < 1 m)
    expected = np.sqrt(2 * D * dt * steps)
    assert abs(dy.std() / expected - 1.0) < 0.03 and abs(dy.mean()) < 0.02 * expected


# ---- run-loop semantics that round 1's review found wrong -------------------------------------------------------
def _small_model(n=60, **cfg):
    from opendrift_b200.models.oceandrift import OceanDrift
    from opendrift_b200.readers import reader_regular_grid
    fx = Fixture('rk4_2d')
    o = OceanDrift(loglevel=50, seed=0)
    rd = reader_regular_grid.Reader(fx.grid_lon, fx.grid_lat, None, fx.times, {common.CUR[0]: fx.u, common.CUR[1]: fx.v}, name='current')
    o.add_reader(rd)
    o.set_config('general:use_auto_landmask', False)
    o.set_config('drift:vertical_advection', False)
    for k, v in cfg.items():
        o.set_config(k, v)
    return o, fx, rd


def test_fallback_follows_the_run_not_the_first_binding():
    """A reader queried before run() is bound without fallbacks; the run's environment:fallback:* must still reach the
    kernels (od_group_set_fallback): an element outside the reader's domain stays where it is instead of becoming NaN."""
    o, fx, rd = _small_model()
    rd.get_variables_interpolated(list(common.CUR), time=fx.start, lon=fx.lon0[:3], lat=fx.lat0[:3], z=np.zeros(3))   # auto-bind
    lon = np.append(fx.lon0[:5], fx.grid_lon[-1] + 3.0)          # the last element is outside the grid
    lat = np.append(fx.lat0[:5], fx.lat0[0])
    o.seed_elements(lon=lon, lat=lat, time=fx.start)
    o.run(steps=3, time_step=600)
    assert np.isfinite(o.elements.lon).all() and np.isfinite(o.elements.lat).all()
    assert abs(o.elements.lon[-1] - np.float32(lon[-1])) < 1e-12 and abs(o.elements.lat[-1] - np.float32(lat[-1])) < 1e-12
    # a second model with another fallback on the same bound reader
    o2, _, _ = _small_model()
    o2.env.readers.clear(), o2.env.priority_list.clear()
    o2.add_reader(rd)
    o2.set_config('environment:fallback:x_sea_water_velocity', 0.5)
    o2.seed_elements(lon=lon, lat=lat, time=fx.start)
    o2.run(steps=3, time_step=600)
    assert o2.elements.lon[-1] - np.float32(lon[-1]) > 0.01      # drifted east with the fallback current


def test_output_buffer_has_the_reference_time_axis_and_backfill():
    """state_to_buffer (basemodel/__init__.py:2384-2403): a fixed axis of output times; elements deactivated on a sub-step between
    output times are written, with their status, into the NEXT output column; the axis is cut at the time reached."""
    o, fx, rd = _small_model(**{'drift:deactivate_east_of': 0.0})
    n = 40
    lon0, lat0 = fx.lon0[:n].astype(np.float64), fx.lat0[:n]
    o.set_config('drift:deactivate_east_of', float(np.median(lon0)))     # half of the elements start outside -> leave at step 0
    o.set_config('drift:max_age_seconds', 1500)
    o.seed_elements(lon=lon0, lat=lat0, time=fx.start)
    res = o.run(steps=6, time_step=600, time_step_output=1800)
    east = np.float32(lon0) > np.float32(np.median(lon0))
    # step 0 is an output step: everybody is written there, the eastern half already with status 'outside'
    assert o.status_categories[:3] == ['active', 'outside', 'retired']
    st0 = res['status'][0]
    assert np.array_equal(st0 == 1, east) and np.array_equal(st0 == 0, ~east)
    # the others retire on the third step (age 1800 >= 1500) and nothing is left: the loop stops at 00:20 and the axis is cut
    # at the last output time reached (00:00); the retired elements were removed before any further state_to_buffer
    assert o.steps_calculation == 2 and len(res['time']) == 1 and res['time'][0] == fx.start
    assert res.lon.values.shape == (n, 1) and res.sizes == {'time': 1, 'trajectory': n}
    assert o.num_elements_active() == 0 and o.num_elements_deactivated() == n
    # without retirement: elements that cross the limit on a sub-step land in the next output column with status 'outside'
    o, fx, rd = _small_model()
    o.set_config('drift:deactivate_north_of', float(np.percentile(lat0, 60)) + 0.002)
    o.seed_elements(lon=lon0, lat=lat0, time=fx.start)
    res = o.run(steps=6, time_step=600, time_step_output=1800)
    assert [t for t in res['time']] == [fx.start + timedelta(seconds=1800 * k) for k in range(3)]     # 00:00, 00:30, 01:00
    status = res.status.values                                                      # [trajectory, time]
    gone = np.asarray(o.elements_deactivated.ID, dtype=np.int64)
    assert len(gone) > 0 and o.num_elements_active() + len(gone) == n
    for i in gone:                # a deactivated element ends with status 'outside' in the last column it appears in
        cols = np.where(status[i] >= 0)[0]
        assert status[i, cols[-1]] == o.status_categories.index('outside') and (status[i, cols[:-1]] == 0).all()
    alive = np.asarray(o.elements.ID, dtype=np.int64)
    assert (status[alive] == 0).all() and np.isfinite(res.lon.values[alive]).all()


def test_status_set_by_a_subclass_through_the_host_view_is_honoured():
    """A subclass may deactivate by writing elements.status directly (the reference tests status != 0 every step)."""
    from opendrift_b200.models.oceandrift import OceanDrift

    class Picky(OceanDrift):
        def update(self):
            super().update()
            if self.steps_calculation == 1:
                st = self.elements.status
                st[::2] = 5
                self.elements.status = st

    o, fx, rd = _small_model()
    p = Picky(loglevel=50, seed=0)
    p.add_reader(rd)
    p.set_config('general:use_auto_landmask', False)
    p.set_config('drift:vertical_advection', False)
    p.seed_elements(lon=fx.lon0[:20], lat=fx.lat0[:20], time=fx.start)
    p.run(steps=4, time_step=600)
    assert p.num_elements_active() == 10 and p.num_elements_deactivated() == 10
    assert np.array_equal(np.asarray(p.elements.ID), np.arange(1, 20, 2))
