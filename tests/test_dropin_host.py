"""The reference-facing Python surface (OceanDrift / Leeway / readers / Environment) end to end WITHOUT a GPU: the tests of
tests/test_gpu_dropin.py (and the drop-in replay of the reference's test_dateline) re-run with the engine swapped for
tests/hostengine.py -- Engine's own methods and ctypes argument structs, with the library replaced by the host build of
the same device sources (tests/hostshim).  What runs here is the host logic of the product: seeding, release, the run
loop, Environment, reader binding and the slab ring, the fused / helper step recipes, draws and their order, compaction.

One difference from the GPU is known and allowed for: torch's CPU float32 sqrt is not correctly rounded (0.7 % of values are
an ulp off; CUDA's sqrt.rn is IEEE), so the wind speed that feeds the analytical diffusivity models can differ by an ulp and
depths of those three fixtures move by up to ~2e-6 m."""
import numpy as np
import pytest

import common
import test_gpu_dropin as T
import test_dateline as TD
from hostengine import HostEngine


@pytest.fixture(autouse=True)
def host_engine(monkeypatch):
    eng = HostEngine()
    import opendrift_b200.engine as E
    import opendrift_b200.models.basemodel as B
    monkeypatch.setattr(E, 'default_engine', lambda device=None: eng)
    monkeypatch.setattr(B, 'default_engine', lambda device=None: eng)
    yield eng


ANALYTIC_MIXING = ('euler_3d_mixing_sundby1983', 'rk4_3d_mixing_large1994', 'rk2_3d_mixing_env_fallback')


@pytest.mark.parametrize('name', common.fixtures())
def test_oceandrift_run_matches_reference(name, host_engine):
    fx = common.Fixture(name)
    o = T._model(fx)
    o.run(steps=fx.steps, time_step=fx.dt, time_step_output=fx.dt)
    assert o.num_elements_active() == fx.n
    lon, lat, z = o.elements.lon, o.elements.lat, o.elements.z
    assert lon.dtype == np.float64 and z.dtype == fx.z.dtype
    assert max(common.max_err_deg(lon, lat, fx.lon, fx.lat)) < 5e-8
    ztol = 1e-5 if name in ANALYTIC_MIXING else common.z_tolerance(fx.meta, exact=1e-9)
    assert np.abs(z - fx.z).max() <= ztol
    assert np.array_equal(o.elements.ID, np.arange(fx.n)[::-1] if fx.dt < 0 else np.arange(fx.n))
    assert len(o.history['time']) == fx.steps + 1
    if name not in ANALYTIC_MIXING:
        # the same steps driven through the bare argument structs (tests/common.py:run_hostshim): bit for bit
        hl, ha, hz = common.run_hostshim(fx, fast=2)
        assert np.array_equal(lon, hl) and np.array_equal(lat, ha) and np.array_equal(z, hz)
    assert 'od_step_oceandrift' in host_engine.lib.calls            # the fused recipe, one launch per step


# the remaining drop-in tests as they are written for the GPU
test_overridden_update_uses_helpers_and_matches = T.test_overridden_update_uses_helpers_and_matches
test_subclass_touching_numpy_state_still_works = T.test_subclass_touching_numpy_state_still_works
test_reader_get_variables_interpolated_and_environment = T.test_reader_get_variables_interpolated_and_environment
test_leeway_model_matches_reference = T.test_leeway_model_matches_reference
test_leeway_missing_forcing_deactivates = T.test_leeway_missing_forcing_deactivates
test_seeding_radius_and_deactivation = T.test_seeding_radius_and_deactivation
test_reference_known_answers_with_constant_environment = T.test_reference_known_answers_with_constant_environment
test_arithmetic_config_selects_the_kernel_policy = T.test_arithmetic_config_selects_the_kernel_policy
test_fallback_follows_the_run_not_the_first_binding = T.test_fallback_follows_the_run_not_the_first_binding
test_output_buffer_has_the_reference_time_axis_and_backfill = T.test_output_buffer_has_the_reference_time_axis_and_backfill
test_status_set_by_a_subclass_through_the_host_view_is_honoured = T.test_status_set_by_a_subclass_through_the_host_view_is_honoured


def test_dropin_model_replays_the_reference_dateline_test():
    TD.test_gpu_dropin_model_replays_the_reference_test()


def test_device_rng_for_diffusion_is_order_independent():
    T.test_device_rng_for_diffusion_is_order_independent()


def test_draw_order_with_mixing_and_horizontal_diffusion():
    """The legacy generator is shared: update() (mixing loop draws) runs before horizontal_diffusion() (two normal
    draws), basemodel/__init__.py:2272-2280.  The fused recipe once drew in the other order (found with this harness)."""
    fx = common.Fixture('rk4_3d_mixing')
    fx.meta['diffusivity'] = 5.0                      # a combination no reference fixture has
    pl, pa, pz = common.run_port(fx)
    o = T._model(fx)
    o.run(steps=fx.steps, time_step=fx.dt, time_step_output=fx.dt)
    assert max(common.max_err_deg(o.elements.lon, o.elements.lat, pl, pa)) < 5e-8
    assert np.abs(o.elements.z - pz).max() < 1e-7


@pytest.mark.parametrize('family,seed', [('basic', 1), ('basic', 9), ('physics', 2), ('physics', 7), ('physics', 13), ('options', 3),
                                         ('options', 11), ('options', 21), ('readers', 3), ('readers', 4), ('readers', 14),
                                         ('wdf', 17), ('wdf', 21), ('wdf', 24)])
def test_random_scenarios_through_the_model_classes(family, seed):
    """The randomised whole-run scenarios of tests/test_hostmath.py (which drive the bare step loop) through
    OceanDrift.run(): readers on their own grids, global / periodic grids, drift-factor arrays, truncation, noise, Stokes,
    mixing.  Must equal the bare loop bit for bit and the port within the position tolerance."""
    import test_hostmath as th
    fx = {'basic': th._random_scenario, 'physics': th._random_physics_scenario, 'options': th._random_options_scenario,
          'readers': th._random_readers_scenario, 'wdf': th._random_wdf_scenario}[family](seed)
    if fx is None:
        pytest.skip('scenario not applicable for this seed')
    o = T._model(fx)
    o.run(steps=fx.steps, time_step=fx.dt, time_step_output=fx.dt)
    lon, lat, z = o.elements.lon, o.elements.lat, np.asarray(o.elements.z)
    hl, ha, hz = common.run_hostshim(fx, fast=2)
    analytic = fx.meta.get('mixing') and fx.meta.get('diffusivity_model') not in (None, 'environment')
    # wind uncertainty without a wind reader: the reference (and the model classes, and the port) add it to the fallback wind and
    # drift with it; the bare loop of tests/common.py has no wind group to hang it on
    wind_noise_only = bool((fx.meta.get('noise') or {}).get('wind')) and not fx.meta['wind']
    if not analytic and not wind_noise_only:           # (torch's CPU sqrt, see the module docstring)
        assert np.array_equal(lon, hl) and np.array_equal(lat, ha) and np.array_equal(z, np.asarray(hz))
    pl, pa, pz = common.run_port(fx)
    assert max(common.max_err_deg(lon, lat, pl, pa)) < 5e-8


@pytest.mark.parametrize('case', __import__('bookkeeping').CASES)
def test_run_loop_bookkeeping_matches_reference(case):
    """Release over several steps, per-element release times, retirement by age, deactivate_north_of and the order of
    the deactivated elements, against the unmodified reference's results (tests/golden/bookkeeping_ref.npz)."""
    import bookkeeping as bk
    o = bk.run_product(common.Fixture('rk4_3d'), case)
    n_act, n_deact = bk.check(o, case)
    assert n_act + n_deact == bk.N
    if case in ('release_max_age', 'release_deactivate_north'):
        assert n_deact > 100


@pytest.mark.parametrize('scheme', __import__('bookkeeping').MULTI_SCHEMES)
def test_two_prioritised_current_readers_match_reference(scheme, host_engine):
    """A nested reader inside a coarser one (the reference loops over the readers on the still-missing elements at every
    get_environment call, environment.py:613-780).  Found with this harness against the live reference (the fused launch
    once sampled the first reader only); results pinned in tests/golden/bookkeeping_ref.npz.  Three ways through the product:
    the reader chain inside the fused kernel, the chain inside od_advect_current (helper recipe), and the staged recipe."""
    import bookkeeping as bk
    from opendrift_b200.models.oceandrift import OceanDrift
    fx = common.Fixture('rk4_2d')
    o = bk.run_product_multireader(fx, scheme)
    bk.check_multireader(o, scheme)
    assert host_engine.lib.calls.count('od_step_oceandrift') == bk.MULTI_STEPS          # one launch per step, chain inside

    class Recipe(OceanDrift):
        def update(self):
            self.advect_ocean_current()
            self.advect_wind()
    host_engine.lib.calls.clear()
    o = bk.run_product_multireader(fx, scheme, cls=Recipe)
    bk.check_multireader(o, scheme)
    assert host_engine.lib.calls.count('od_advect_current') == bk.MULTI_STEPS and 'od_step_oceandrift' not in host_engine.lib.calls

    class Staged(Recipe):
        def _current_chain(self, t):
            return None
    host_engine.lib.calls.clear()
    o = bk.run_product_multireader(fx, scheme, cls=Staged)
    bk.check_multireader(o, scheme)
    assert 'od_advect_current' not in host_engine.lib.calls and 'od_step_oceandrift' not in host_engine.lib.calls


def test_models_refuse_reader_lists_they_cannot_follow():
    """Leeway's fused launch and the mixing kernel sample one reader per variable: with several candidates they raise
    instead of silently ignoring all but the first."""
    from opendrift_b200.models.leeway import Leeway
    from opendrift_b200.readers import reader_regular_grid
    fx = common.LeewayFixture('leeway_piw1')
    o = Leeway(loglevel=50, seed=1)
    cur = {common.CUR[0]: fx.u, common.CUR[1]: fx.v}
    o.add_reader([reader_regular_grid.Reader(fx.grid_lon, fx.grid_lat, None, fx.times, cur, name='c1'),
                  reader_regular_grid.Reader(fx.grid_lon, fx.grid_lat, None, fx.times, cur, name='c2'),
                  reader_regular_grid.Reader(fx.grid_lon, fx.grid_lat, None, fx.times, {'x_wind': fx.x_wind, 'y_wind': fx.y_wind})])
    o.seed_elements(lon=fx.lon0[:10], lat=fx.lat0[:10], time=fx.start, object_type=1)
    with pytest.raises(NotImplementedError, match='several readers'):
        o.run(steps=2, time_step=600)


@pytest.mark.parametrize('name', common.fixtures())
def test_every_fixture_through_the_helper_recipe(name):
    """A subclass that overrides update() with the reference's own recipe (oceandrift.py:185-211) gets the helpers as
    separate launches; they must give the fixture's result like the fused launch does (this found the start-of-step
    environment ignoring drift:truncate_ocean_model_below_m on that path)."""
    from opendrift_b200.models.oceandrift import OceanDrift

    class Recipe(OceanDrift):
        def update(self):
            self.advect_ocean_current()
            self.advect_wind()
            self.stokes_drift()
            if self.get_config('drift:vertical_mixing'):
                self.vertical_mixing()
            self.vertical_advection()
    fx = common.Fixture(name)
    o = T._model(fx)
    o.__class__ = Recipe
    o.run(steps=fx.steps, time_step=fx.dt)
    assert max(common.max_err_deg(o.elements.lon, o.elements.lat, fx.lon, fx.lat)) < 5e-8
    assert np.abs(o.elements.z - fx.z).max() <= 1e-5


@pytest.mark.parametrize('case', list(__import__('bookkeeping').LEEWAY_CASES))
def test_leeway_release_and_backward_cases_match_reference(case):
    """Leeway with a release interval, backward runs, capsizing in both directions: the unmodified reference's final
    positions, orientation, capsized flags and (jibed) crosswind slopes (tests/golden/bookkeeping_ref.npz)."""
    import bookkeeping as bk
    bk.check_leeway(bk.run_product_leeway(common.LeewayFixture('leeway_piw1'), case), case)


@pytest.mark.parametrize('case', list(__import__('bookkeeping').od_cases()))
def test_option_combinations_match_reference(case):
    """OceanDrift option combinations no fixture has -- constant / fallback forcing next to gridded readers, a reader that
    ends mid-run, relative wind, wind-drift depth, uncertainty + diffusion, drift-factor arrays, and the vertical-mixing variants
    (terminal velocity, mixing at the surface, shallow sea floor, non-divisor inner step, constant model, backward) -- against
    results of the unmodified reference driven through its own model class (tests/golden/bookkeeping_ref.npz)."""
    import bookkeeping as bk
    bk.check_od(bk.run_product_od(case), case)


def test_constant_reader_replays_the_reference_known_answers():
    """opendrift/readers/reader_constant.py on the drop-in classes: the reference's tests/readers/test_variables.py:107-128
    (wind from speed + direction, 15 one-hour steps, default wind drift factor) as the reference writes it, and constant
    current / wind readers next to each other (values from a run of the unmodified reference)."""
    from datetime import datetime, timedelta
    from opendrift_b200.models.oceandrift import OceanDrift
    from opendrift_b200.readers import reader_constant
    for direction, lon, lat in ((225, 3.932, 59.966), (45, 4.068, 60.034)):
        r = reader_constant.Reader({'wind_speed': 5, 'wind_to_direction': direction, 'land_binary_mask': 0})
        o = OceanDrift(loglevel=50)
        o.set_config('general:use_auto_landmask', False)
        o.add_reader(r)
        o.seed_elements(lon=4, lat=60, time=datetime.now())
        o.run(steps=15)
        np.testing.assert_almost_equal(o.elements.lon, lon, 3)
        np.testing.assert_almost_equal(o.elements.lat, lat, 3)
    o = OceanDrift(loglevel=50)
    o.set_config('general:use_auto_landmask', False)
    o.add_reader([reader_constant.Reader({'x_sea_water_velocity': 1, 'y_sea_water_velocity': 0, 'x_wind': 10, 'y_wind': 0})])
    o.seed_elements(lon=4.0, lat=60.0, time=datetime(2026, 1, 1), current_drift_factor=0.3, wind_drift_factor=0.02)
    o.run(steps=2, time_step=3600)
    assert o.elements.lon[0] == pytest.approx(4.064516123645828, abs=1e-12)       # the unmodified reference, same script
    assert o.elements.lat[0] == pytest.approx(59.99999590372567, abs=1e-12)
    with pytest.raises(NotImplementedError):
        reader_constant.Reader({'x_wind': np.arange(3.0), 'element_ID': np.arange(3)})


@pytest.mark.parametrize('case', list(__import__('bookkeeping').run_cases()[1]))
def test_run_argument_variants_match_reference(case):
    """run(duration= / end_time= / steps=, time_step_output=, timedelta steps, config time steps, backward duration),
    number_per_point seeding with a radius, deactivate_west/south_of: step count, end time, surviving IDs and positions of
    the unmodified reference."""
    import bookkeeping as bk
    bk.check_runcase(bk.run_product_runcase(case), case)


def test_three_reader_chain_3d_with_everything_on(host_engine):
    """Three current readers with different extents and level tables (one of them 2-D), vertical velocity and wind on
    readers of their own, current uncertainty and horizontal diffusion, RK4: the reader chain inside the fused launch and
    the staged recipe against the unmodified reference."""
    import bookkeeping as bk
    from opendrift_b200.models.oceandrift import OceanDrift
    bk.check_chain3d(bk.run_product_chain3d())
    assert host_engine.lib.calls.count('od_step_oceandrift') == bk.CHAIN3D_STEPS

    class Staged(OceanDrift):
        def _current_chain(self, t):
            return None
    host_engine.lib.calls.clear()
    bk.check_chain3d(bk.run_product_chain3d(cls=Staged))
    assert 'od_step_oceandrift' not in host_engine.lib.calls

    class Recipe(OceanDrift):                      # helper recipe: the chain (and the per-stage noise) inside od_advect_current
        def update(self):
            self.advect_ocean_current()
            self.advect_wind()
            self.stokes_drift()
            self.vertical_advection()
    host_engine.lib.calls.clear()
    bk.check_chain3d(bk.run_product_chain3d(cls=Recipe))
    assert host_engine.lib.calls.count('od_advect_current') == bk.CHAIN3D_STEPS and 'od_step_oceandrift' not in host_engine.lib.calls


@pytest.mark.parametrize('seed', [2, 4, 10, 15, 19])
def test_random_reader_chains_fused_equals_staged(seed):
    """A random sub-box reader in front of a surface-only reader over the whole (possibly periodic / descending-axis) grid of a
    random scenario, forward or backward: the reader chain inside the fused launch against the staged recipe -- two
    independent implementations of the reference's reader loop (24 seeds during development)."""
    import test_hostmath as th
    from opendrift_b200.models.oceandrift import OceanDrift
    from opendrift_b200.readers import reader_regular_grid
    fx = th._random_scenario(seed)
    rng = np.random.default_rng(seed)
    nx, ny = len(fx.grid_lon), len(fx.grid_lat)
    i0, i1 = sorted(rng.choice(np.arange(2, nx - 2), 2, replace=False))
    j0, j1 = sorted(rng.choice(np.arange(2, ny - 2), 2, replace=False))
    i1, j1 = max(i1, min(nx, i0 + 3)), max(j1, min(ny, j0 + 3))
    three_d = fx.grid_z is not None
    cur = common.CUR

    def run(staged):
        class M(OceanDrift):
            pass
        if staged:
            M._current_chain = lambda self, t: None
        eng = HostEngine()
        o = M(loglevel=50, seed=0, engine=eng)
        o.add_reader([reader_regular_grid.Reader(fx.grid_lon[i0:i1], fx.grid_lat[j0:j1], fx.grid_z, fx.times,
                                                 {cur[0]: fx.u[..., j0:j1, i0:i1].copy(), cur[1]: fx.v[..., j0:j1, i0:i1].copy()}, name='A'),
                      reader_regular_grid.Reader(fx.grid_lon, fx.grid_lat, None, fx.times,
                                                 {cur[0]: (0.5 * (fx.u[:, 0] if three_d else fx.u)).astype(np.float32),
                                                  cur[1]: (-0.8 * (fx.v[:, 0] if three_d else fx.v)).astype(np.float32)}, name='B')])
        o.set_config('general:use_auto_landmask', False)
        o.set_config('drift:advection_scheme', fx.meta['scheme'])
        o.set_config('drift:vertical_advection', False)
        o.seed_elements(lon=fx.lon0, lat=fx.lat0, z=fx.z0, time=fx.start)
        o.run(steps=fx.steps, time_step=fx.dt, time_step_output=fx.dt)
        return np.asarray(o.elements.lon), np.asarray(o.elements.lat), eng.lib.calls
    fl, fa, fcalls = run(False)
    sl, sa, scalls = run(True)
    assert 'od_step_oceandrift' in fcalls and 'od_step_oceandrift' not in scalls
    assert max(common.max_err_deg(fl, fa, sl, sa)) < 5e-8


@pytest.mark.parametrize('kind,scheme', __import__('bookkeeping').HANDOVER_CASES)
def test_step_straddling_a_reader_hand_over(kind, scheme):
    """Readers that follow each other in time (daily files as separate readers): the reference makes one get_environment call
    per Runge-Kutta stage, each picking the readers that cover ITS time, so a step that starts in one reader's coverage and
    ends in the next one's samples both; a reader whose coverage begins inside a step serves the later stages.  Found against
    the live reference with this harness (the stage after the hand-over once got the fallback)."""
    import bookkeeping as bk
    bk.check_handover(bk.run_product_handover(kind, scheme), kind, scheme)


def test_subclass_helpers_wind_speed_current_speed_direction():
    """wind_speed() / current_speed() / simulation_direction(): what model subclasses read inside update()."""
    from opendrift_b200.models.oceandrift import OceanDrift
    seen = {}

    class Probe(OceanDrift):
        def update(self):
            seen['w'], seen['c'], seen['d'] = self.wind_speed(), self.current_speed(), self.simulation_direction()
            self.advect_ocean_current()
    fx = common.Fixture('rk4_3d_full')
    o = T._model(fx)
    o.__class__ = Probe
    o.run(steps=1, time_step=fx.dt)
    assert seen['d'] == 1 and seen['w'].dtype == np.float32 and seen['w'].shape == (fx.n,) and seen['c'].max() > 0
    lib = common.hostshim()
    wind = common.HsField(fx.wind_lon, fx.wind_lat, None, [fx.x_wind, fx.y_wind], fx.times)
    xw, yw = wind.sample(lib, fx.start, fx.lon0.astype(np.float64), fx.lat0.astype(np.float64), fx.z0, True)
    assert np.array_equal(seen['w'], np.sqrt(xw**2 + yw**2))


def test_leeway_with_wind_uncertainty_scatters_the_elements():
    """drift:wind_uncertainty reaches the fused Leeway launch (it was refused until the launch learnt to add the step's draws): the
    same seed drifts elsewhere with it."""
    from opendrift_b200.models.leeway import Leeway
    from opendrift_b200.readers import reader_constant
    lons = {}
    for wu in (0.0, 2.0):
        o = Leeway(loglevel=50, seed=1)
        o.add_reader(reader_constant.Reader({'x_wind': 5, 'y_wind': 0, 'x_sea_water_velocity': 0.1, 'y_sea_water_velocity': 0}))
        o.set_config('drift:wind_uncertainty', wu)
        o.seed_elements(lon=4, lat=60, number=10, time=__import__('datetime').datetime(2026, 1, 1), object_type=1)
        o.run(steps=2, time_step=600)
        lons[wu] = np.asarray(o.elements.lon).copy()
    assert np.abs(lons[2.0] - lons[0.0]).max() > 1e-4         # (the runs with the reference's draws are in tests/bookkeeping.py)


@pytest.mark.parametrize('case', list(__import__('bookkeeping').SUBCLASS_CASES))
def test_subclass_recipes_on_the_helpers_match_reference(case):
    """Subclasses that drive the helpers with per-element factors -- drift in sea ice as OpenOil's advect_oil does it
    (advect_ocean_current / advect_wind with 1 - k_ice, stokes_drift(factor array), advect_with_sea_ice) -- and the combined
    swell / wind-sea Stokes profile, against the same subclass body on the unmodified reference."""
    import bookkeeping as bk
    o = bk.run_product_subclass(case)
    assert bk.check_subclass(o, case) > 0.01


@pytest.mark.parametrize('case', list(__import__('bookkeeping').SUBBLOCK_CASES))
def test_subblock_reader_matches_reference(case):
    """A reader that hands out sub-blocks around the elements: the device blocks take the block's own index geometry
    (od_group_set_window), as the reference's ReaderBlock does; compared with the unmodified reference run on the same kind of
    reader.  (With the whole grid instead of the block the first, float32-position step differs by ~4e-8 deg.)"""
    import bookkeeping as bk
    o, rd = bk.run_product_subblock(case)
    e, dz, window, n_windows = bk.check_subblock(o, rd, case)
    full = (len(common.Fixture('rk4_3d_full').grid_lat), len(common.Fixture('rk4_3d_full').grid_lon))
    assert window[0] < full[0] and window[1] < full[1] and n_windows >= 1
    assert e < 2e-9 and dz == 0.0, (e, dz)     # the reference's block, the reference's index arithmetic


def test_subblock_reader_is_asked_for_new_blocks_when_the_elements_leave():
    import bookkeeping as bk
    o, rd = bk.run_product_rewindow(True)
    f, rf = bk.run_product_rewindow(False)
    assert rd.windows_set >= 2 and rf.windows_set == 0
    g = rd.group_of(common.CUR[0])[0]
    assert g.desc.nx < 40 and g.desc.ny < 36                 # a window, not the 36 x 40 grid
    e = max(common.max_err_deg(np.asarray(o.elements.lon), np.asarray(o.elements.lat), np.asarray(f.elements.lon), np.asarray(f.elements.lat)))
    # block-relative against whole-grid index arithmetic (float32 on the first step), grown over 24 one-hour steps: 2e-7 deg
    assert e < 1e-6, e


@pytest.mark.parametrize('which', __import__('bookkeeping').HOOK_CASES)
def test_mixing_loop_hooks_of_a_subclass_match_reference(which, host_engine):
    """A subclass that overrides surface_stick / surface_wave_mixing / update_terminal_velocity / prepare_vertical_mixing gets the
    inner loop one launch per iteration with its hooks in between, in the reference's order, the legacy generator's draws
    interleaved as the reference interleaves them; compared with the same subclass body on the unmodified reference."""
    import bookkeeping as bk
    o = bk.run_product_hooks(which)
    dz = bk.check_hooks(o, which)
    assert dz <= 1e-9, dz
    n_mix = host_engine.lib.calls.count('od_vertical_mixing')
    assert n_mix == 3 * 10                                  # 3 steps x 10 inner iterations, one launch each
    if which == 'all':
        assert o.prepared == 3


def test_per_iteration_mixing_launches_continue_the_device_generator(host_engine):
    """gpu:rng = philox: ten launches of one inner iteration each (a subclass with a hook) draw what the fused ten-iteration
    launch draws -- the depths are the same bit for bit."""
    import bookkeeping as bk
    from opendrift_b200.models.oceandrift import OceanDrift
    from opendrift_b200.readers import reader_regular_grid
    make = lambda lon, lat, z, t, f, name: reader_regular_grid.Reader(lon, lat, z, t, f, name=name)     # noqa: E731

    class Philox(OceanDrift):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self.set_config('gpu:rng', 'philox')
    hooked = bk.run_hook_case('same_stick', Philox, make)
    plain = bk.run_hook_case('none', Philox, make)
    assert host_engine.lib.calls.count('od_vertical_mixing') == 30 + 3
    assert np.array_equal(np.asarray(hooked.elements.z), np.asarray(plain.elements.z))
    assert np.array_equal(np.asarray(hooked.elements.lon), np.asarray(plain.elements.lon))


@pytest.mark.parametrize('case', list(__import__('bookkeeping').PROJ_CASES))
def test_readers_on_a_projected_plane_match_reference(case, host_engine):
    """Gridded readers whose grid lies on a stereographic plane (polar and oblique aspects): positions are projected before the
    index arithmetic and the current / wind components rotated from the grid's axes to east / north afterwards, inside the
    launches -- fused OceanDrift step (RK4 3-D with w, RK2 with a wind reader), vertical mixing, Leeway -- against the unmodified
    reference running readers of the same projection (pyproj stood in for by oracle/proj_stere.py + oracle/geod_karney.py)."""
    import bookkeeping as bk
    o = bk.run_product_proj(case)
    e, dz, moved = bk.check_proj(o, case)
    assert moved > 0.01 and e < 5e-8 and dz <= 1e-9, (e, dz)
    step = 'od_leeway_step' if case.endswith('leeway') else 'od_step_oceandrift'
    assert step in host_engine.lib.calls                          # the fused launches, not the staged recipe


def test_projected_reader_interpolation_with_and_without_rotation(host_engine):
    """get_variables_interpolated of a projected reader: components along the grid's axes unless rotate_to_proj names a
    geographic CRS (variables.py:799-837); the rotated pair has the same speed, turned by the local grid azimuth."""
    import bookkeeping as bk
    from opendrift_b200.readers import reader_regular_grid
    c, xs, ys, zs, times, fields, lon0, lat0, z0 = bk.proj_setup('stere_polar_rk4_3d_w')
    rd = reader_regular_grid.Reader(xs, ys, zs, times, fields['cur3d'], name='cur', proj4=c['proj4'])
    lon, lat, z = lon0.astype(np.float64), lat0.astype(np.float64), z0
    raw, _ = rd.get_variables_interpolated(list(common.CUR), time=times[1], lon=lon, lat=lat, z=z)
    rot, _ = rd.get_variables_interpolated(list(common.CUR), time=times[1], lon=lon, lat=lat, z=z, rotate_to_proj='+proj=latlong')
    u0, v0, u1, v1 = (np.asarray(a, dtype=np.float64) for a in (raw[common.CUR[0]], raw[common.CUR[1]], rot[common.CUR[0]], rot[common.CUR[1]]))
    assert np.allclose(np.hypot(u0, v0), np.hypot(u1, v1), rtol=2e-6)
    turn = np.degrees(np.arctan2(u1, v1) - np.arctan2(u0, v0))
    turn = (turn + 180) % 360 - 180
    # polar stereographic with lon_0 = 10: the grid's y axis points along the meridian 10 E + 180, so east of it the grid is turned
    # clockwise by (lon - 10) degrees
    assert np.abs(turn - (-(lon - 10.0))).max() < 0.2 or np.abs(turn - (lon - 10.0)).max() < 0.2
    assert np.abs(turn).max() > 1.0


def test_services_a_reference_style_update_calls(host_engine):
    """timer_start / timer_end (timer.py:26-34), store_message (:4736-4740), water_column_stretching (oceandrift.py:299-313) and
    list_configspec (config.py:34-52): an update() written for the reference calls them; they exist with the reference's meaning."""
    from datetime import datetime
    from opendrift_b200.models.oceandrift import OceanDrift

    class Model(OceanDrift):
        def update(self):
            self.timer_start('main loop:updating elements:my physics')
            self.water_column_stretching()
            self.advect_ocean_current()
            self.timer_end('main loop:updating elements:my physics')
            self.store_message('step %d' % self.steps_calculation)

    o = Model(loglevel=50)
    o.set_config('environment:constant:x_sea_water_velocity', 1)
    o.set_config('environment:constant:land_binary_mask', 0)
    o.seed_elements(lon=3, lat=60, time=datetime(2024, 1, 1))
    o.run(steps=2)
    assert abs(float(o.elements.lon[0]) - 3.129) < 1e-3
    assert 'my physics' in o.performance() and o.timing['main loop:updating elements:my physics'].total_seconds() >= 0
    assert o.get_messages() == 'step 0\nstep 1\n'
    o.list_configspec('drift:advection')
