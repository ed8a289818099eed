"""GPU: run-loop bookkeeping of the drop-in classes (release over several steps, per-element release times, retirement by
age, deactivate_north_of, order of the deactivated elements) against the unmodified reference's results
(tests/golden/bookkeeping_ref.npz; cases and checks in tests/bookkeeping.py, host-verified in tests/test_dropin_host.py)."""
import pytest

import bookkeeping as bk
import common

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('case', bk.CASES)
def test_run_loop_bookkeeping_matches_reference_on_gpu(case):
    o = bk.run_product(common.Fixture('rk4_3d'), case)
    n_act, n_deact = bk.check(o, case)
    assert n_act + n_deact == bk.N


def test_draw_order_with_mixing_and_horizontal_diffusion_on_gpu():
    """(see tests/test_dropin_host.py) the diffusion draws come after the mixing loop's"""
    import numpy as np
    from test_gpu_dropin import _model
    fx = common.Fixture('rk4_3d_mixing')
    fx.meta['diffusivity'] = 5.0
    pl, pa, pz = common.run_port(fx)
    o = _model(fx)
    o.run(steps=fx.steps, time_step=fx.dt, time_step_output=fx.dt)
    assert max(common.max_err_deg(o.elements.lon, o.elements.lat, pl, pa)) < 5e-8
    assert np.abs(o.elements.z - pz).max() < 1e-7


@pytest.mark.parametrize('scheme', bk.MULTI_SCHEMES)
def test_two_prioritised_current_readers_match_reference_on_gpu(scheme):
    o = bk.run_product_multireader(common.Fixture('rk4_2d'), scheme)
    bk.check_multireader(o, scheme)


@pytest.mark.parametrize('case', list(bk.LEEWAY_CASES))
def test_leeway_release_and_backward_cases_match_reference_on_gpu(case):
    bk.check_leeway(bk.run_product_leeway(common.LeewayFixture('leeway_piw1'), case), case)


@pytest.mark.parametrize('case', list(bk.od_cases()))
def test_option_combinations_match_reference_on_gpu(case):
    bk.check_od(bk.run_product_od(case), case)


def test_constant_reader_known_answers_on_gpu():
    """reader_constant on the GPU: tests/readers/test_variables.py:107-128 of the reference, as it is written there."""
    from datetime import datetime
    import numpy as np
    from opendrift_b200.models.oceandrift import OceanDrift
    from opendrift_b200.readers import reader_constant
    for direction, lon, lat in ((225, 3.932, 59.966), (45, 4.068, 60.034)):
        o = OceanDrift(loglevel=50)
        o.set_config('general:use_auto_landmask', False)
        o.add_reader(reader_constant.Reader({'wind_speed': 5, 'wind_to_direction': direction, 'land_binary_mask': 0}))
        o.seed_elements(lon=4, lat=60, time=datetime.now())
        o.run(steps=15)
        np.testing.assert_almost_equal(o.elements.lon, lon, 3)
        np.testing.assert_almost_equal(o.elements.lat, lat, 3)


@pytest.mark.parametrize('case', list(bk.run_cases()[1]))
def test_run_argument_variants_match_reference_on_gpu(case):
    bk.check_runcase(bk.run_product_runcase(case), case)


def test_three_reader_chain_3d_on_gpu():
    """step_chain_kernel: three current readers (different extents / level tables), w and wind readers, noise, diffusion."""
    bk.check_chain3d(bk.run_product_chain3d())


@pytest.mark.parametrize('kind,scheme', bk.HANDOVER_CASES)
def test_step_straddling_a_reader_hand_over_on_gpu(kind, scheme):
    bk.check_handover(bk.run_product_handover(kind, scheme), kind, scheme)


@pytest.mark.parametrize('case', list(bk.SUBCLASS_CASES))
def test_subclass_recipes_on_the_helpers_on_gpu(case):
    """Drift in sea ice (per-element factors on every helper, advect_with_sea_ice) and the windsea_swell Stokes profile."""
    assert bk.check_subclass(bk.run_product_subclass(case), case) > 0.01


@pytest.mark.parametrize('case', list(bk.SUBBLOCK_CASES))
def test_subblock_reader_on_gpu(case):
    """Sub-block readers: od_group_set_window + od_bbox on the device, blocks replaced under way, prefetch on the copy stream."""
    o, rd = bk.run_product_subblock(case)
    e, dz, window, n_windows = bk.check_subblock(o, rd, case)
    assert e < 5e-8 and dz <= 1e-5 and n_windows >= 1, (e, dz, window, n_windows)


def test_subblock_rewindow_on_gpu():
    o, rd = bk.run_product_rewindow(True)
    f, rf = bk.run_product_rewindow(False)
    assert rd.windows_set >= 2
    import common
    import numpy as np
    e = max(common.max_err_deg(np.asarray(o.elements.lon), np.asarray(o.elements.lat), np.asarray(f.elements.lon), np.asarray(f.elements.lat)))
    assert e < 1e-6, e


@pytest.mark.parametrize('which', bk.HOOK_CASES)
def test_mixing_loop_hooks_on_gpu(which):
    assert bk.check_hooks(bk.run_product_hooks(which), which) <= 1e-9


@pytest.mark.parametrize('case', list(bk.PROJ_CASES))
def test_readers_on_a_projected_plane_on_gpu(case):
    e, dz, moved = bk.check_proj(bk.run_product_proj(case), case)
    assert moved > 0.01 and e < 5e-8 and dz <= 1e-9, (e, dz)
