"""Global (east-west periodic) readers: the reference's known-answer test tests/readers/test_interpolation.py:43-148
(test_dateline), replayed with in-memory readers -- a current reader on lon 0..359 and a wind reader on lon -180..179.
The fixtures ref_dateline_*.npz were produced by the UNMODIFIED reference (oracle/make_golden.py:run_dateline_case);
the numbers asserted here are the ones the reference's test asserts."""
import numpy as np
import pytest

import common
from common import Fixture


def _known_answers_euler(lon, lat):
    # test_interpolation.py:100-112: simulation across the 0 meridian and across the dateline
    np.testing.assert_array_almost_equal(lon[:4], [-2.129, 2.129, -175.129, 175.129], decimal=3)
    np.testing.assert_array_almost_equal(lat[:4], [60.006, 59.994, 60.006, 59.994], decimal=3)


def _known_answers_spread(lon, lat):
    # test_interpolation.py:126-142 (last output step of the globally spread elements)
    for idx, lo, la in ((25, -61.55, -59.92), (100, -80.83, 20.08), (35, 141.55, -60.07), (112, 160.83, 19.92)):
        assert abs(lon[idx] - lo) < 0.006 and abs(lat[idx] - la) < 0.006, (idx, lon[idx], lat[idx])


def test_reference_fixtures_hold_the_reference_known_answers():
    fx = Fixture('dateline_piecewise_euler')
    _known_answers_euler(fx.lon, fx.lat)
    fx = Fixture('dateline_piecewise_spread')
    _known_answers_spread(fx.lon, fx.lat)


def test_port_and_host_math_on_global_readers():
    for name in ('dateline_piecewise_euler', 'dateline_piecewise_spread', 'dateline_smooth_rk4', 'dateline_smooth_euler'):
        fx = Fixture(name)
        lon, lat, _ = common.run_port(fx)
        assert np.array_equal(lon, fx.lon) and np.array_equal(lat, fx.lat), name          # the port is the reference, bit for bit
        for mode in (0, 2):
            hl, ha, _ = common.run_hostshim(fx, fast=mode)
            e = common.max_err_deg(hl, ha, fx.lon, fx.lat)
            assert max(e) < (1e-11 if fx.meta['scheme'] == 'euler' else 2e-9), (name, mode, e)
    hl, ha, _ = common.run_hostshim(Fixture('dateline_piecewise_euler'), fast=2)
    _known_answers_euler(hl, ha)


def test_geometry_of_global_grids():
    from opendrift_b200.engine import grid_geometry
    lat = np.arange(-88, 89)
    g = grid_geometry(np.arange(0, 360), lat)
    assert g['wrap_x'] == 1 and g['global_coverage'] and g['xspan'] == 360.0 and g['x0'] == 0.0
    g = grid_geometry(np.arange(-180, 180), lat)
    assert g['wrap_x'] == 1 and g['xspan'] == 360.0 and g['x0'] == -180.0
    g = grid_geometry(np.arange(160, 280), lat)                       # the Pacific wind reader of the reference test
    assert g['wrap_x'] == 0 and g['global_x'] == 0 and not g['global_coverage'] and g['xspan'] == 119.0
    g = grid_geometry(np.arange(0, 361), lat)                         # global with a duplicated end column: no virtual column
    assert g['wrap_x'] == 0 and g['global_coverage'] and g['global_x'] == 1
    g = grid_geometry(np.linspace(-180.01, 175.01, 9), lat)           # global by the reference's rule (xmin - 2 dx <= -180 ...), not periodic
    assert g['wrap_x'] == 0 and g['global_x'] == 1
    g = grid_geometry(np.arange(0.25, 360, 0.5), lat)
    assert g['wrap_x'] == 1 and abs(g['xspan'] - 360.0) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['dateline_piecewise_euler', 'dateline_piecewise_spread', 'dateline_smooth_rk4',
                                  'dateline_smooth_euler'])
def test_gpu_engine_on_global_readers(name):
    fx = Fixture(name)
    lon, lat, _ = common.run_engine(fx, fused=True)
    e = common.max_err_deg(lon, lat, fx.lon, fx.lat)
    assert max(e) < (1e-11 if fx.meta['scheme'] == 'euler' else 2e-9), (name, e)
    if name == 'dateline_piecewise_euler':
        _known_answers_euler(lon, lat)
    if name == 'dateline_piecewise_spread':
        _known_answers_spread(lon, lat)


@pytest.mark.gpu
def test_gpu_dropin_model_replays_the_reference_test():
    """The reference's test, written the way the reference writes it, on the GPU classes."""
    from datetime import datetime, timedelta
    from opendrift_b200.models.oceandrift import OceanDrift
    from opendrift_b200.readers import reader_regular_grid
    lat = np.arange(-88, 89)
    start_time = datetime(2021, 1, 1)
    time = [start_time + i * timedelta(hours=24) for i in range(3)]
    xcurr = np.zeros((3, len(lat), 360), np.float32); ycurr = np.zeros_like(xcurr)
    xcurr[:, :, 0:180] = 1
    xcurr[:, :, 180:] = -1
    xwind = np.zeros((3, len(lat), 360), np.float32); ywind = np.zeros_like(xwind)
    ywind[:, :, 0:180] = 1
    ywind[:, :, 180:] = -1
    xw2 = np.zeros((3, len(lat), 120), np.float32); yw2 = np.zeros_like(xw2)
    yw2[:, :, 0:20] = -1
    yw2[:, :, 20:] = 1

    def readers(pacific_wind=False):
        rc = reader_regular_grid.Reader(np.arange(0, 360), lat, None, time,
                                        {'x_sea_water_velocity': xcurr, 'y_sea_water_velocity': ycurr}, name='current')
        if pacific_wind:
            rw = reader_regular_grid.Reader(np.arange(160, 280), lat, None, time, {'x_wind': xw2, 'y_wind': yw2}, name='wind2')
        else:
            rw = reader_regular_grid.Reader(np.arange(-180, 180), lat, None, time, {'x_wind': xwind, 'y_wind': ywind}, name='wind')
        return rc, rw

    rc, rw = readers()
    assert list(rw.covers_positions(np.array([-175, 0, 175]), np.array([60, 60, 60]))[0]) == [0, 1, 2]
    assert list(readers(True)[1].covers_positions(np.array([-175, 0, 175]), np.array([60, 60, 60]))[0]) == [0, 2]
    for seeds, expect_lon, pacific in (([-2, 2], [-2.129, 2.129], False), ([-175, 175], [-175.129, 175.129], False),
                                       ([-175, 175], [-175.129, 175.129], True)):
        o = OceanDrift(loglevel=50)
        o.add_reader(list(readers(pacific)))
        o.set_config('general:use_auto_landmask', False)
        o.seed_elements(lon=seeds, lat=[60, 60], time=start_time, wind_drift_factor=.1)
        o.run(steps=2)
        np.testing.assert_array_almost_equal(o.elements.lon, expect_lon, decimal=3)
        np.testing.assert_array_almost_equal(o.elements.lat, [60.006, 59.994], decimal=3)
