"""The NumPy port (oracle/advect_port.py) is pinned to the UNMODIFIED reference: bit-identical final
positions on every committed fixture (tests/golden/ref_*.npz, written by oracle/make_golden.py from a
live run of /root/reference), and - when the reference tree is present - on a fresh random case."""
import numpy as np
import pytest

import common
from common import Fixture, fixtures, run_port


@pytest.mark.parametrize('name', fixtures())
def test_port_matches_reference_fixture(name):
    fx = Fixture(name)
    lon, lat, z = run_port(fx)
    assert np.array_equal(lon, fx.lon)
    assert np.array_equal(lat, fx.lat)
    assert np.array_equal(z, fx.z)
    assert np.abs(fx.lon - fx.lon0).max() > 1e-3        # the particles did move


@pytest.mark.parametrize('name', common.leeway_fixtures())
def test_leeway_port_matches_reference_fixture(name):
    fx = common.LeewayFixture(name)
    lon, lat, el = common.run_leeway_port(fx)
    assert np.array_equal(lon, fx.lon) and np.array_equal(lat, fx.lat)
    assert np.array_equal(el['orientation'], fx.orientation)
    assert np.array_equal(el['crosswind_slope'], fx.crosswind_slope)
    if fx.capsized is not None:                   # processes:capsizing: the same elements capsized
        assert np.array_equal(np.asarray(el['capsized'], dtype=np.float64), fx.capsized) and fx.capsized.sum() > 100


def test_port_matches_live_reference():
    from oracle import refrun
    if not refrun.available():
        pytest.skip('reference tree not present (GPU box)')
    from oracle import advect_port as ap
    from opendrift_b200 import synthetic as syn
    g = syn.GridSpec(nx=30, ny=28, nz=6, lon0=1.0, dlon=0.07, lat0=57.0, dlat=0.04, dz=15.0)
    times = syn.slab_times(3)
    U, V = zip(*[syn.double_gyre_uv(g, (t - syn.T0).total_seconds()) for t in times])
    fields = {common.CUR[0]: np.stack(U), common.CUR[1]: np.stack(V)}
    rng = np.random.default_rng(99)
    lon = rng.uniform(1.3, 2.8, 400).astype(np.float32)
    lat = rng.uniform(57.2, 58.0, 400).astype(np.float32)
    z = rng.uniform(-80, 0, 400).astype(np.float32)
    o = refrun.run_oceandrift([refrun.make_grid_reader(g.lon, g.lat, g.z, times, fields)], lon, lat, z, syn.T0,
                              900, 5, config={'drift:advection_scheme': 'runge-kutta4',
                                              'drift:vertical_advection': False})
    pl, pa, _ = ap.run_oceandrift([ap.GridReader(g.lon, g.lat, g.z, times, fields)], lon, lat, z, syn.T0, 900, 5)
    assert np.array_equal(pl, o.elements.lon) and np.array_equal(pa, o.elements.lat)
