"""The specialised step (opendrift_b200/csrc/od_spec.cuh) against the general step on the host build of the device sources: the
same runs with the specialisation on and off must agree BIT FOR BIT, and the specialised path must actually have been taken
(tests/hostshim counts its particles and the ones it flagged and handed back to the general step).  The cases are chosen to
hit what the specialised sampler flags: particles on and beyond the edge of the block, on the last row / column, longitudes
outside [-180, 180), steps on a reader time (time mode 1) and between two (mode 0), polar moves outside the series' range."""
import ctypes as C
from datetime import datetime, timedelta

import numpy as np
import pytest

import bigcases as bc
import common
from hostengine import HostEngine


@pytest.fixture()
def host_engine(monkeypatch):
    eng = HostEngine()
    import opendrift_b200.engine as E
    import opendrift_b200.models.basemodel as B
    monkeypatch.setattr(E, 'default_engine', lambda device=None: eng)
    monkeypatch.setattr(B, 'default_engine', lambda device=None: eng)
    yield eng
    common.hostshim().hs_spec_mode(1)


def _counts(reset=True):
    n, r = C.c_int64(0), C.c_int64(0)
    common.hostshim().hs_spec_counts(C.byref(n), C.byref(r), 1 if reset else 0)
    return n.value, r.value


def _run_pair(make):
    shim = common.hostshim()
    out = []
    for on in (1, 0):
        shim.hs_spec_mode(on)
        _counts()
        o = make()
        o.run(**o._test_run_args)
        out.append((np.asarray(o.elements.lon).copy(), np.asarray(o.elements.lat).copy(), np.asarray(o.elements.z).copy(),
                    np.asarray(o.elements.ID).copy(), _counts()))
    shim.hs_spec_mode(1)
    return out


@pytest.mark.parametrize('kind', ['cfg2', 'cfg4'])
def test_specialised_step_equals_general_step_on_benchmarked_configurations(kind, host_engine):
    if kind not in bc.KINDS:
        pytest.skip('fixture not generated')
    c = bc.BigCase(kind)
    sub = slice(0, 20000)

    def make():
        o = c.model(subset=sub, **{'gpu:sort_interval_steps': 0})
        o._test_run_args = dict(steps=c.steps, time_step=c.dt, time_step_output=c.steps * c.dt)
        return o

    (lon1, lat1, z1, id1, (n_on, redo_on)), (lon0, lat0, z0, id0, (n_off, _)) = _run_pair(make)
    assert n_off == 0 and n_on >= 20000 * (c.steps - 1), (n_on, n_off)      # (the first step has float32 positions: general step)
    assert redo_on < 0.02 * n_on
    assert np.array_equal(id1, id0)
    assert np.array_equal(lon1, lon0) and np.array_equal(lat1, lat0) and np.array_equal(z1, z0)


def _edge_model(lon_mode_pm180, seed, polar=False):
    """A small 3-D u/v/w block with particles scattered over and around it (edges, last row / column, outside), strong currents so
    that mid-points leave the block, run from a reader time (mode 1 at the first stage) over a reader time step."""
    from opendrift_b200.models.oceandrift import OceanDrift
    from opendrift_b200.readers import reader_regular_grid
    rng = np.random.default_rng(seed)
    nx, ny, nz = 24, 20, 6
    # (node coordinates that float32 holds exactly: the seeded positions are float32 until the first update)
    if polar:
        lon = -30.0 + 2.5 * np.arange(nx)
        lat = 86.0 + 0.1875 * np.arange(ny)
    elif lon_mode_pm180:
        lon = 174.0 + 0.25 * np.arange(nx)           # mid-points beyond 180 wrap: flagged
        lat = 58.0 + 0.125 * np.arange(ny)
    else:
        lon = 3.0 + 0.25 * np.arange(nx)
        lat = 58.0 + 0.125 * np.arange(ny)
    z = -np.array([0.0, 5.0, 10.0, 25.0, 50.0, 100.0])
    t0 = datetime(2024, 3, 1)
    times = [t0 + timedelta(hours=h) for h in range(3)]
    sc = 40.0 if polar else 3.0
    u = (sc * rng.standard_normal((3, nz, ny, nx))).astype(np.float32)
    v = (sc * rng.standard_normal((3, nz, ny, nx))).astype(np.float32)
    w = (0.01 * rng.standard_normal((3, nz, ny, nx))).astype(np.float32)
    u[:, :, 3:5, 7:9] = np.nan                            # a land hole (NaN fill)
    v[:, :, 3:5, 7:9] = np.nan
    for a in (u, v):                                      # no flow on the rim: particles seeded on it stay on it
        a[:, :, 0, :] = a[:, :, -1, :] = 0.0
        a[:, :, :, 0] = a[:, :, :, -1] = 0.0
    fields = {'x_sea_water_velocity': u, 'y_sea_water_velocity': v, 'upward_sea_water_velocity': w}
    o = OceanDrift(loglevel=50, seed=0)
    o.add_reader(reader_regular_grid.Reader(lon, lat, z, times, fields, name='current'))
    o.set_config('general:use_auto_landmask', False)
    o.set_config('general:coastline_action', 'none')
    o.set_config('drift:advection_scheme', 'runge-kutta4')
    o.set_config('drift:vertical_advection', True)
    o.set_config('drift:vertical_mixing', False)
    o.set_config('environment:fallback:x_sea_water_velocity', 0.3)
    o.set_config('environment:fallback:y_sea_water_velocity', -0.2)
    o.set_config('gpu:sort_interval_steps', 0)
    n = 4000
    dlon, dlat = lon[1] - lon[0], lat[1] - lat[0]
    plon = rng.uniform(lon[0] - 2 * dlon, lon[-1] + 2 * dlon, n)
    plat = rng.uniform(lat[0] - 2 * dlat, min(lat[-1] + 2 * dlat, 89.99), n)
    # exactly on nodes, on the first / last row and column
    plon[:200] = rng.choice(lon, 200)
    plat[:200] = rng.choice(lat, 200)
    plon[200:300] = lon[-1]
    plat[300:400] = lat[-1]
    plon[400:450] = lon[0]
    plat[450:500] = lat[0]
    pz = -rng.uniform(0, 120, n)
    pz[::7] = 0.0
    o.seed_elements(lon=plon, lat=plat, z=pz, time=t0)
    o._test_run_args = dict(steps=9, time_step=600, time_step_output=9 * 600)
    return o


@pytest.mark.parametrize('pm180,polar,seed', [(False, False, 1), (True, False, 2), (False, True, 3)])
def test_specialised_step_equals_general_step_on_edges(pm180, polar, seed, host_engine):
    (lon1, lat1, z1, id1, (n_on, redo_on)), (lon0, lat0, z0, id0, (n_off, _)) = _run_pair(lambda: _edge_model(pm180, seed, polar))
    assert n_off == 0 and n_on > 0
    assert 0 < redo_on < n_on                      # some were flagged, not all
    assert np.array_equal(id1, id0)
    assert np.array_equal(lon1, lon0, equal_nan=True) and np.array_equal(lat1, lat0, equal_nan=True) and np.array_equal(z1, z0, equal_nan=True)
