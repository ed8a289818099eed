"""Known-answer tests the reference holds for the interpolation chain of this path
(/root/reference/tests/readers/test_interpolation.py), restated on the oracle port, on the host build of the device
sampler and (gpu) on od_interp.  SURVEY.md section 8(c) lists them:
  * :261-283 test_interpolation_vertical  - Linear1DInterpolator exact small cases, incl. extrapolation to surface / bottom
  * :216-229 test_flipped                 - Linear2DInterpolator on d = 5 X and on the same data with a flipped x axis
  * :150-182 synthetic block with holes   - NaN handling of linearNDFast (3x3 max-dilation, <= 10 passes), hole counts
Ensemble blocks (:231-259) are outside the hot path (no ensemble readers in the product)."""
import ctypes as C
from datetime import datetime, timedelta

import numpy as np
import pytest

import common
from common import hostshim, HsField

T0 = datetime(2026, 1, 1)
TIMES = [T0, T0 + timedelta(hours=1)]


# ---- the reference's numbers -----------------------------------------------------------------------------------------
VERT_CASES = [
    # zgrid, z, expected (test_interpolation.py:264-283); data[k] = k on every layer
    (np.array([0.0, 1.0, 3.0, 10.0]), np.array([0.5, 3.0, 9.0]), [0.5, 2.0, 2.85714286]),
    (np.array([1.0, 3.0, 5.0, 10.0]), np.array([0.5, 6.0, 12.0]), [0.0, 2.2, 3.0]),
]
FLIP_X0 = np.array([0, 7, 7.3, 7.41, 9])
FLIP_Y0 = np.array([5, 5, 8, 8.2, 5])


def _layers(zgrid):
    lon = np.array([0.0, 1.0, 2.0], np.float32)
    lat = np.array([0.0, 1.0, 2.0], np.float32)
    data = np.stack([np.full((3, 3), k, np.float32) for k in range(len(zgrid))])
    return lon, lat, data


def _flipped():
    x = np.arange(10).astype(np.float32)
    y = np.arange(20).astype(np.float32)
    X, _ = np.meshgrid(x, y)
    d = (X * 5).astype(np.float32)
    return x, y, d


def synthetic_block_with_holes():
    """get_synthetic_data_dict (test_interpolation.py:150-182), the 2-D slice on a grid scaled into valid lon/lat ranges
    (the index arithmetic only sees relative positions)."""
    xg, yg = np.meshgrid(np.linspace(-70, 470, 200), np.linspace(10, 340, 100))
    a = np.cos(np.radians(xg)) + np.sin(np.radians(yg))
    a[0:40, 50:60] = np.nan
    a[40:60, 100:120] = np.nan
    a[20:22, 30:32] = np.nan
    lon = np.linspace(1.0, 55.0, 200).astype(np.float32)
    lat = np.linspace(1.0, 34.0, 100).astype(np.float32)
    return lon, lat, a.astype(np.float32)


# ---- oracle port -----------------------------------------------------------------------------------------------------
def test_port_vertical_known_answers():
    from oracle import advect_port as ap
    for zgrid, z, expected in VERT_CASES:
        lon, lat, data = _layers(zgrid)
        r = ap.GridReader(lon, lat, zgrid, TIMES, {'upward_sea_water_velocity': np.stack([data, data])})
        n = len(z)
        env = ap.reader_interpolate(r, ['upward_sea_water_velocity'], T0, np.full(n, 1.0), np.full(n, 1.0), z.astype(np.float32))
        assert np.allclose(env['upward_sea_water_velocity'], expected)


def test_port_flipped_known_answers():
    from oracle import advect_port as ap
    x, y, d = _flipped()
    for xs, ds in ((x, d), (np.flip(x), np.flip(d, axis=1))):
        r = ap.GridReader(xs, y, None, TIMES, {'x_wind': np.stack([ds, ds]), 'y_wind': np.stack([ds, ds])})
        env = ap.reader_interpolate(r, ['x_wind'], T0, FLIP_X0.astype(np.float64), FLIP_Y0.astype(np.float64), np.zeros(5, np.float32))
        np.testing.assert_array_almost_equal(env['x_wind'], 5 * FLIP_X0)


def test_expand_array_hole_counts():
    """expand_numpy_array on the synthetic block: every pass removes the rim of each hole (test_interpolation.py:455-478
    checks such counts on a ROMS file that is not available here; the counts below are the port's own, i.e. the
    reference routine's, and pin the device fill)."""
    from oracle.advect_port import expand_numpy_array
    _, _, a = synthetic_block_with_holes()
    counts = [int(np.isnan(a).sum())]
    for _ in range(10):
        expand_numpy_array(a)
        counts.append(int(np.isnan(a).sum()))
    assert counts[0] == 40 * 10 + 20 * 20 + 4 and counts[1] < counts[0] and counts[-1] == 0
    assert counts == sorted(counts, reverse=True)


# ---- host build of the device sampler ---------------------------------------------------------------------------------
def test_host_sampler_vertical_and_flipped():
    lib = hostshim()
    for zgrid, z, expected in VERT_CASES:
        lon, lat, data = _layers(zgrid)
        f = HsField(lon, lat, zgrid, [[data, data]], TIMES, (float('nan'),))
        n = len(z)
        o0, _ = f.sample(lib, T0, np.full(n, 1.0), np.full(n, 1.0), z.astype(np.float32), False)
        assert np.allclose(o0, expected)
    x, y, d = _flipped()
    for xs, ds in ((x, d), (np.flip(x), np.flip(d, axis=1))):
        f = HsField(xs, y, None, [[ds, ds], [ds, ds]], TIMES, (float('nan'), float('nan')))
        o0, _ = f.sample(lib, T0, FLIP_X0.astype(np.float64), FLIP_Y0.astype(np.float64), np.zeros(5, np.float32), False)
        np.testing.assert_array_almost_equal(o0, 5 * FLIP_X0)


# ---- device ----------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_sampler_known_answers():
    from opendrift_b200.engine import Engine
    eng = Engine(0)
    for zgrid, z, expected in VERT_CASES:
        lon, lat, data = _layers(zgrid)
        grp = eng.add_group(lon, lat, zgrid, 1, TIMES, lambda ti, c: data, (float('nan'),))
        n = len(z)
        out = eng.interp(grp, T0, eng.to_device(np.full(n, 1.0)), eng.to_device(np.full(n, 1.0)), eng.to_device(z.astype(np.float32)), raw=True)
        assert np.allclose(out[0].cpu().numpy(), expected)
    x, y, d = _flipped()
    for xs, ds in ((x, d), (np.flip(x).copy(), np.flip(d, axis=1).copy())):
        grp = eng.add_group(xs, y, None, 2, TIMES, lambda ti, c: ds, (float('nan'), float('nan')))
        out = eng.interp(grp, T0, eng.to_device(FLIP_X0.astype(np.float64)), eng.to_device(FLIP_Y0.astype(np.float64)), None, raw=True)
        np.testing.assert_array_almost_equal(out[0].cpu().numpy(), 5 * FLIP_X0)
    eng.close()


@pytest.mark.gpu
def test_gpu_nan_fill_matches_expand_numpy_array():
    """od_group_fill_nan after k passes == k calls of the reference's expand_numpy_array: same cells, same values, same
    remaining-hole counts (incl. the report path of the C-ABI)."""
    from oracle.advect_port import expand_numpy_array
    from opendrift_b200.engine import Engine
    eng = Engine(0)
    lon, lat, a = synthetic_block_with_holes()
    for passes in (1, 2, 3, 10):
        grp = eng.add_group(lon, lat, None, 1, TIMES, lambda ti, c: a, (float('nan'),))
        grp.fill_nan = 0
        s = grp.slot_of(0)
        remaining = eng.fill_nan(grp.gid, s, 0, passes, report=True)
        ref = a.copy()
        for _ in range(passes):
            expand_numpy_array(ref)
        assert remaining == int(np.isnan(ref).sum())
        dev = eng.slot_tensor(grp, s, 0).cpu().numpy().reshape(ref.shape)
        assert np.array_equal(np.isnan(dev), np.isnan(ref))
        assert np.array_equal(dev[~np.isnan(dev)], ref[~np.isnan(ref)])
        eng.free_group(grp)
    eng.close()


def _north_to_south_case():
    rng = np.random.default_rng(0)
    lon = np.linspace(2, 5, 31).astype(np.float32)
    lat = np.linspace(60, 57, 41).astype(np.float32)            # north -> south, as many model grids are stored
    z = np.array([0, -5, -20, -50.0])
    u = rng.normal(size=(2, 4, 41, 31)).astype(np.float32)
    v = rng.normal(size=(2, 4, 41, 31)).astype(np.float32)
    n = 5000
    return lon, lat, z, u, v, rng.uniform(1.9, 5.1, n), rng.uniform(56.9, 60.1, n), rng.uniform(-60, 2, n).astype(np.float32)


def test_host_sampler_on_a_north_to_south_grid_is_bit_exact():
    from oracle import advect_port as ap
    lon, lat, z, u, v, plon, plat, pz = _north_to_south_case()
    t = T0 + timedelta(seconds=1234)
    env = ap.get_environment([ap.GridReader(lon, lat, z, TIMES, {common.CUR[0]: u, common.CUR[1]: v})], common.CUR, t, plon, plat, pz)
    f = HsField(lon, lat, z, [[u[0], u[1]], [v[0], v[1]]], TIMES, (0.0, 0.0))
    o0, o1 = f.sample(hostshim(), t, plon, plat, pz, False)
    assert np.array_equal(o0, env[common.CUR[0]]) and np.array_equal(o1, env[common.CUR[1]])
    assert (o0 == 0).sum() > 100          # the uncovered rim took the fallback


@pytest.mark.gpu
def test_gpu_sampler_on_a_north_to_south_grid_is_bit_exact():
    from oracle import advect_port as ap
    from opendrift_b200.engine import Engine
    lon, lat, z, u, v, plon, plat, pz = _north_to_south_case()
    t = T0 + timedelta(seconds=1234)
    env = ap.get_environment([ap.GridReader(lon, lat, z, TIMES, {common.CUR[0]: u, common.CUR[1]: v})], common.CUR, t, plon, plat, pz)
    eng = Engine(0)
    grp = eng.add_group(lon, lat, z, 2, TIMES, lambda ti, c: (u, v)[c][ti], (0.0, 0.0))
    o = eng.interp(grp, t, eng.to_device(plon), eng.to_device(plat), eng.to_device(pz))
    assert np.array_equal(o[0].cpu().numpy(), env[common.CUR[0]]) and np.array_equal(o[1].cpu().numpy(), env[common.CUR[1]])
    eng.close()


def test_modulate_longitude_known_answers():
    """tests/readers/test_variables.py:31-90 (test_modulate_longitude_360 / _180) on the product's reader base class."""
    from opendrift_b200.readers.basereader import StructuredReader

    class R(StructuredReader):
        def __init__(self, xmin, xmax):
            self.proj4 = '+proj=lonlat +ellps=WGS84'
            self.xmin, self.xmax, self.ymin, self.ymax = xmin, xmax, -80, 80
            self.variables = []
            self.name = 'domain'
            self.start_time = self.end_time = self.time_step = None
            super().__init__()

    r = R(0, 340)
    lons = np.linspace(0, 300, 100)
    assert (r.modulate_longitude(lons) == lons).all()
    assert (r.modulate_longitude(np.array([-180, -90])) == np.array([360 - 180, 360 - 90])).all()
    assert (r.modulate_longitude(np.array([0, 90])) == np.array([0, 90])).all()
    assert (r.modulate_longitude(np.array([100, 180])) == np.array([100, 180])).all()
    assert (r.modulate_longitude(np.array([240, 350])) == np.array([240, 350])).all()
    r = R(-150, 180)
    lons = np.linspace(-180, 150, 100)
    assert (r.modulate_longitude(lons) == lons).all()
    assert (r.modulate_longitude(np.array([-180, -90])) == np.array([-180, -90])).all()
    assert (r.modulate_longitude(np.array([0, 90])) == np.array([0, 90])).all()
    assert (r.modulate_longitude(np.array([100, 180])) == np.array([100, -180])).all()
    assert (r.modulate_longitude(np.array([240])) == np.array([-120])).all()


@pytest.mark.gpu
def test_leeway_forward_backward_symmetry():
    """tests/models/test_basemodel.py:96-146 (test_simulation_matches_forw_backward) without the automatic landmask:
    a forward run and a backward run with mirrored forcing end at the same positions.  (The reference test seeds
    FISHING-VESSEL-1 from its OBJECTPROP.DAT; the product ships the four PIW categories only, so PIW-4 stands in.)"""
    from opendrift_b200.models.leeway import Leeway

    def run(sign):
        lee = Leeway(loglevel=50)
        lee.set_config('general:use_auto_landmask', False)
        lee.set_config('environment:constant:land_binary_mask', 0)
        lee.set_config('environment:fallback:x_wind', -1.5 * sign)
        lee.set_config('environment:fallback:y_wind', -10 * sign)
        lee.set_config('environment:fallback:x_sea_water_velocity', -1.5 * sign)
        lee.set_config('environment:fallback:y_sea_water_velocity', 0)
        lee.seed_elements(lon=4.5, lat=60, number=100, object_type=4, time=datetime(2015, 1, 1))
        lee.run(steps=2, time_step=sign * 10 * 3600, time_step_output=sign * 10 * 3600)
        return lee
    leef, leeb = run(1), run(-1)
    assert leef.num_elements_active() == leeb.num_elements_active() == 100
    assert leef.num_elements_deactivated() == leeb.num_elements_deactivated()
    np.testing.assert_array_almost_equal(np.sort(leef.elements.lon), np.sort(leeb.elements.lon))
    np.testing.assert_array_almost_equal(np.sort(leef.elements.lat), np.sort(leeb.elements.lat), decimal=5)
    assert np.abs(leef.elements.lon - 4.5).max() > 0.05
