"""GPU parity for BASELINE configs[0]: the analytical double-gyre reader on its stereographic plane
(examples/example_double_gyre_advection_schemes.py) through the C-ABI (od_analytic_interp, od_analytic_advect) and through
the drop-in classes, against the fixtures the unmodified reference produced (tests/golden/ref_gyre_*.npz) and the port.

Tolerance: 1e-9 deg = 1.1e-4 m on the 2 m x 1 m box (see tests/test_gyre.py); the north-star 1e-6 deg is 0.11 m there.
(This file sorts last on purpose: it was written after the round's GPU minutes were spent)"""
from datetime import timedelta

import numpy as np
import pytest

import common
import gyre_common as gc
from opendrift_b200 import _lib

pytestmark = pytest.mark.gpu
TOL_M = 1.1e-4


def test_sampler_matches_port_on_gpu():
    from opendrift_b200.engine import default_engine
    eng = default_engine()
    fx = gc.GyreFixture('gyre_rk4')
    rd, pr = fx.product_reader(), fx.port_reader()
    rng = np.random.default_rng(0)
    n = 200000
    lon, lat = rd.xy2lonlat(rng.uniform(-0.02, 2.02, n), rng.uniform(-0.02, 1.02, n))
    d = rd.analytic_desc(with_fallback=False)
    for tsec, f32 in ((0.0, False), (1.35, False), (4.2, True)):
        lo = lon.astype(np.float32) if f32 else lon
        la = lat.astype(np.float32) if f32 else lat
        e = pr.interpolate(common.CUR, fx.t0 + timedelta(seconds=tsec), lo, la, None)
        u, v = eng.analytic_interp(d, tsec, eng.to_device(lo.astype(np.float64)), eng.to_device(la.astype(np.float64)), pos_f32=f32)
        u, v = u.cpu().numpy(), v.cpu().numpy()
        for got, k in ((u, common.CUR[0]), (v, common.CUR[1])):
            ref = e[k].astype(np.float32)
            assert np.array_equal(np.isnan(got), np.isnan(ref))
            ok = ~np.isnan(ref)
            # the device's libm is not the host's: a float64 result that differs in its last bits rounds to the
            # neighbouring float32 once in ~1e7 samples
            assert np.max(np.abs(got[ok] - ref[ok])) <= np.spacing(np.abs(ref[ok]).max())
            assert np.mean(got[ok] == ref[ok]) > 0.999


@pytest.mark.parametrize('mode', [_lib.OD_MATH_SERIES, _lib.OD_MATH_EXACT, _lib.OD_MATH_FAST])
@pytest.mark.parametrize('name', gc.gyre_fixtures())
def test_c_abi_matches_reference_fixture(name, mode):
    fx = gc.GyreFixture(name)
    lon, lat = gc.run_engine(fx, mode)
    assert gc.plane_error_m(fx, lon, lat, fx.lon, fx.lat) < TOL_M


@pytest.mark.parametrize('name', gc.gyre_fixtures())
def test_dropin_model_matches_reference_fixture(name):
    """The example script's calls on the drop-in classes."""
    fx = gc.GyreFixture(name)
    lon, lat = gc.run_model(fx)
    assert lon.dtype == np.float64
    assert gc.plane_error_m(fx, lon, lat, fx.lon, fx.lat) < TOL_M
    pl, pa = gc.run_port(fx)
    assert gc.plane_error_m(fx, lon, lat, pl, pa) < TOL_M


def test_large_particle_count_and_errors():
    """One launch over 2 M particles (grid sizing), and the library's argument checks."""
    import ctypes as C
    from opendrift_b200.engine import default_engine
    eng = default_engine()
    fx = gc.GyreFixture('gyre_rk4')
    rd = fx.product_reader()
    rd.bind(eng, {k: 0.0 for k in common.CUR})
    d = rd.analytic_desc()
    rng = np.random.default_rng(4)
    n = 2_000_000
    lon, lat = rd.xy2lonlat(rng.uniform(0.0, 2.0, n), rng.uniform(0.0, 1.0, n))
    dl, da = eng.to_device(lon), eng.to_device(lat)
    before = eng.launches()
    eng.analytic_advect(d, 'runge-kutta4', (0.0, 0.05, 0.1), 0.1, dl, da)
    eng.sync()
    assert eng.launches() == before + 1
    sub = slice(0, 3000)
    fx_small = gc.GyreFixture('gyre_rk4')
    from oracle import advect_port as ap
    # the port for the same single step on a subsample (float64 seeds: not the float32 first-step path)
    pr = fx_small.port_reader()
    env = ap.get_environment([pr], common.CUR, fx.t0, lon[sub], lat[sub], np.zeros(3000))
    pl, pa = ap.advect_ocean_current([pr], 'runge-kutta4', fx.t0, 0.1, lon[sub], lat[sub], np.zeros(3000), np.ones(3000),
                                     np.ones(3000, dtype=np.int32), env)
    k = gc.R_SPHERE * np.pi / 180
    assert np.max(np.hypot((dl.cpu().numpy()[sub] - pl) * k, (da.cpu().numpy()[sub] - pa) * k)) < 1e-7
    bad = rd.analytic_desc()
    bad.kind = 99
    with pytest.raises(RuntimeError, match='analytical reader kind'):
        eng.analytic_advect(bad, 'euler', (0.0, 0.05, 0.1), 0.1, dl, da)
    bad = rd.analytic_desc()
    bad.proj.a = -1.0
    with pytest.raises(RuntimeError, match='a > 0'):
        eng.analytic_interp(bad, 0.0, dl, da)


@pytest.mark.parametrize('proj4', gc.ASPECTS)
def test_other_projection_aspects_on_gpu(proj4):
    """Oblique / polar aspects and a plane across the dateline: C-ABI and drop-in classes against the port."""
    fx = gc.AspectCase(proj4)
    lon, lat = gc.run_engine(fx)
    assert fx.error_m(lon, lat) < 1e-5
    lon, lat = gc.run_model(fx)
    assert fx.error_m(lon, lat) < 1e-5
