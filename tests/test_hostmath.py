"""The per-particle device math (opendrift_b200/csrc/*.cuh) compiled for the host by tests/hostshim and
checked against the oracle - the arithmetic of the CUDA kernels, verified without a GPU.

Tolerances: float64-only paths (geodesic, Euler with float64 factors) agree to round-off; paths with a
float32 arctan2 (RK mid-points, float32 final move) differ from NumPy's own (non correctly rounded,
SIMD-dependent) float32 arctan2 by an ulp of azimuth, i.e. ~1e-9 deg of position per step."""
import ctypes as C
import os

import numpy as np
import pytest

import common
from datetime import timedelta
from opendrift_b200 import synthetic as syn

from common import Fixture, fixtures, run_hostshim, hostshim, _p, GOLDEN


def test_geodesic_vs_exact_integrals():
    lib = hostshim()
    g = np.load(os.path.join(GOLDEN, 'geod_mpmath.npz'))
    n = len(g['lon1'])
    lo, la = np.empty(n), np.empty(n)
    lib.hs_geod_direct(C.c_int64(n), _p(g['lon1']), _p(g['lat1']), _p(g['azi1']), _p(g['s12']), _p(lo), _p(la))
    dlon = (lo - g['lon2'] + 180.0) % 360.0 - 180.0
    assert np.abs(la - g['lat2']).max() < 1e-12
    assert np.abs(dlon * np.cos(np.radians(g['lat2']))).max() < 1e-12


@pytest.mark.parametrize('name', fixtures())
def test_step_vs_reference_fixture(name):
    fx = Fixture(name)
    lon, lat, z = run_hostshim(fx)
    elon, elat = common.max_err_deg(lon, lat, fx.lon, fx.lat)
    tol = 1e-11 if fx.meta['scheme'] == 'euler' and fx.cdf is None else 2e-8
    assert elon < tol and elat < tol, (elon, elat)
    assert np.abs(z - fx.z).max() <= common.z_tolerance(fx.meta)


@pytest.mark.parametrize('name', fixtures())
def test_series_mode_vs_reference_fixture(name):
    """OD_MATH_SERIES (the engine's default): same tolerance as the exact replay."""
    fx = Fixture(name)
    lon, lat, z = run_hostshim(fx, fast=2)
    elon, elat = common.max_err_deg(lon, lat, fx.lon, fx.lat)
    tol = 1e-11 if fx.meta['scheme'] == 'euler' and fx.cdf is None else 2e-8
    assert elon < tol and elat < tol, (elon, elat)
    assert np.abs(z - fx.z).max() <= common.z_tolerance(fx.meta)


def _wrap(d):
    return d - 360.0 * np.round(d / 360.0)


def test_series_geodesic_vs_exact_integrals():
    """The short-arc series against the mpmath evaluation of the geodesic integrals, where it applies."""
    lib = hostshim()
    g = np.load(os.path.join(GOLDEN, 'geod_mpmath.npz'))
    n = len(g['lon1'])
    xn = g['s12'] * np.cos(np.radians(g['azi1']))
    ye = g['s12'] * np.sin(np.radians(g['azi1']))
    lo, la, used = np.empty(n), np.empty(n), np.zeros(n, dtype=np.int32)
    lib.hs_geod_series(C.c_int64(n), _p(g['lon1']), _p(g['lat1']), _p(xn), _p(ye), _p(lo), _p(la), _p(used))
    m = used == 1
    assert m.sum() > 300                      # the golden set holds many step-sized cases
    err_lon = np.abs(_wrap(lo - g['lon2']) * np.cos(np.radians(g['lat2'])))
    assert np.abs(la - g['lat2'])[m].max() < 1e-13 and err_lon[m].max() < 1e-13
    assert np.abs(la - g['lat2']).max() < 1e-12 and err_lon.max() < 1e-12       # fallback = the full solution


def test_series_geodesic_vs_karney_random():
    """Series vs the oracle's full Karney solution on random step-sized moves; the series must take every move of
    an ordinary drift step (<= 10 km below 80 degrees of latitude) and hand long or polar moves to the full solution."""
    from oracle import geod_karney as gk
    lib = hostshim()
    rng = np.random.default_rng(11)
    n = 200000
    lon = rng.uniform(-180, 180, n)
    az = rng.uniform(-180, 180, n)
    for lat_max, s_max, all_series in ((60.0, 10000.0, True), (80.0, 3000.0, True), (89.99, 50000.0, False)):
        lat = rng.uniform(-lat_max, lat_max, n)
        s = rng.uniform(0.0, s_max, n)
        xn, ye = s * np.cos(np.radians(az)), s * np.sin(np.radians(az))
        lo, la, used = np.empty(n), np.empty(n), np.zeros(n, dtype=np.int32)
        lib.hs_geod_series(C.c_int64(n), _p(lon), _p(lat), _p(xn), _p(ye), _p(lo), _p(la), _p(used))
        l2, a2 = gk.direct(lon, lat, az, s)
        assert np.abs(la - a2).max() < 1e-13
        assert np.abs(_wrap(lo - l2) * np.cos(np.radians(a2))).max() < 1e-13
        if all_series:
            assert used.all()
        else:
            assert 0 < used.sum() < n
    # degenerate inputs: zero-length move returns the start point; NaN propagates; a pole start falls back
    lon1 = np.array([10.0, 10.0, 10.0, -179.9999999]); lat1 = np.array([60.0, 60.0, 90.0, 0.0])
    xn = np.array([0.0, np.nan, -100.0, 0.0]); ye = np.array([0.0, 1.0, 0.0, -50.0])
    lo, la, used = np.empty(4), np.empty(4), np.zeros(4, dtype=np.int32)
    lib.hs_geod_series(C.c_int64(4), _p(lon1), _p(lat1), _p(xn), _p(ye), _p(lo), _p(la), _p(used))
    assert lo[0] == 10.0 and la[0] == 60.0 and used[0] == 1
    assert np.isnan(lo[1]) and np.isnan(la[1]) and used[1] == 0
    assert used[2] == 0 and abs(la[2] - (90.0 - 100.0 / 111693.9)) < 1e-6
    assert used[3] == 1 and 179.999 < lo[3] <= 180.0 and la[3] == 0.0       # crosses the date line westwards


@pytest.mark.parametrize('name', fixtures())
def test_fast_mode_vs_reference_fixture(name):
    fx = Fixture(name)
    lon, lat, z = run_hostshim(fx, fast=True)
    elon, elat = common.max_err_deg(lon, lat, fx.lon, fx.lat)
    assert elon < 1e-7 and elat < 1e-7, (elon, elat)


@pytest.mark.parametrize('name', common.leeway_fixtures())
def test_leeway_step_vs_reference_fixture(name):
    fx = common.LeewayFixture(name)
    lon, lat, el = common.run_leeway_hostshim(fx)
    e = common.max_err_deg(lon, lat, fx.lon, fx.lat)
    assert max(e) < 5e-8, e                       # float32 sin / cos / arctan2 differ from NumPy's SIMD versions by an ulp
    assert np.array_equal(el['orientation'], fx.orientation)
    assert np.array_equal(el['crosswind_slope'], fx.crosswind_slope)
    if fx.capsized is not None:                   # processes:capsizing: the same elements capsized
        assert np.array_equal(np.asarray(el['capsized'], dtype=np.float64), fx.capsized) and fx.capsized.sum() > 100


def test_interpolation_bit_exact():
    """od_interp arithmetic == ReaderBlock/Linear2DInterpolator/Linear1DInterpolator/time lerp/float32 cast."""
    from datetime import timedelta
    from oracle import advect_port as ap
    lib = hostshim()
    fx = Fixture('rk4_3d_offgrid')
    r = ap.GridReader(fx.grid_lon, fx.grid_lat, fx.grid_z, fx.times, {common.CUR[0]: fx.u, common.CUR[1]: fx.v})
    cur = common.HsField(fx.grid_lon, fx.grid_lat, fx.grid_z, [fx.u, fx.v], fx.times)
    rng = np.random.default_rng(3)
    n = 20000
    lon = rng.uniform(fx.grid_lon.min() - 0.2, fx.grid_lon.max() + 0.2, n)
    lat = rng.uniform(fx.grid_lat.min() - 0.1, fx.grid_lat.max() + 0.1, n)
    lon[:50] = fx.grid_lon[-1]              # exactly on the last column / row
    lat[50:100] = fx.grid_lat[-1]
    lon[100:150] = fx.grid_lon[0]
    z = rng.uniform(fx.grid_z.min() * 1.2, 5.0, n).astype(np.float32)
    for off, pos32 in [(0, 0), (1234, 0), (3600, 0), (5000, 0), (1234, 1)]:
        t = fx.times[0] + timedelta(seconds=off)
        lo, la = (lon.astype(np.float32), lat.astype(np.float32)) if pos32 else (lon, lat)
        env = ap.get_environment([r], common.CUR, t, lo, la, z)
        o0, o1 = np.empty(n, np.float32), np.empty(n, np.float32)
        pr = cur.pair(t)
        lib.hs_interp(C.byref(cur.g), C.byref(pr), C.c_int64(n), _p(lo.astype(np.float64)),
                      _p(la.astype(np.float64)), _p(z), C.c_int(pos32), _p(o0), _p(o1))
        assert np.array_equal(o0, env[common.CUR[0]]), (off, pos32)
        assert np.array_equal(o1, env[common.CUR[1]]), (off, pos32)
        assert (o0 == 0).sum() > 100          # some samples fell outside coverage -> fallback 0


def test_reciprocal_division_is_correctly_rounded():
    """The device forms (x - x0) / xspan as q0 = RN(d r), e = fma(-q0, s, d), q = fma(e, r, q0) with r = RN(1/s)
    (od_interp.cuh div_rn).  Emulate the three correctly rounded operations with exact rationals and compare with the
    IEEE quotient the reference computes, for the spans of the bench grid, of the fixtures and for random spans."""
    from fractions import Fraction
    from opendrift_b200 import synthetic as syn
    rng = np.random.default_rng(21)
    g = syn.GridSpec()
    spans = [float(np.float32(g.lon[-1] - g.lon[0])), float(np.float32(g.lat[-1] - g.lat[0]))]
    fx = Fixture('rk4_3d')
    spans += [float(np.float32(fx.grid_lon[-1] - fx.grid_lon[0])), float(np.float32(fx.grid_lat[-1] - fx.grid_lat[0]))]
    spans += [float(np.float32(s)) for s in rng.uniform(0.01, 360.0, 20)] + list(rng.uniform(1e-3, 400.0, 20))
    bad = 0
    for s in spans:
        r = 1.0 / s
        fs, fr = Fraction(s), Fraction(r)
        for d in np.concatenate([rng.uniform(-1.0, 1.2, 1500) * s, rng.uniform(-400, 400, 500), [0.0, s, -s, 1e-249, -3e-200]]):            # |d| < 1e-250 takes the true division
            d = float(d)
            q0 = d * r
            e = float(Fraction(d) - Fraction(q0) * fs)            # exact: the FMA rounds once, and this residual is representable
            q = float(Fraction(q0) + Fraction(e) * fr)
            bad += q != d / s
    assert bad == 0


@pytest.mark.parametrize('seed', range(12))
def test_interpolation_bit_exact_on_random_geometries(seed):
    """The sampler against the port on random block geometries: grid sizes down to 2 x 2 x 2, ascending and descending
    longitude / latitude / depth axes, irregular depth levels, longitude conventions on both sides of Greenwich and
    east-west periodic grids, 1- and 2-component groups, float64 and float32 positions, particles on and off the grid."""
    from oracle import advect_port as ap
    lib = hostshim()
    rng = np.random.default_rng(1000 + seed)
    nx, ny = int(rng.integers(2, 40)), int(rng.integers(2, 40))
    nz = int(rng.choice([1, 2, 3, 7, 12]))
    periodic = seed % 4 == 3
    if periodic:
        nx = int(rng.choice([36, 72, 90]))
        dx = 360.0 / nx
        lon = (rng.choice([0.0, -180.0]) + dx * np.arange(nx)).astype(np.float32)
    else:
        x0 = rng.uniform(-170, 170) if seed % 2 else rng.uniform(1, 300)
        dx = rng.uniform(0.01, 0.5)
        lon = (x0 + dx * np.arange(nx)).astype(np.float32)
        if rng.uniform() < 0.3:
            lon = lon[::-1].copy()
    lat = (rng.uniform(-80, 60) + rng.uniform(0.01, 0.4) * np.arange(ny)).astype(np.float32)
    if rng.uniform() < 0.4:
        lat = lat[::-1].copy()
    z = None
    if nz > 1:
        # levels representable in float32: the reference clamps a float32 copy of z to the float64 level range and scipy's
        # interp1d raises when the rounded value lands outside it (interpolators.py:176-183), so such grids cannot be replayed
        z = (-np.cumsum(rng.uniform(0.5, 20.0, nz)) + rng.uniform(0, 3)).astype(np.float32).astype(np.float64)
        if rng.uniform() < 0.5:
            z = z[::-1].copy()
    times = [common.syn.T0 + timedelta(hours=i) for i in range(3)]
    shape = (3, nz, ny, nx) if nz > 1 else (3, ny, nx)
    ncomp = 1 if seed % 3 == 0 else 2
    names = ['upward_sea_water_velocity'] if ncomp == 1 else list(common.CUR)
    fields = [rng.normal(size=shape).astype(np.float32) for _ in range(ncomp)]
    r = ap.GridReader(lon, lat, z, times, dict(zip(names, fields)))
    f = common.HsField(lon, lat, z, fields, times, tuple([0.0] * ncomp))
    assert bool(f.g.wrap_x) == periodic
    n = 4000
    plon = rng.uniform(float(lon.min()) - 2 * abs(dx), float(lon.max()) + 2 * abs(dx), n)
    plat = rng.uniform(float(lat.min()) - 0.3, float(lat.max()) + 0.3, n)
    plon[:20], plat[20:40] = float(lon[-1]), float(lat[-1])             # on the last column / row
    plon[40:60], plat[60:80] = float(lon[0]), float(lat[0])
    pz = (rng.uniform(float(z.min()) - 5, 3.0, n) if z is not None else np.zeros(n)).astype(np.float32)
    for off, pos32 in ((0, 0), (1800, 0), (4321, 0), (4321, 1)):
        t = times[0] + timedelta(seconds=off)
        lo, la = (plon.astype(np.float32), plat.astype(np.float32)) if pos32 else (plon, plat)
        env = ap.get_environment([r], names, t, lo, la, pz)
        outs = f.sample(lib, t, lo.astype(np.float64), la.astype(np.float64), pz, bool(pos32))
        for k, nme in enumerate(names):
            assert np.array_equal(outs[k], env[nme]), (seed, off, pos32, nme)


# ---- random whole-step scenarios: the device math (host build) against the port ---------------------------------------
class _Fx:
    pass

def _random_scenario(seed):
    rng=np.random.default_rng(5000+seed)
    fx=_Fx()
    nx, ny = int(rng.integers(8, 40)), int(rng.integers(8, 40))
    nz = int(rng.choice([1, 3, 7]))
    periodic = seed % 5 == 4
    if periodic:
        nx = 72; dx=5.0
        lon=(rng.choice([0.0,-180.0]) + dx*np.arange(nx)).astype(np.float32)
        lat=(rng.uniform(-60,20)+0.5*np.arange(ny)).astype(np.float32)
    else:
        dx=rng.uniform(0.02,0.2)
        lon=(rng.uniform(-170,150)+dx*np.arange(nx)).astype(np.float32)
        lat=(rng.uniform(-70,60)+rng.uniform(0.02,0.1)*np.arange(ny)).astype(np.float32)
        if rng.uniform()<0.3: lon=lon[::-1].copy()
    if rng.uniform()<0.4: lat=lat[::-1].copy()
    z=None
    if nz>1:
        z=(-np.cumsum(rng.uniform(1,15,nz))+rng.uniform(0,1)).astype(np.float32).astype(np.float64)
        z[0]=0.0 if rng.uniform()<0.5 else z[0]
        if rng.uniform()<0.5: z=z[::-1].copy()
    steps=int(rng.integers(3,7)); dt=int(rng.choice([300,600,900,-600]))
    nsl=syn.n_slabs_for(steps,dt)+1
    shape=(nsl,nz,ny,nx) if nz>1 else (nsl,ny,nx)
    amp=0.5
    fx.u=(amp*rng.normal(size=shape)).astype(np.float32); fx.v=(amp*rng.normal(size=shape)).astype(np.float32)
    with_w = nz>1 and rng.uniform()<0.5
    fx.w=(1e-3*rng.normal(size=shape)).astype(np.float32) if with_w else None
    wind = rng.uniform()<0.5
    fx.x_wind=(5*rng.normal(size=(nsl,ny,nx))).astype(np.float32) if wind else None
    fx.y_wind=(5*rng.normal(size=(nsl,ny,nx))).astype(np.float32) if wind else None
    fx.wind_lon, fx.wind_lat = lon, lat
    fx.grid_lon, fx.grid_lat, fx.grid_z = lon, lat, z
    fx.cdf=None; fx.kdiff=None; fx.stokes=None
    n=800
    fx.lon0=rng.uniform(float(lon.min())-abs(dx), float(lon.max())+abs(dx), n).astype(np.float32)
    fx.lat0=rng.uniform(float(lat.min())-0.05, float(lat.max())+0.05, n).astype(np.float32)
    fx.z0=(rng.uniform((z.min()-3) if z is not None else -1, 0.5, n)).astype(np.float32)
    fx.z0[fx.z0>0]=0
    if z is None: fx.z0[:]=0
    fx.times=syn.slab_times(nsl,3600)
    fx.dt=dt; fx.steps=steps; fx.n=n
    fx.start=fx.times[-1] if dt<0 else syn.T0+timedelta(seconds=int(rng.choice([0,450])))
    scheme=str(rng.choice(['euler','runge-kutta','runge-kutta4']))
    fx.meta=dict(scheme=scheme,with_w=with_w,wind=wind,diffusivity=0.0,seed=0,wind_drift_depth=None,mixing=False,dt_mix=60.0,stokes=None,noise=None,start_offset_s=0)
    fx.props=lambda: (np.float32(1)*np.ones(n), np.float32(0.02)*np.ones(n), np.ones(n,dtype=np.int32))
    fx.wind_drift_depth=lambda: 0.1
    return fx



@pytest.mark.parametrize('seed', range(10))
def test_random_step_scenarios_vs_port(seed):
    """Random block geometry (ascending / descending axes, periodic grids, a level at z = 0 or not), scheme, time step
    (also backward), start offset, wind drift and vertical advection; particles inside and outside the block: the device
    math must follow the port (= the reference) to the same 2e-8 deg as on the fixtures, in both float64 arithmetic modes."""
    fx = _random_scenario(seed)
    pl, pa, pz = common.run_port(fx)
    for mode in (0, 2, 1):          # exact replay, default (series), fast (float32 sampling + series moves)
        hl, ha, hz = run_hostshim(fx, fast=mode)
        assert np.array_equal(np.isfinite(pl), np.isfinite(hl))
        m = np.isfinite(pl)
        e = common.max_err_deg(hl[m], ha[m], pl[m], pa[m])
        assert max(e) < (1e-7 if mode == 1 else 2e-8), (seed, mode, e)     # (measured: fast <= 9e-9 over 120 scenarios up to 86 N)
        assert np.nanmax(np.abs(hz.astype(float) - pz.astype(float))) <= 1e-5


def _random_physics_scenario(seed):
    rng=np.random.default_rng(9000+seed)
    fx=_random_scenario(seed+1000)
    n=fx.n
    three_d = fx.grid_z is not None
    nsl=fx.u.shape[0]; ny,nx=len(fx.grid_lat),len(fx.grid_lon)
    m=fx.meta
    if fx.dt<0:
        return None
    # always wind for stokes / analytic mixing
    if fx.x_wind is None:
        fx.x_wind=(5*rng.normal(size=(nsl,ny,nx))).astype(np.float32); fx.y_wind=(5*rng.normal(size=(nsl,ny,nx))).astype(np.float32)
    m['wind']=True
    choice=seed%4
    if choice==0 and three_d:      # env mixing
        nz=len(fx.grid_z)
        fx.kdiff=(0.01*rng.uniform(0.1,1.0,size=(nsl,nz,ny,nx))).astype(np.float32)
        m['mixing']=True; m['dt_mix']=float(rng.choice([60.0,100.0]))
    elif choice==1:                # analytic mixing
        m['mixing']=True; m['dt_mix']=60.0; m['diffusivity_model']=str(rng.choice(['windspeed_Large1994','windspeed_Sundby1983'])); m['background_diffusivity']=float(rng.choice([1.2e-5,1e-4]))
    elif choice==2:                # stokes
        prof=str(rng.choice(['Phillips','monochromatic','exponential']))
        fx.stokes={'sea_surface_wave_stokes_drift_x_velocity':(0.1*rng.normal(size=(nsl,ny,nx))).astype(np.float32),
                   'sea_surface_wave_stokes_drift_y_velocity':(0.1*rng.normal(size=(nsl,ny,nx))).astype(np.float32)}
        if rng.uniform()<0.5:
            fx.stokes['sea_surface_wave_significant_height']=(1+rng.uniform(size=(nsl,ny,nx))).astype(np.float32)
        m['stokes']=prof
    else:
        m['diffusivity']=float(rng.choice([1.0,10.0]))
    return fx



@pytest.mark.parametrize('seed', [1, 2, 3, 5, 6, 7, 18, 42, 46])
def test_random_physics_scenarios_vs_port(seed):
    """As above with the analytical mixing models, a Stokes profile (with vertical advection in the same step: the Stokes
    move must see the start-of-step depth) or horizontal diffusion switched on."""
    fx = _random_physics_scenario(seed)
    if fx is None:
        pytest.skip('backward run or 2-D block drawn for this seed')
    pl, pa, pz = common.run_port(fx)
    for mode in (0, 2):
        hl, ha, hz = run_hostshim(fx, fast=mode)
        assert np.array_equal(np.isfinite(pl), np.isfinite(hl))
        m = np.isfinite(pl)
        assert max(common.max_err_deg(hl[m], ha[m], pl[m], pa[m])) < 2e-8, (seed, mode)
        assert np.nanmax(np.abs(hz.astype(float) - pz.astype(float))) <= (1e-6 if fx.meta.get('mixing') else 1e-5)


class _L:
    pass
def _random_leeway_scenario(seed):
    rng=np.random.default_rng(20000+seed)
    b=_random_scenario(seed+7000)
    fx=_L()
    fx.grid_lon,fx.grid_lat=b.grid_lon,b.grid_lat
    nsl=b.u.shape[0]; ny,nx=len(b.grid_lat),len(b.grid_lon)
    fx.u=(0.3*rng.normal(size=(nsl,ny,nx))).astype(np.float32); fx.v=(0.3*rng.normal(size=(nsl,ny,nx))).astype(np.float32)
    fx.x_wind=(8*rng.normal(size=(nsl,ny,nx))).astype(np.float32); fx.y_wind=(8*rng.normal(size=(nsl,ny,nx))).astype(np.float32)
    n=600
    lo,hi=float(fx.grid_lon.min()),float(fx.grid_lon.max()); la0,la1=float(fx.grid_lat.min()),float(fx.grid_lat.max())
    fx.lon0=rng.uniform(lo+0.2*(hi-lo), hi-0.2*(hi-lo), n).astype(np.float32)
    fx.lat0=rng.uniform(la0+0.2*(la1-la0), la1-0.2*(la1-la0), n).astype(np.float32)
    fx.times=b.times; fx.dt=abs(b.dt); fx.steps=b.steps; fx.n=n; fx.start=syn.T0
    caps = [8.0,5.0] if seed%2 else None
    from opendrift_b200.models.leeway import read_object_properties
    prop=read_object_properties()[int(rng.choice([1,2,3,4]))]
    fx.prop={k:v for k,v in prop.items() if k not in('OBJKEY','Description')}
    fx.meta=dict(seed=int(seed),capsizing=caps)
    fx.capsized=None
    return fx


@pytest.mark.parametrize('seed', range(6))
def test_random_leeway_scenarios_vs_port(seed):
    fx = _random_leeway_scenario(seed)
    pl, pa, pel = common.run_leeway_port(fx)
    hl, ha, hel = common.run_leeway_hostshim(fx)
    assert np.array_equal(np.isfinite(pl), np.isfinite(hl))
    m = np.isfinite(pl)
    assert max(common.max_err_deg(hl[m], ha[m], pl[m], pa[m])) < 5e-8
    assert np.array_equal(hel['orientation'], pel['orientation']) and np.array_equal(hel['crosswind_slope'], pel['crosswind_slope'])
    if fx.meta['capsizing'] is not None:
        assert np.array_equal(np.asarray(hel['capsized'], dtype=np.float64), pel['capsized'])


def _random_options_scenario(seed):
    rng=np.random.default_rng(30000+seed)
    fx=_random_scenario(seed+5000)
    m=fx.meta; n=fx.n
    three_d=fx.grid_z is not None
    if rng.uniform()<0.4 and three_d:
        m['truncate']=float(rng.uniform(2, abs(fx.grid_z).max()))
    if m['with_w'] and rng.uniform()<0.5: m['w_at_surface']=True
    if m['wind']:
        wdd=float(rng.choice([0.0,0.1,0.5,2.0])); m['wind_drift_depth']=wdd
        fx.wind_drift_depth=(lambda w=wdd: w)
    if rng.uniform()<0.3:
        cdf=rng.uniform(0.3,1.2,n).astype(np.float32); fx.cdf=cdf
        fx.props=(lambda c=cdf,n=n: (c.astype(np.float32), np.float32(0.02)*np.ones(n), np.ones(n,dtype=np.int32)))
    # particles exactly on nodes / levels
    k=40
    ii=rng.integers(1,len(fx.grid_lon)-1,k); jj=rng.integers(1,len(fx.grid_lat)-1,k)     # interior nodes (a motionless particle exactly on the boundary is a 1-ulp knife edge of the geodesic)
    fx.lon0[:k]=fx.grid_lon[ii]; fx.lat0[:k]=fx.grid_lat[jj]
    if three_d:
        kk=rng.integers(0,len(fx.grid_z),k); fx.z0[k:2*k]=np.minimum(fx.grid_z[kk],0).astype(np.float32)
    # land holes
    if rng.uniform()<0.4:
        ny,nx=len(fx.grid_lat),len(fx.grid_lon)
        iy,ix=np.meshgrid(np.arange(ny),np.arange(nx),indexing='ij')
        cx,cy=rng.integers(0,nx),rng.integers(0,ny); hole=(ix-cx)**2+(iy-cy)**2<=rng.integers(2,20)
        for a in (fx.u,fx.v):
            if a.ndim==4:
                a[:,:,hole]=np.nan
                a[:,a.shape[1]//2:,(ix-cx-3)**2+(iy-cy)**2<=9]=np.nan
            else: a[:,hole]=np.nan
    if rng.uniform()<0.25:
        m['noise']={'current':0.1} if rng.uniform()<0.5 else {'current_uniform':0.05,'wind':0.5}
    return fx


@pytest.mark.parametrize('seed', [0, 3, 6, 17, 19, 21, 27, 31])
def test_random_option_scenarios_vs_port(seed):
    """Random scenarios with drift:truncate_ocean_model_below_m, vertical advection at the surface, wind-drift depths,
    per-element float32 drift factors, uncertainty draws, land holes (NaN fill) and particles exactly on interior grid
    nodes and on depth levels."""
    fx = _random_options_scenario(seed)
    pl, pa, pz = common.run_port(fx)
    for mode in (0, 2):
        hl, ha, hz = run_hostshim(fx, fast=mode)
        assert np.array_equal(np.isfinite(pl), np.isfinite(hl))
        m = np.isfinite(pl)
        assert max(common.max_err_deg(hl[m], ha[m], pl[m], pa[m])) < 2e-8, (seed, mode)
        assert np.nanmax(np.abs(hz.astype(float) - pz.astype(float))) <= 1e-5


def _random_readers_scenario(seed):
    rng=np.random.default_rng(50000+seed)
    fx=_random_scenario(seed+11000)
    if fx.grid_z is None or fx.dt<0: return None
    nsl=fx.u.shape[0]
    # w on its own (coarser / finer, differently ordered) grid covering the same box, different levels
    lo,hi=float(fx.grid_lon.min()),float(fx.grid_lon.max()); la0,la1=float(fx.grid_lat.min()),float(fx.grid_lat.max())
    nxw,nyw,nzw=int(rng.integers(5,30)),int(rng.integers(5,30)),int(rng.choice([2,4,9]))
    fx.w_lon=np.linspace(lo+0.01,hi-0.01,nxw).astype(np.float32); fx.w_lat=np.linspace(la0-0.01,la1+0.01,nyw).astype(np.float32)
    if rng.uniform()<0.5: fx.w_lat=fx.w_lat[::-1].copy()
    zmin=float(fx.grid_z.min())
    fx.w_z=np.linspace(0.0,zmin*rng.uniform(0.6,1.2),nzw).astype(np.float32).astype(np.float64)
    if rng.uniform()<0.5: fx.w_z=fx.w_z[::-1].copy()
    fx.w=(2e-3*rng.normal(size=(nsl,nzw,nyw,nxw))).astype(np.float32)
    fx.meta['with_w']=True
    if rng.uniform()<0.5: fx.meta['w_at_surface']=True
    # wind on a global periodic grid
    if seed%2:
        fx.wind_lon=(rng.choice([0.0,-180.0])+5.0*np.arange(72)).astype(np.float32); fx.wind_lat=np.linspace(-88,88,45).astype(np.float32)
        fx.x_wind=(6*rng.normal(size=(nsl,45,72))).astype(np.float32); fx.y_wind=(6*rng.normal(size=(nsl,45,72))).astype(np.float32)
        fx.meta['wind']=True
    return fx


@pytest.mark.parametrize('seed', [0, 1, 4, 9, 14, 19])
def test_random_reader_combinations_vs_port(seed):
    """The vertical velocity from a reader of its own (another grid, other levels: the kernel cannot share the cell and
    weights of the current sample), the wind from a global periodic grid over a regional current, grids that are global by
    the reference's rule without being periodic (north-south coverage test only, edge value in the east-west gap)."""
    fx = _random_readers_scenario(seed)
    if fx is None:
        pytest.skip('backward run or 2-D block drawn for this seed')
    pl, pa, pz = common.run_port(fx)
    for mode in (0, 2):
        hl, ha, hz = run_hostshim(fx, fast=mode)
        assert np.array_equal(np.isfinite(pl), np.isfinite(hl))
        m = np.isfinite(pl)
        assert max(common.max_err_deg(hl[m], ha[m], pl[m], pa[m])) < 2e-8, (seed, mode)
        assert np.nanmax(np.abs(hz.astype(float) - pz.astype(float))) <= 1e-5


def _random_wdf_scenario(seed):
    rng=np.random.default_rng(70000+seed)
    fx=_random_scenario(seed+15000)
    n=fx.n
    nsl=fx.u.shape[0]; ny,nx=len(fx.grid_lat),len(fx.grid_lon)
    if fx.x_wind is None:
        fx.x_wind=(5*rng.normal(size=(nsl,ny,nx))).astype(np.float32); fx.y_wind=(5*rng.normal(size=(nsl,ny,nx))).astype(np.float32)
    fx.meta['wind']=True
    wdf=rng.uniform(0.0,0.05,n).astype(np.float32); wdf[:50]=0
    fx.wdf_array=wdf
    cdf=(rng.uniform(0.5,1.0,n).astype(np.float32)) if seed%2 else None
    fx.cdf=cdf
    fx.props=(lambda c=cdf,w=wdf,n=n: ((c if c is not None else np.float32(1)*np.ones(n)), w, np.ones(n,dtype=np.int32)))
    wdd=float(rng.choice([0.0,0.1,1.0])); fx.meta['wind_drift_depth']=wdd; fx.wind_drift_depth=(lambda w=wdd: w)
    return fx


@pytest.mark.parametrize('seed', range(4))
def test_random_float32_drift_factor_arrays_vs_port(seed):
    """Per-element float32 wind_drift_factor (and current_drift_factor) arrays: the wind move and the final move then run in
    float32 azimuth / speed (update_positions, basemodel/__init__.py:4643-4650), with the surface taper of drift:wind_drift_depth."""
    fx = _random_wdf_scenario(seed)
    pl, pa, pz = common.run_port(fx)
    for mode in (0, 2):
        hl, ha, hz = run_hostshim(fx, fast=mode)
        assert np.array_equal(np.isfinite(pl), np.isfinite(hl))
        m = np.isfinite(pl)
        assert max(common.max_err_deg(hl[m], ha[m], pl[m], pa[m])) < 5e-8, (seed, mode)
