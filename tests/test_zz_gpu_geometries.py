"""od_interp on the device against the port on random block geometries (the host build of the same source is checked on
160 geometries in tests/test_hostmath.py; this runs a handful through the C-ABI on the GPU)."""
from datetime import timedelta

import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('seed', [3, 10, 47, 69, 101])
def test_device_sampler_bit_exact_on_random_geometries(seed):
    from oracle import advect_port as ap
    from opendrift_b200.engine import Engine
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
        z = (-np.cumsum(rng.uniform(0.5, 20.0, nz)) + rng.uniform(0, 3)).astype(np.float32).astype(np.float64)
        if rng.uniform() < 0.5:
            z = z[::-1].copy()
    times = [common.syn.T0 + timedelta(hours=i) for i in range(3)]
    shape = (3, nz, ny, nx) if nz > 1 else (3, ny, nx)
    ncomp = 1 if seed % 3 == 0 else 2
    names = ['upward_sea_water_velocity'] if ncomp == 1 else list(common.CUR)
    fields = [rng.normal(size=shape).astype(np.float32) for _ in range(ncomp)]
    r = ap.GridReader(lon, lat, z, times, dict(zip(names, fields)))
    eng = Engine(0)
    grp = eng.add_group(lon, lat, z, ncomp, times, lambda ti, c: fields[c][ti], tuple([0.0] * ncomp))
    n = 4000
    plon = rng.uniform(float(lon.min()) - 2 * abs(dx), float(lon.max()) + 2 * abs(dx), n)
    plat = rng.uniform(float(lat.min()) - 0.3, float(lat.max()) + 0.3, n)
    plon[:20], plat[20:40] = float(lon[-1]), float(lat[-1])
    plon[40:60], plat[60:80] = float(lon[0]), float(lat[0])
    pz = (rng.uniform(float(z.min()) - 5, 3.0, n) if z is not None else np.zeros(n)).astype(np.float32)
    for off, pos32 in ((0, 0), (1800, 0), (4321, 0), (4321, 1)):
        t = times[0] + timedelta(seconds=off)
        lo, la = (plon.astype(np.float32), plat.astype(np.float32)) if pos32 else (plon, plat)
        env = ap.get_environment([r], names, t, lo, la, pz)
        outs = eng.interp(grp, t, eng.to_device(lo.astype(np.float64)), eng.to_device(la.astype(np.float64)), eng.to_device(pz),
                          pos_f32=bool(pos32))
        for k, nme in enumerate(names):
            assert np.array_equal(outs[k].cpu().numpy(), env[nme]), (seed, off, pos32, nme)
    eng.close()
