"""Synthetic forcing fields and particle clouds of BASELINE.json's configs
(SURVEY.md section 8(d)): a depth-attenuated, time-dependent double gyre on a regular
lon/lat/z grid.  Pure NumPy input generation shared by the bench, the tests and the
golden-vector generator; nothing here is on the measured path.

The analytic form follows the reference's analytical reader
(opendrift/readers/reader_double_gyre.py:66-73): a = eps sin(wt), b = 1 - 2 eps sin(wt),
f = a X^2 + b X, u = -pi A sin(pi f) cos(pi Y), v = pi A cos(pi f) sin(pi Y) df/dX,
here mapped onto [lon0, lon0+Lx] x [lat0, lat0+Ly] with X in [0, 2], Y in [0, 1] and
multiplied by g(z) = exp(z / 50).
"""
from datetime import datetime, timedelta

import numpy as np

T0 = datetime(2026, 1, 1, 0, 0, 0)


class GridSpec:
    def __init__(self, nx=512, ny=512, nz=50, lon0=0.0, dlon=0.02, lat0=55.0, dlat=0.01, dz=2.0):
        self.nx, self.ny, self.nz = nx, ny, nz
        # float32 coordinates, as netCDF readers hand them to ReaderBlock
        # (reference: opendrift/readers/reader_netCDF_CF_generic.py:586-587)
        self.lon = (lon0 + dlon * np.arange(nx)).astype(np.float32)
        self.lat = (lat0 + dlat * np.arange(ny)).astype(np.float32)
        self.z = -dz * np.arange(nz, dtype=np.float64) if nz > 1 else None
        self.Lx = float(self.lon[-1]) - float(self.lon[0])
        self.Ly = float(self.lat[-1]) - float(self.lat[0])


def double_gyre_uv(grid, t_seconds, A=0.25, eps=0.25, omega=2 * np.pi / 36000.0, three_d=True):
    """(u, v) float32 slabs of shape (nz, ny, nx) (or (ny, nx)) at time t."""
    X = 2.0 * (grid.lon.astype(np.float64) - float(grid.lon[0])) / grid.Lx
    Y = (grid.lat.astype(np.float64) - float(grid.lat[0])) / grid.Ly
    a = eps * np.sin(omega * t_seconds)
    b = 1.0 - 2.0 * eps * np.sin(omega * t_seconds)
    f = a * X * X + b * X
    dfdx = 2.0 * a * X + b
    u2 = (-np.pi * A * np.sin(np.pi * f))[None, :] * np.cos(np.pi * Y)[:, None]
    v2 = (np.pi * A * np.cos(np.pi * f) * dfdx)[None, :] * np.sin(np.pi * Y)[:, None]
    if not three_d or grid.z is None:
        return u2.astype(np.float32), v2.astype(np.float32)
    g = np.exp(grid.z / 50.0)[:, None, None]
    return (g * u2[None]).astype(np.float32), (g * v2[None]).astype(np.float32)


def upward_w(grid):
    """Static w slab (nz, ny, nx) float32: 1e-3 sin(pi X / 2) sin(pi Y) sin(pi z / zmin)."""
    X = 2.0 * (grid.lon.astype(np.float64) - float(grid.lon[0])) / grid.Lx
    Y = (grid.lat.astype(np.float64) - float(grid.lat[0])) / grid.Ly
    zz = np.sin(np.pi * grid.z / grid.z.min())
    w = 1e-3 * zz[:, None, None] * (np.sin(np.pi * Y)[:, None] * np.sin(np.pi * X / 2)[None, :])[None]
    return w.astype(np.float32)


def vertical_diffusivity(grid, t_seconds=0.0):
    """K(z) slab (nz, ny, nx) float32: 0.01 exp(z / 20) with a horizontal and a slow temporal modulation (cfg 4)."""
    X = 2.0 * (grid.lon.astype(np.float64) - float(grid.lon[0])) / grid.Lx
    Y = (grid.lat.astype(np.float64) - float(grid.lat[0])) / grid.Ly
    mod = 1.0 + 0.5 * np.sin(np.pi * X / 2)[None, :] * np.sin(np.pi * Y)[:, None]
    amp = 0.01 * (1.0 + 0.2 * np.sin(2 * np.pi * t_seconds / 43200.0))
    return (amp * np.exp(grid.z / 20.0)[:, None, None] * mod[None]).astype(np.float32)


def wind_xy(grid, t_seconds, speed=10.0):
    """2-D wind slabs (ny, nx) float32: a slowly rotating, spatially modulated 10 m/s wind."""
    X = 2.0 * (grid.lon.astype(np.float64) - float(grid.lon[0])) / grid.Lx
    Y = (grid.lat.astype(np.float64) - float(grid.lat[0])) / grid.Ly
    th = 2 * np.pi * t_seconds / 86400.0
    mod = 1.0 + 0.3 * np.sin(np.pi * X)[None, :] * np.cos(np.pi * Y)[:, None]
    return (speed * np.cos(th) * mod).astype(np.float32), (speed * np.sin(th) * mod).astype(np.float32)


def stokes_xy(grid, t_seconds, speed=0.12):
    """Surface Stokes drift slabs (ny, nx) float32, roughly aligned with the synthetic wind."""
    wx, wy = wind_xy(grid, t_seconds, speed=1.0)
    X = 2.0 * (grid.lon.astype(np.float64) - float(grid.lon[0])) / grid.Lx
    mod = (1.0 + 0.4 * np.cos(np.pi * X))[None, :]
    return (speed * wx * mod).astype(np.float32), (speed * wy * mod).astype(np.float32)


def wave_height(grid, t_seconds):
    """Significant wave height slab (ny, nx) float32, 1-3 m."""
    Y = (grid.lat.astype(np.float64) - float(grid.lat[0])) / grid.Ly
    X = 2.0 * (grid.lon.astype(np.float64) - float(grid.lon[0])) / grid.Lx
    return (2.0 + np.sin(np.pi * Y)[:, None] * np.cos(np.pi * X / 2)[None, :]).astype(np.float32)


def slab_times(n_slabs, step_seconds=3600):
    return [T0 + timedelta(seconds=step_seconds * i) for i in range(n_slabs)]


def n_slabs_for(steps, dt_seconds, step_seconds=3600):
    return int(np.ceil(steps * abs(dt_seconds) / step_seconds)) + 2


def particle_cloud(n, seed=0, three_d=True):
    """Seed positions of cfg 2: lon~U(1,9.2), lat~U(55.5,59.6), z~U(-90,0); float32-rounded as
    the reference's element constructor does (opendrift/elements/elements.py:156-158)."""
    rng = np.random.default_rng(seed)
    lon = rng.uniform(1.0, 9.2, n).astype(np.float32)
    lat = rng.uniform(55.5, 59.6, n).astype(np.float32)
    z = rng.uniform(-90.0, 0.0, n).astype(np.float32) if three_d else np.zeros(n, dtype=np.float32)
    return lon, lat, z
