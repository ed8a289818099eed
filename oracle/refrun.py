"""ORACLE (test infrastructure): import and run the UNMODIFIED reference Python from
/root/reference in THIS container.

The reference's own arithmetic for the hot path (OpenDriftSimulation.run ->
Environment.get_environment -> StructuredReader/ReaderBlock -> Linear2DInterpolator
-> PhysicsMethods.advect_ocean_current -> update_positions) is pure NumPy/SciPy and
imports fine once the plotting / IO third-party packages that are absent from this
image are replaced by MagicMock modules.  The single piece of missing *arithmetic*
is ``pyproj.Geod.fwd``; a fake ``pyproj`` module backed by oracle/geod_karney.py
(validated against mpmath, see oracle/geod_exact.py) is injected for it.

This module cannot travel to the GPU box (no /root/reference there): it is used only
by oracle/make_golden.py to write tests/golden/*.npz, and by the CPU tests that are
skipped when /root/reference is absent.
"""
import os
import sys
import types
import logging
from unittest.mock import MagicMock
from datetime import datetime, timedelta

import numpy as np

REFERENCE_ROOT = os.environ.get('OPENDRIFT_REFERENCE', '/root/reference')
if not os.path.isdir(os.path.join(REFERENCE_ROOT, 'opendrift')):
    # the GPU box: the byte-for-byte copy that oracle/build_ref.py made in the build container (oracle/_ref, git-ignored)
    REFERENCE_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')

_MOCKED = [
    'xarray', 'netCDF4', 'matplotlib', 'matplotlib.pyplot', 'matplotlib.animation',
    'matplotlib.patches', 'matplotlib.path', 'matplotlib.widgets', 'matplotlib.colors',
    'matplotlib.tri', 'matplotlib.cm', 'matplotlib.dates', 'matplotlib.ticker',
    'mpl_toolkits', 'mpl_toolkits.axes_grid1',
    'cartopy', 'cartopy.crs', 'cartopy.feature', 'cartopy.io', 'cartopy.io.shapereader',
    'cartopy.mpl', 'cartopy.mpl.gridliner', 'cartopy.mpl.ticker',
    'geojson', 'roaring_landmask', 'cmocean', 'coloredlogs', 'copernicusmarine', 'dotenv',
    'geopandas', 'shapely', 'shapely.geometry', 'shapely.ops', 'shapely.prepared',
    'shapely.vectorized', 'shapely.strtree', 'utm', 'pynucos', 'nc_time_axis', 'pykdtree',
    'pykdtree.kdtree', 'xhistogram', 'xhistogram.xarray', 'adios_db', 'cftime', 'trajan',
    'dask', 'dask.array', 'h5py', 'pygrib', 'cfgrib', 'earthaccess', 'pyresample',
]


def _fake_pyproj():
    from oracle.geod_karney import Geod

    class _CRS:
        def __init__(self, geographic=True, srs='+proj=latlong'):
            self.is_geographic = geographic
            self.srs = srs

        def __eq__(self, other):
            return isinstance(other, _CRS) and other.srs == self.srs

        def __hash__(self):
            return hash(self.srs)

    class Proj:
        """'+proj=latlong' style CRSs (identity) and the spherical stereographic projection of
        reader_double_gyre.py (oracle/proj_stere.py)."""

        def __init__(self, projparams=None, **kw):
            s = str(projparams)
            self.srs = s
            self.definition_string = lambda: s
            if any(k in s for k in ('latlong', 'longlat', 'lonlat', 'latlon')):
                self.impl = None
                self.crs = _CRS(True, s)
            elif '+proj=stere' in s:
                from oracle.proj_stere import Stere, parse_proj4
                pp = parse_proj4(s)
                sphere = 'R' in pp or not any(k in pp for k in ('ellps', 'datum', 'rf', 'f', 'b', 'es')) and float(pp.get('e', 0.0)) == 0.0
                if sphere:
                    self.impl = Stere(s)
                else:
                    from oracle.proj_conformal import StereEllipsoid
                    self.impl = StereEllipsoid(s)
                self.crs = _CRS(False, s)
            elif '+proj=merc' in s or '+proj=lcc' in s:
                from oracle.proj_conformal import make
                self.impl = make(s)
                self.crs = _CRS(False, s)
            else:
                raise NotImplementedError('fake pyproj supports latlong, spherical stere, merc and lcc only: ' + s)

        def __call__(self, x, y, inverse=False):
            if self.impl is None:
                return x, y
            scalar = np.isscalar(x)
            a, b = (self.impl.inverse if inverse else self.impl.forward)(x, y)
            if scalar:
                return float(a), float(b)
            return a, b

    class Transformer:
        """Transformer.from_proj(a, b).transform(x, y): through geographic coordinates, no datum shift (PROJ applies
        none between a sphere-based projection without datum and a latlong CRS: 'ballpark' transformation)."""

        def __init__(self, pf, pt):
            self.pf, self.pt = pf, pt

        @classmethod
        def from_proj(cls, proj_from, proj_to, **kw):
            pf = proj_from if isinstance(proj_from, Proj) else Proj(proj_from)
            pt = proj_to if isinstance(proj_to, Proj) else Proj(proj_to)
            return cls(pf, pt)

        def transform(self, x, y):
            lon, lat = self.pf(x, y, inverse=True)
            return self.pt(lon, lat)

    m = types.ModuleType('pyproj')
    m.Proj = Proj
    m.Geod = Geod
    m.CRS = MagicMock()
    m.Transformer = Transformer
    m.__version__ = '0.0-fake'
    return m


def _fake_xarray():
    """xarray stand-in: everything is a MagicMock except Dataset, which holds the variables of the reference's result block as
    plain NumPy arrays -- enough for what run() does with it outside state_to_buffer: the float32 copies of the first output
    column that serve as `_elements_previous` / `_environment_previous` (basemodel/__init__.py:2164-2165; read and written by
    release_elements :928-931, update_previous_state :642-669 and interact_with_coastline :671-746)."""
    class Var(np.ndarray):
        def assign_attrs(self, *a, **k):
            return self

    class Dataset:
        def __init__(self, coords=None, data_vars=None, attrs=None):
            self._vars = {}
            for k, v in (data_vars or {}).items():
                self._vars[k] = (np.array(v[1]) if isinstance(v, tuple) else np.array(v)).view(Var)
            self.attrs = dict(attrs or {})
            self.coords = coords or {}

        @property
        def data_vars(self):
            return self._vars

        @property
        def sizes(self):
            a = next(iter(self._vars.values()), np.zeros((0, 0)))
            return {'trajectory': a.shape[0], 'time': a.shape[1] if a.ndim > 1 else 1}

        def _sub(self, d):
            o = Dataset(attrs=self.attrs)
            o._vars = d
            o.coords = self.coords
            return o

        def __getitem__(self, k):
            if isinstance(k, (list, tuple)):
                return self._sub({n: self._vars[n] for n in k})
            return self._vars[k]

        def __setitem__(self, k, v):
            self._vars[k] = v

        def __getattr__(self, k):
            v = self.__dict__.get('_vars', {})
            if k in v:
                return v[k]
            c = self.__dict__.get('coords', {})
            if k in c:
                return c[k][1] if isinstance(c[k], tuple) else c[k]
            raise AttributeError(k)

        def __iter__(self):
            return iter(self._vars)

        def __contains__(self, k):
            return k in self._vars

        def isel(self, time=None, trajectory=None, drop=False):
            d = {}
            for n, a in self._vars.items():
                if time is not None and a.ndim > 1:
                    a = a[:, time]
                if trajectory is not None:
                    a = a[np.asarray(trajectory)]
                d[n] = a
            return self._sub(d)

        def copy(self, deep=True):
            return self._sub({n: (np.array(a, copy=True) if deep else a) for n, a in self._vars.items()})

    m = MagicMock()
    m.Dataset = Dataset
    return m


_ready = False


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'opendrift'))


def setup():
    """Install the stub modules and put the reference on sys.path (idempotent)."""
    global _ready
    if _ready:
        return
    if not available():
        raise RuntimeError('reference tree not found at %s' % REFERENCE_ROOT)
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if here not in sys.path:
        sys.path.insert(0, here)
    for name in _MOCKED:
        if name not in sys.modules:
            sys.modules[name] = MagicMock()
    sys.modules['pyproj'] = _fake_pyproj()
    sys.modules['xarray'] = _fake_xarray()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import opendrift.readers.basereader  # noqa: F401  (must precede interpolation.structured)
    from opendrift.models.basemodel import OpenDriftSimulation
    # xarray is mocked: the result buffer cannot be written. It is not on the hot path.
    OpenDriftSimulation.state_to_buffer = lambda self, final=False: None
    logging.getLogger('opendrift').setLevel(logging.CRITICAL)
    _ready = True


def make_grid_reader(lon, lat, z, times, fields, name='synthetic_grid', subblocks=False, proj4='+proj=latlong'):
    """A reference StructuredReader (subclass of the reference base class) serving
    in-memory regular lon/lat(/z) slabs.  ``fields[var]`` has shape (nt, nz, ny, nx) or
    (nt, ny, nx) float32.  get_variables returns the FULL grid block (like
    reader_constant_2d.py:45-49) with float32 x/y (as reader_netCDF_CF_generic.py:586-587).

    A grid that is global east-west by the reference's rule (variables.py:289-301) and exactly periodic
    (nx * dx == 360) returns the full circle plus ONE wrapped column (x = lon[-1] + dx, values of column 0), so that the
    block is monotonic, stays inside [-180, 360] as ReaderBlock demands (interpolation/structured.py:35-48) and covers
    the seam cell between the last and the first column -- what reader_netCDF_CF_generic assembles from two parts when
    the particle cloud straddles its longitude border (reader_netCDF_CF_generic.py:452-463).
    """
    setup()
    from opendrift.readers.basereader.structured import StructuredReader

    class Reader(StructuredReader):
        def __init__(self):
            self.proj4 = proj4              # a projected plane: lon / lat are then the x / y axes in metres
            self.lon = np.asarray(lon, dtype=np.float32)
            self.lat = np.asarray(lat, dtype=np.float32)
            self.zlev = None if z is None else np.asarray(z, dtype=np.float64)
            self.xmin, self.xmax = float(self.lon.min()), float(self.lon.max())
            self.ymin, self.ymax = float(self.lat.min()), float(self.lat.max())
            self.delta_x = float(self.lon[1] - self.lon[0])
            self.delta_y = float(self.lat[1] - self.lat[0])
            self.numx, self.numy = len(self.lon), len(self.lat)
            self.variables = list(fields.keys())
            self.fields = fields
            self.times = list(times)
            self.start_time, self.end_time = times[0], times[-1]
            self.time_step = (times[1] - times[0]) if len(times) > 1 else None
            self.name = name
            super().__init__()
            self.periodic = bool(self.global_coverage()) and abs(self.numx * self.delta_x - 360.0) < 1e-3 * self.delta_x
            self.block_x = self.lon
            if self.periodic:
                self.block_x = np.append(self.lon, np.float32(self.lon[-1] + np.float32(self.delta_x))).astype(np.float32)

        def get_variables(self, requested_variables, time=None, x=None, y=None, z=None):
            it = self.times.index(time)
            sx = sy = slice(None)
            if subblocks and x is not None and y is not None and not self.periodic:
                # a sub-block around the requested positions, widened by self.buffer cells (the role of
                # reader_netCDF_CF_generic.py:436-466; self.buffer is set by the reference's set_buffer_size)
                ix = np.floor(np.abs(np.asarray(x, dtype=np.float64) - float(self.lon[0])) / self.delta_x).astype(int)
                iy = np.floor(np.abs(np.asarray(y, dtype=np.float64) - float(self.lat[0])) / self.delta_y).astype(int)
                b = int(self.buffer)
                sx = slice(max(0, int(ix.min()) - b), min(int(ix.max()) + b + 1, self.numx))
                sy = slice(max(0, int(iy.min()) - b), min(int(iy.max()) + b + 1, self.numy))
                self.blocks_served = getattr(self, 'blocks_served', 0) + 1
                self.block_shapes = getattr(self, 'block_shapes', []) + [(sy.stop - sy.start, sx.stop - sx.start)]
            out = {'x': self.block_x[sx], 'y': self.lat[sy], 'time': time}
            three_d = False
            for v in requested_variables:
                a = self.fields[v][it]
                three_d |= a.ndim == 3
                a = np.array(a[..., sy, sx], dtype=np.float32, copy=True)
                if self.periodic:
                    a = np.concatenate([a, a[..., :1]], axis=-1)
                out[v] = a
            out['z'] = self.zlev if three_d else 0
            return out

    return Reader()


def run_oceandrift(readers, lon, lat, z, start_time, time_step, steps, config=None,
                   seed_kwargs=None, model='OceanDrift', seed=0):
    """Run the reference model and return final (lon, lat, z, status-free) arrays."""
    setup()
    import tempfile
    if model == 'OceanDrift':
        from opendrift.models.oceandrift import OceanDrift as Model
    elif model == 'Leeway':
        from opendrift.models.leeway import Leeway as Model
    else:
        raise ValueError(model)
    logfile = os.path.join(tempfile.gettempdir(), 'oracle_refrun.log')
    o = Model(loglevel=50, logfile=logfile, seed=seed)
    for r in readers:
        o.add_reader(r)
    cfg = {'general:use_auto_landmask': False,
           'environment:constant:land_binary_mask': 0,
           'general:coastline_action': 'none'}
    cfg.update(config or {})
    for k, v in cfg.items():
        o.set_config(k, v)
    o.seed_elements(lon=lon, lat=lat, z=z, time=start_time, **(seed_kwargs or {}))
    o.run(steps=steps, time_step=time_step, time_step_output=time_step)
    return o
