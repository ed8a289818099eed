"""Analytical double-gyre current on a projected plane: the reader of BASELINE configs[0]
(examples/example_double_gyre_advection_schemes.py), same constructor, attributes and helper methods as
opendrift/readers/reader_double_gyre.py (a ContinuousReader, opendrift/readers/basereader/continuous.py).

The field, the projection and the vector rotation are evaluated per particle on the device (csrc/od_analytic.cuh);
this class carries the parameters (od_analytic_desc) and the reference's reader surface:
`get_variables_interpolated`, `xy2lonlat` / `lonlat2xy`, coverage attributes."""
from datetime import datetime

import numpy as np

from .. import _lib
from ..errors import OutsideSpatialCoverageError
from .projection import SphericalStereographic


class Reader:
    always_valid = False

    def __init__(self, initial_time=datetime(2000, 1, 1, 0, 0), epsilon=0.1, omega=0.628, A=0.25,
                 proj4='+proj=stere +lat_0=0 +lon_0=0 +lat_ts=0 +units=m +a=6.371e+06 +e=0 +no_defs'):
        self.fileName = 'double_gyre'
        self.name = 'double_gyre'
        self.proj4 = proj4
        self.proj = SphericalStereographic(proj4)
        self.xmin, self.xmax, self.ymin, self.ymax = 0., 2., 0., 1.
        self.zmin, self.zmax = -np.inf, np.inf
        self.A, self.epsilon, self.omega = A, epsilon, omega
        self.initial_time = initial_time
        self.start_time = self.end_time = self.time_step = None
        self.variables = ['x_sea_water_velocity', 'y_sea_water_velocity']
        self._engine = None
        self._fallback = {}
        # modulate_longitude (basereader/variables.py:259-280): the longitude convention follows the corner longitudes
        exlons, _ = self.xy2lonlat(np.array([self.xmin, self.xmin, self.xmax, self.xmax]),
                                   np.array([self.ymin, self.ymax, self.ymax, self.ymin]))
        self._lon_0to360 = not (np.min(exlons) < 0)

    # -- the reference's helper surface --------------------------------------------------------------------------
    def xy2lonlat(self, x, y):
        return self.proj(x, y, inverse=True)

    def lonlat2xy(self, lon, lat):
        return self.proj(lon, lat, inverse=False)

    def modulate_longitude(self, lons):
        lons = np.asarray(lons)
        return np.mod(lons, 360) if self._lon_0to360 else np.mod(lons + 180, 360) - 180

    def covers_time(self, time):
        return True

    def global_coverage(self):
        return False

    def covers_positions(self, lon, lat, z=0):
        x, y = self.lonlat2xy(self.modulate_longitude(np.atleast_1d(lon)), np.atleast_1d(lat))
        ind = np.where((x >= self.xmin) & (x <= self.xmax) & (y >= self.ymin) & (y <= self.ymax))[0]
        return ind, x[ind], y[ind]

    # -- device binding ------------------------------------------------------------------------------------------
    def bind(self, engine, fallback=None):
        self._engine = engine
        self._fallback = dict(fallback or {})

    def unbind(self):
        self._engine = None

    def analytic_desc(self, with_fallback=True):
        """od_analytic_desc of this reader (include/odcuda.h)."""
        d = _lib.AnalyticDesc()
        d.kind = _lib.OD_ANALYTIC_DOUBLE_GYRE
        d.lon_mode = _lib.OD_LON_0_360 if self._lon_0to360 else _lib.OD_LON_PM180
        d.proj = self.proj.desc()
        d.xmin, d.xmax, d.ymin, d.ymax = self.xmin, self.xmax, self.ymin, self.ymax
        d.par[0], d.par[1], d.par[2], d.par[3] = float(self.A), float(self.epsilon), float(self.omega), 0.0
        d.rot_delta = 10.0                       # rotate_vectors: 10 m along the y axis of a projected plane
        for c, v in enumerate(self.variables):
            fb = self._fallback.get(v) if with_fallback else None
            d.fallback[c] = float('nan') if fb is None else float(fb)
        return d

    def seconds(self, time):
        """reader_double_gyre.py:66: t = (time - initial_time).total_seconds()"""
        return (time - self.initial_time).total_seconds()

    def device_sample(self, engine, time, d_lon, d_lat, pos_f32=False):
        """{variable: float32 device tensor}, NaN where the reader does not cover the position (no fallback)."""
        u, v = engine.analytic_interp(self.analytic_desc(with_fallback=False), self.seconds(time), d_lon, d_lat, pos_f32)
        return {'x_sea_water_velocity': u, 'y_sea_water_velocity': v}

    # -- the reference's public entry point (basereader/variables.py:860-920) -----------------------------------
    def get_variables_interpolated(self, variables, profiles=None, profiles_depth=None, time=None,
                                   lon=None, lat=None, z=None, rotate_to_proj=None):
        """East / north velocity at lon, lat (always rotated to geographic axes, which is what the model asks for with
        rotate_to_proj='+proj=latlong'); masked where uncovered.  Values are float32, as Environment stores them."""
        if isinstance(variables, str):
            variables = [variables]
        assert set(variables).issubset(self.variables), f'{variables} is not subset of {self.variables}'
        if self._engine is None:
            from ..engine import default_engine
            self.bind(default_engine())
        eng = self._engine
        lon_in, lat_in = np.atleast_1d(lon), np.atleast_1d(lat)
        pos_f32 = lon_in.dtype == np.float32 and lat_in.dtype == np.float32
        out = self.device_sample(eng, time, eng.to_device(lon_in.astype(np.float64)), eng.to_device(lat_in.astype(np.float64)), pos_f32)
        env = {v: np.ma.masked_invalid(out[v].cpu().numpy()) for v in variables}
        if all(np.ma.getmaskarray(a).all() for a in env.values()):
            raise OutsideSpatialCoverageError('All %s particles are outside domain of %s' % (len(lon_in), self.name))
        env_profiles = None
        if profiles is not None:                 # continuous.py:40-46: the value itself at both ends of the profile
            env_profiles = {'z': [0, -profiles_depth]}
            for var in profiles:
                env_profiles[var] = np.ma.array([env[var], env[var]])
        return env, env_profiles

    def __repr__(self):
        return 'Reader: %s  x [%s..%s] m, y [%s..%s] m on %s' % (self.name, self.xmin, self.xmax, self.ymin, self.ymax, self.proj4)
