"""ORACLE (test infrastructure, never the product path): NumPy restatement of the reference's analytical double-gyre
reader on its stereographic plane -- BASELINE.json configs[0], examples/example_double_gyre_advection_schemes.py.

What the reference does for every get_environment call with this reader, restated step by step:
  Variables.get_variables_interpolated (readers/basereader/variables.py:860-920): modulate_longitude (:259-280) ->
  lonlat2xy (:129-143, pyproj.Proj forward) -> get_variables_interpolated_xy (:709-858): covers_positions_xy (:229-257)
  -> ContinuousReader._get_variables_interpolated_ (basereader/continuous.py:32-48) -> Reader.get_variables
  (readers/reader_double_gyre.py:57-82: the analytical field, float64) -> rotate_vectors (variables.py:59-109: azimuth
  of the reader's y axis from a 10 m line through Transformer + Geod.inv, then a rotation by minus that angle) ->
  NaN for uncovered particles (:841-853).
The projection and the geodesic are third-party (pyproj): restated in oracle/proj_stere.py and oracle/geod_karney.py.

Pinned against the real reference (oracle/refrun.py with those two restatements injected as the fake pyproj) by
tests/test_oracle_gyre.py on the fixtures tests/golden/ref_gyre_*.npz (oracle/make_golden.py).
"""
import numpy as np

from . import geod_karney
from .proj_stere import Stere

DEFAULT_PROJ4 = '+proj=stere +lat_0=0 +lon_0=0 +lat_ts=0 +units=m +a=6.371e+06 +e=0 +no_defs'


class DoubleGyreReader:
    """reader_double_gyre.Reader for the port's get_environment (oracle/advect_port.py)."""

    variables = ['x_sea_water_velocity', 'y_sea_water_velocity']

    def __init__(self, initial_time, epsilon=0.1, omega=0.628, A=0.25, proj4=DEFAULT_PROJ4):
        self.initial_time = initial_time
        self.epsilon, self.omega, self.A = epsilon, omega, A
        self.proj = Stere(proj4)
        self.xmin, self.xmax, self.ymin, self.ymax = 0., 2., 0., 1.
        self.start_time = self.end_time = self.time_step = None
        self.global_coverage = False
        # modulate_longitude (variables.py:259-280): the longitude range follows the corner longitudes
        exlons, _ = self.proj.inverse(np.array([self.xmin, self.xmin, self.xmax, self.xmax]),
                                      np.array([self.ymin, self.ymax, self.ymax, self.ymin]))
        self.lon_0to360 = not (np.min(exlons) < 0)
        self._geod = geod_karney.Geod()

    def field(self, time, x, y):
        """reader_double_gyre.py:66-73"""
        t = (time - self.initial_time).total_seconds()
        a = self.epsilon * np.sin(self.omega * t)
        b = 1 - 2 * self.epsilon * np.sin(self.omega * t)
        f = a * x * x + b * x
        dfdx = 2 * a * x + b
        u = -np.pi * self.A * np.sin(np.pi * f) * np.cos(np.pi * y)
        v = np.pi * self.A * np.cos(np.pi * f) * np.sin(np.pi * y) * dfdx
        return u, v

    def rotation(self, x, y):
        """Azimuth (radians) of the reader's y axis at (x, y): variables.py:84-98."""
        lon1, lat1 = self.proj.inverse(x, y)
        lon2, lat2 = self.proj.inverse(x, y + 10)
        return np.radians(self._geod.inv(lon1, lat1, lon2, lat2)[0])

    def interpolate(self, variables, time, lon, lat, z):
        lon = np.mod(lon, 360) if self.lon_0to360 else np.mod(lon + 180, 360) - 180     # in the dtype of lon
        x, y = self.proj.forward(lon, lat)
        n = len(x)
        covered = np.where((x >= self.xmin) & (x <= self.xmax) & (y >= self.ymin) & (y <= self.ymax))[0]
        out = {v: np.full(n, np.nan) for v in variables}
        if len(covered) == 0:
            return out
        xc, yc = x[covered], y[covered]
        u, v = self.field(time, xc, yc)
        rot = -self.rotation(xc, yc)
        u_rot = u * np.cos(rot) - v * np.sin(rot)
        v_rot = u * np.sin(rot) + v * np.cos(rot)
        res = {'x_sea_water_velocity': u_rot, 'y_sea_water_velocity': v_rot}
        for name in variables:
            out[name][covered] = res[name]
        return out
