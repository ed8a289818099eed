"""A reader that gives the same value everywhere, always: opendrift/readers/reader_constant.py (the reference's tests and
gallery scripts use it for wind, current, waves).  Same constructor -- a {variable: value} map or keyword arguments.

On the GPU path it is a tiny periodic global grid (4 x 2 points, one static slab) bound like any other gridded reader, so
constant forcing takes part in the fused step kernel like gridded forcing does.  The per-element form (arrays + element_ID)
is not on the GPU path; the reference itself recommends seeding the environment instead."""
import numpy as np

from .basereader import StructuredReader


# [x_component, y_component, magnitude, direction_to] (basereader/consts.py:27-36)
_VECTOR_FAMILIES = [
    ('x_wind', 'y_wind', 'wind_speed', 'wind_to_direction'),
    ('sea_ice_x_velocity', 'sea_ice_y_velocity', 'sea_ice_speed', 'direction_of_sea_ice_velocity'),
    ('x_sea_water_velocity', 'y_sea_water_velocity', 'sea_water_speed', 'sea_water_to_direction'),
    ('sea_surface_wave_stokes_drift_x_velocity', 'sea_surface_wave_stokes_drift_y_velocity',
     'sea_surface_wave_stokes_drift_speed', 'sea_surface_wave_stokes_drift_to_direction')]


class Reader(StructuredReader):
    always_valid = True

    def __init__(self, *args, **kwargs):
        if len(args) == 1 and isinstance(args[0], dict):
            parameter_value_map = dict(args[0])
        elif kwargs:
            parameter_value_map = dict(kwargs)
        else:
            raise ValueError('reader_constant.Reader needs a {variable: value} map or keyword arguments')
        if 'element_ID' in parameter_value_map or any(np.size(v) != 1 for v in parameter_value_map.values()):
            raise NotImplementedError('element-dependent constant readers (arrays + element_ID) are not on the GPU path; '
                                      'seed the values as element properties / environment instead')
        self._parameter_value_map = {k: np.atleast_1d(v) for k, v in parameter_value_map.items()}
        # x / y components derived from (magnitude, direction_to) when only those are given: the automatic environment
        # mapping of the reference's readers (basereader/variables.py:536-545, vector_from_speed_and_direction :468-472)
        for xname, yname, speed, direction in _VECTOR_FAMILIES:
            m = self._parameter_value_map
            if speed in m and direction in m and xname not in m and yname not in m:
                ang = np.radians(m[direction].astype(np.float64))
                m[xname] = m[speed] * np.cos(ang)
                m[yname] = m[speed] * np.sin(ang)
        self.variables = list(self._parameter_value_map)
        self.proj4 = '+proj=latlong'
        self.xmin, self.xmax, self.ymin, self.ymax = -180, 180, -90, 90
        self.delta_x, self.delta_y = 90.0, 180.0
        self.start_time = self.end_time = self.time_step = None
        self.name = 'constant_reader'
        self._lon = np.array([-180.0, -90.0, 0.0, 90.0], dtype=np.float32)      # periodic: 4 * 90 deg
        self._lat = np.array([-90.0, 90.0], dtype=np.float32)
        super().__init__()

    def get_variables(self, requested_variables, time=None, x=None, y=None, z=None):
        out = {'x': self._lon, 'y': self._lat, 'z': 0, 'time': time}
        for v in requested_variables:
            out[v] = np.full((2, 4), float(self._parameter_value_map[v][0]), dtype=np.float32)
        return out
