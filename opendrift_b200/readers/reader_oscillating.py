"""A value oscillating in time, the same everywhere: opendrift/readers/reader_oscillating.py (tidal currents, sea surface height in
the reference's examples and tests), same constructor.  A ContinuousReader: the value is exact at every time it is asked for,
Runge-Kutta stage times included, not a lerp between slabs."""
from datetime import datetime, timedelta

import numpy as np

from .continuous import ContinuousReader


class Reader(ContinuousReader):

    def __init__(self, variable, amplitude, period=timedelta(hours=24), period_seconds=None, phase=0, zero_time=datetime(2017, 1, 1, 0)):
        if period_seconds is not None:
            raise ValueError('Input parameter "period_seconds" is deprecated, please use "period" (timedelta) instead')
        self.variables = [variable]
        self.amplitude = amplitude
        self.period_seconds = period.total_seconds()
        self.zero_time = zero_time
        self.proj4 = '+proj=latlong +datum=WGS84'
        self.name = 'oscillating_reader'
        super().__init__()

    def get_variables(self, requestedVariables, time=None, x=None, y=None, z=None):
        # (reader_oscillating.py:60-66: the constructor's `phase` is accepted and not used there either)
        angle = ((time - self.zero_time).total_seconds() / self.period_seconds) * np.pi
        value = self.amplitude * np.sin(angle)
        return {'time': time, 'x': x, 'y': y, 'z': z, self.variables[0]: value * np.ones(np.shape(x))}
