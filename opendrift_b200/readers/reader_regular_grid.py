"""In-memory regular lon/lat(/z) grid reader: the role the reference's reader_netCDF_CF_generic (block
supplier, opendrift/readers/reader_netCDF_CF_generic.py:404-626) and reader_constant_2d play for gridded
forcing -- `get_variables()` hands out full-grid blocks, float32 coordinates (:586-587)."""
import numpy as np

from .basereader import StructuredReader


class Reader(StructuredReader):
    def __init__(self, lon, lat, z=None, times=None, fields=None, name='regular_grid', subblocks=False, proj4='+proj=latlong'):
        """fields: dict variable -> array (nt, [nz,] ny, nx) float32 (NumPy or CUDA tensors), or a callable
        fields[var](time_index) -> ([nz,] ny, nx).  subblocks: hand out only the part of the grid around the requested
        positions plus `buffer` cells, like a file reader does (reader_netCDF_CF_generic.py:436-466)."""
        self.subblocks = bool(subblocks)
        self.proj4 = proj4              # a projected plane (spherical +proj=stere): lon / lat are then its x / y axes in metres
        self.lon = np.asarray(lon, dtype=np.float32)
        self.lat = np.asarray(lat, dtype=np.float32)
        self.zlev = None if z is None else np.asarray(z, dtype=np.float64)
        self.xmin, self.xmax = float(self.lon.min()), float(self.lon.max())
        self.ymin, self.ymax = float(self.lat.min()), float(self.lat.max())
        self.delta_x = float(self.lon[1] - self.lon[0])
        self.delta_y = float(self.lat[1] - self.lat[0])
        self.numx, self.numy = len(self.lon), len(self.lat)
        self.fields = fields
        self.variables = list(fields.keys())
        self.times = list(times) if times is not None else [None]
        self.start_time, self.end_time = self.times[0], self.times[-1]
        self.time_step = (self.times[1] - self.times[0]) if len(self.times) > 1 else None
        self.name = name
        super().__init__()

    def get_variables(self, requested_variables, time=None, x=None, y=None, z=None):
        ti = self.times.index(time) if time is not None else 0
        sx = sy = slice(None)
        if self.subblocks and x is not None and y is not None:
            # the cells the positions fall in, widened by the buffer on every side
            ix = np.floor(np.abs(np.asarray(x, dtype=np.float64) - float(self.lon[0])) / self.delta_x).astype(int)
            iy = np.floor(np.abs(np.asarray(y, dtype=np.float64) - float(self.lat[0])) / self.delta_y).astype(int)
            b = int(self.buffer)
            sx = slice(max(0, int(ix.min()) - b), min(int(ix.max()) + b + 1, self.numx))
            sy = slice(max(0, int(iy.min()) - b), min(int(iy.max()) + b + 1, self.numy))
            if sx.stop - sx.start < 2:
                sx = slice(max(0, sx.start - 1), min(sx.start + 2, self.numx))
            if sy.stop - sy.start < 2:
                sy = slice(max(0, sy.start - 1), min(sy.start + 2, self.numy))
            self.blocks_served = getattr(self, 'blocks_served', 0) + 1
        out = {'x': self.lon[sx], 'y': self.lat[sy], 'time': time}
        three_d = False
        for v in requested_variables:
            f = self.fields[v]
            a = f(ti) if callable(f) else f[ti]
            three_d |= getattr(a, 'ndim', 2) == 3
            if sx != slice(None) or sy != slice(None):
                a = a[..., sy, sx]
                a = a.contiguous() if hasattr(a, 'is_cuda') else np.ascontiguousarray(a)
            out[v] = a
        out['z'] = self.zlev if three_d else 0
        return out
