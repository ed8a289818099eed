"""Static 2-D arrays on a regular grid, valid at all times: opendrift/readers/reader_constant_2d.py, same constructor
(`Reader(x, y, {'variable': array[ny, nx], ...}, proj4='+proj=latlong')`).  A gridded reader with one slab that never changes:
bound once, sampled by the same kernels as any other field group (bilinear; land_binary_mask from the nearest grid point)."""
import numpy as np

from . import reader_regular_grid


class Reader(reader_regular_grid.Reader):

    def __init__(self, x, y, array_dict, proj4='+proj=latlong'):
        x, y = np.asarray(x), np.asarray(y)
        fields = {}
        for name, a in array_dict.items():
            a = np.ma.filled(np.ma.masked_invalid(np.asarray(a, dtype=np.float64)), np.nan).astype(np.float32)
            if a.shape != (len(y), len(x)):
                raise ValueError('%s: expected an array of shape (len(y), len(x)) = %s, got %s' % (name, (len(y), len(x)), a.shape))
            fields[name] = a[None]
        super().__init__(x, y, None, None, fields, name='reader_constant_2d', proj4=proj4)
        self.array_dict = dict(array_dict, x=x, y=y, z=0)
