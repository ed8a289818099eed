"""TEST INFRASTRUCTURE: an object with the part of opendrift_b200.engine.Engine's surface that the model classes use
when no gridded reader is bound (analytical / constant readers), backed by the host build of the device math
(tests/hostshim).  It lets the CPU suite run the drop-in model classes end to end -- seeding, release, the run loop,
Environment, the reader glue, the ctypes argument structs of Engine itself -- without a GPU.  The methods are Engine's
own (unbound functions re-used), only the library object they call is swapped for an adapter that forwards
od_* calls to the hs_* functions.  Never imported by the product."""
import ctypes as C

import numpy as np

import common
from opendrift_b200 import _lib
from opendrift_b200.engine import Engine

_P = C.c_void_p


class _HostLib:
    """od_* entry points used by the borrowed Engine methods -> hostshim"""

    def __init__(self, shim):
        self.shim = shim
        self.calls = []
        shim.hs_analytic_interp.restype = C.c_int
        shim.hs_analytic_interp.argtypes = [C.POINTER(_lib.AnalyticDesc), C.c_double, C.c_int64, _P, _P, C.c_int, _P, _P]
        shim.hs_analytic_advect.restype = C.c_int
        shim.hs_analytic_advect.argtypes = [C.POINTER(_lib.AnalyticDesc), C.POINTER(_lib.AnalyticAdvectArgs)]
        shim.hs_update_positions.restype = None
        shim.hs_update_positions.argtypes = [C.c_int64, _P, _P, _P, _P, C.c_int, _P, C.c_double]
        shim.hs_history_scatter.restype = C.c_int
        shim.hs_history_scatter.argtypes = [C.POINTER(_lib.HistoryArgs)]
        shim.hs_minmax_f32.restype = None
        shim.hs_minmax_f32.argtypes = [C.c_int64, _P, _P, C.POINTER(C.c_float), C.POINTER(C.c_float)]

    def od_analytic_interp(self, ctx, desc, t, n, lon, lat, flags, u, v):
        self.calls.append('od_analytic_interp')
        return self.shim.hs_analytic_interp(desc, t, n, lon, lat, flags, u, v)

    def od_analytic_advect(self, ctx, desc, args):
        self.calls.append('od_analytic_advect')
        return self.shim.hs_analytic_advect(desc, args)

    def od_update_positions(self, ctx, n, lon, lat, xv, yv, f64, moving, dt):
        self.calls.append('od_update_positions')
        self.shim.hs_update_positions(n, lon, lat, xv, yv, f64, moving, dt)
        return 0

    def od_minmax_f32(self, ctx, n, a, b, lo, hi):
        self.calls.append('od_minmax_f32')
        self.shim.hs_minmax_f32(n, a, b, lo, hi)
        return 0

    def od_history_scatter(self, ctx, args):
        self.calls.append('od_history_scatter')
        return self.shim.hs_history_scatter(args)

    def od_last_error(self, ctx):
        return b'hostshim call failed'


class HostEngine:
    def __init__(self):
        import torch
        self.torch = torch
        self.device = torch.device('cpu')
        self.lib = _HostLib(common.hostshim())
        self.ctx = None
        self.math_mode = _lib.OD_MATH_SERIES

    def sync(self):
        pass

    def launches(self):
        return len(self.lib.calls)

    _check = Engine._check
    to_device = Engine.to_device
    empty = Engine.empty
    analytic_interp = Engine.analytic_interp
    analytic_advect = Engine.analytic_advect
    update_positions = Engine.update_positions
    minmax = Engine.minmax
    history_scatter = Engine.history_scatter
