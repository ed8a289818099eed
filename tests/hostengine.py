"""TEST INFRASTRUCTURE: an object with the part of opendrift_b200.engine.Engine's surface that the model classes use
when no gridded reader is bound (analytical / constant readers), backed by the host build of the device math
(tests/hostshim).  It lets the CPU suite run the drop-in model classes end to end -- seeding, release, the run loop,
Environment, the reader glue, the ctypes argument structs of Engine itself -- without a GPU.  The methods are Engine's
own (unbound functions re-used), only the library object they call is swapped for an adapter that forwards
od_* calls to the hs_* functions.  Never imported by the product."""
import ctypes as C
import weakref

import numpy as np

import common
from common import HsGroup, HsPair
from opendrift_b200 import _lib
from opendrift_b200.engine import Engine

_P = C.c_void_p


def _np_from_ptr(ptr, n, dtype):
    """A NumPy view of n elements at a raw address (the caller's tensor keeps the memory alive)."""
    addr = ptr.value if isinstance(ptr, C.c_void_p) else int(ptr)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(addr)
    return np.frombuffer(buf, dtype=dtype, count=n)


class _HsChain(C.Structure):
    _fields_ = [('g', C.POINTER(HsGroup) * _lib.OD_MAX_CHAIN), ('t', (HsPair * 3) * _lib.OD_MAX_CHAIN)]


class _HostGroup:
    """The library's field group on the host: geometry, level table, the ring of raw slabs, pair texels on demand."""

    def __init__(self, desc, levels):
        self.desc = _lib.GroupDesc.from_buffer_copy(desc)
        self.levels = None if levels is None else np.ascontiguousarray(levels, dtype=np.float64)
        self.slabs = {}            # (slot, comp) -> float32 [nz, ny, nx]
        self.version = {}          # slot -> counter
        self._tex = {}
        d = self.desc
        g = HsGroup()
        g.ncomp, g.nx, g.ny, g.nz, g.lon_mode, g.wrap_x, g.global_x = d.ncomp, d.nx, d.ny, d.nz, d.lon_mode, d.wrap_x, d.global_x
        g.x0, g.xspan, g.y0, g.yspan = d.x0, d.xspan, d.y0, d.yspan
        g.xmin, g.xmax, g.ymin, g.ymax = d.xmin, d.xmax, d.ymin, d.ymax
        g.fallback[0], g.fallback[1] = d.fallback[0], d.fallback[1]
        g.proj, g.rotate_vectors = d.proj, d.rotate_vectors
        g.z_levels = None if self.levels is None else self.levels.ctypes.data
        self.hs = g

    def pair(self, ts):
        """hs_pair of an od_time_sample: (c0A, c1A, c0B, c1B) texels, as pack_pair*_kernel builds them."""
        pr = HsPair()
        pr.mode, pr.w = ts.mode, ts.w
        if ts.mode == _lib.OD_T_MISSING:
            return pr
        sa = ts.slot_a if ts.mode != _lib.OD_T_SECOND else ts.slot_b
        sb = ts.slot_b if ts.mode == _lib.OD_T_LERP else sa
        key = (sa, sb, self.version.get(sa), self.version.get(sb))
        if key not in self._tex:
            if len(self._tex) > 6:
                self._tex.clear()
            nc = self.desc.ncomp
            self._tex[key] = np.ascontiguousarray(np.stack([self.slabs[(sa, c)] for c in range(nc)] +
                                                           [self.slabs[(sb, c)] for c in range(nc)], axis=-1), dtype=np.float32)
        pr.tex = self._tex[key].ctypes.data
        return pr


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
        HG, HP = C.POINTER(HsGroup), C.POINTER(HsPair)
        shim.hs2_interp.restype = C.c_int
        shim.hs2_interp.argtypes = [HG, HP, C.c_int64, _P, _P, _P, C.c_int, _P, _P]
        shim.hs2_advect.restype = C.c_int
        shim.hs2_advect.argtypes = [C.POINTER(_lib.AdvectArgs), HG, HP, C.POINTER(_HsChain)]
        shim.hs2_step.restype = C.c_int
        shim.hs2_step.argtypes = [C.POINTER(_lib.StepArgs), HG, HP, HG, HP, HG, HP, C.POINTER(_HsChain)]
        shim.hs2_mix.restype = C.c_int
        shim.hs2_mix.argtypes = [C.POINTER(_lib.MixArgs), HG, HP]
        shim.hs2_leeway.restype = C.c_int
        shim.hs2_leeway.argtypes = [C.POINTER(_lib.LeewayArgs), HG, HP, HG, HP]
        shim.hs_stokes.restype = C.c_int        # takes od_stokes_args as is: hs_stokes_args has the same layout
        shim.hs_geod_direct.restype = None
        shim.hs_geod_direct.argtypes = [C.c_int64, _P, _P, _P, _P, _P, _P]
        self.groups = {}

    # -- field groups: the slab ring lives here, NaN fill with the oracle's dilation ------------------------------------
    def od_group_define(self, ctx, gid, desc, levels):
        d = desc._obj if hasattr(desc, '_obj') else desc
        lv = None if levels is None else np.array(levels[:d.nz], dtype=np.float64)
        self.groups[gid] = _HostGroup(d, lv)
        return 0

    def od_group_free(self, ctx, gid):
        self.groups.pop(gid, None)
        return 0

    def od_group_upload(self, ctx, gid, slot, comp, src, on_device):
        g = self.groups[gid]
        d = g.desc
        a = _np_from_ptr(src, d.nx * d.ny * d.nz, np.float32).reshape(d.nz, d.ny, d.nx).copy()
        g.slabs[(slot, comp)] = a
        g.version[slot] = g.version.get(slot, 0) + 1
        return 0

    def od_group_fill_nan(self, ctx, gid, slot, comp, iterations, h_remaining):
        from oracle.advect_port import expand_numpy_array
        a = self.groups[gid].slabs[(slot, comp)]
        for lay in a:
            for _ in range(iterations):
                if not np.isnan(lay).any():
                    break
                expand_numpy_array(lay)
        self.groups[gid].version[slot] += 1
        if h_remaining is not None:
            h_remaining._obj.value = int(np.isnan(a).sum())
        return 0

    def od_group_set_window(self, ctx, gid, desc):
        d = desc._obj if hasattr(desc, '_obj') else desc
        old = self.groups[gid]
        assert (d.ncomp, d.nz, d.n_slots) == (old.desc.ncomp, old.desc.nz, old.desc.n_slots)
        g = _HostGroup(d, old.levels)
        g.version = dict(old.version)            # slot contents are stale: the product re-uploads before it samples
        self.groups[gid] = g
        return 0

    def od_bbox(self, ctx, n, lon, lat, out):
        x, y = _np_from_ptr(lon, n, np.float64), _np_from_ptr(lat, n, np.float64)
        out[0], out[1], out[2], out[3] = np.nanmin(x), np.nanmax(x), np.nanmin(y), np.nanmax(y)
        return 0

    def od_group_set_fallback(self, ctx, gid, f0, f1):
        g = self.groups[gid]
        g.desc.fallback[0], g.desc.fallback[1] = f0, f1
        g.hs.fallback[0], g.hs.fallback[1] = f0, f1
        return 0

    def od_group_touch(self, ctx, gid, slot):
        self.groups[gid].version[slot] = self.groups[gid].version.get(slot, 0) + 1
        return 0

    def _gp(self, gid, ts):
        g = self.groups[gid]
        return C.byref(g.hs), g.pair(ts)

    def od_interp(self, ctx, gid, ts, n, lon, lat, z, flags, out0, out1):
        self.calls.append('od_interp')
        g, pr = self._gp(gid, ts._obj)
        return self.shim.hs2_interp(g, C.byref(pr), n, lon, lat, z, flags, out0, out1)

    def _t3(self, gid, a):
        arr = (HsPair * 3)()
        g = self.groups[gid]
        if not a.d_k1_u:
            arr[0] = g.pair(a.t_start)
        if a.scheme != _lib.OD_EULER:
            arr[1] = g.pair(a.t_mid)
        if a.scheme == _lib.OD_RK4:
            arr[2] = g.pair(a.t_end)
        return arr

    def _chain(self, a):
        ch = _HsChain()
        for k in range(a.n_chain):
            g = self.groups[a.chain_group[k]]
            ch.g[k] = C.pointer(g.hs)
            if not a.d_k1_u:
                ch.t[k][0] = g.pair(a.chain_t[k][0])
            if a.scheme != _lib.OD_EULER:
                ch.t[k][1] = g.pair(a.chain_t[k][1])
            if a.scheme == _lib.OD_RK4:
                ch.t[k][2] = g.pair(a.chain_t[k][2])
        return ch

    def od_advect_current(self, ctx, args):
        self.calls.append('od_advect_current')
        a = args._obj
        t3 = self._t3(a.group_uv, a)
        return self.shim.hs2_advect(args, C.byref(self.groups[a.group_uv].hs), t3, C.byref(self._chain(a)))

    def od_step_oceandrift(self, ctx, args):
        self.calls.append('od_step_oceandrift')
        a = args._obj
        t3 = self._t3(a.cur.group_uv, a.cur)
        gw = tw = gz = tz = None
        if a.group_wind >= 0:
            gw, tw_ = self._gp(a.group_wind, a.t_wind)
            tw = C.byref(tw_)
        if a.group_w >= 0:
            gz, tz_ = self._gp(a.group_w, a.t_w)
            tz = C.byref(tz_)
        return self.shim.hs2_step(args, C.byref(self.groups[a.cur.group_uv].hs), t3, gw, tw, gz, tz, C.byref(self._chain(a.cur)))

    def od_vertical_mixing(self, ctx, args):
        self.calls.append('od_vertical_mixing')
        a = args._obj
        if a.model == _lib.OD_MIX_ENVIRONMENT:
            g, pr = self._gp(a.group_k, a.t_k)
            return self.shim.hs2_mix(args, g, C.byref(pr))
        return self.shim.hs2_mix(args, None, None)

    def od_leeway_step(self, ctx, args):
        self.calls.append('od_leeway_step')
        a = args._obj
        gw, tw = self._gp(a.group_wind, a.t_wind)
        gc, tc = self._gp(a.group_cur, a.t_cur)
        return self.shim.hs2_leeway(args, gw, C.byref(tw), gc, C.byref(tc))

    def od_stokes_drift(self, ctx, args):
        self.calls.append('od_stokes_drift')
        return self.shim.hs_stokes(args)

    def od_geod_fwd(self, ctx, n, lon, lat, az, dist):
        self.calls.append('od_geod_fwd')
        self.shim.hs_geod_direct(n, lon, lat, az, dist, lon, lat)
        return 0

    def od_partition_active(self, ctx, n, status, perm, h_keep):
        st = _np_from_ptr(status, n, np.int32)
        keep, drop = np.where(st == 0)[0], np.where(st != 0)[0]
        _np_from_ptr(perm, n, np.int32)[:] = np.concatenate([keep, drop]).astype(np.int32)
        h_keep._obj.value = len(keep)
        return 0

    def od_permute(self, ctx, n, perm, src, dst, es):
        pm = _np_from_ptr(perm, n, np.int32)
        s_ = _np_from_ptr(src, n * es, np.uint8).reshape(n, es)
        _np_from_ptr(dst, n * es, np.uint8).reshape(n, es)[:] = s_[pm]
        return 0

    def od_unpermute(self, ctx, n, perm, src, dst, es):
        pm = _np_from_ptr(perm, n, np.int32)
        s_ = _np_from_ptr(src, n * es, np.uint8).reshape(n, es)
        _np_from_ptr(dst, n * es, np.uint8).reshape(n, es)[pm] = s_
        return 0

    def od_sort_by_cell(self, ctx, gid, n, lon, lat, z, perm):
        """Any stable ordering by (level, 4x4-cell tile) will do: the device arrays' order is not observable (IDs travel)."""
        d = self.groups[gid].desc
        lo, la = _np_from_ptr(lon, n, np.float64), _np_from_ptr(lat, n, np.float64)
        x = np.mod(lo, 360) if d.lon_mode == _lib.OD_LON_0_360 else np.mod(lo + 180, 360) - 180
        ix = np.clip(np.nan_to_num((x - d.x0) / d.xspan * (d.nx - 1)), 0, d.nx - 1).astype(np.int64) // 4
        iy = np.clip(np.nan_to_num((la - d.y0) / d.yspan * (d.ny - 1)), 0, d.ny - 1).astype(np.int64) // 4
        iz = np.zeros(n, dtype=np.int64)
        lv = self.groups[gid].levels
        if z and lv is not None:
            zz = _np_from_ptr(z, n, np.float32).astype(np.float64)
            order = np.argsort(lv)
            iz = order[np.clip(np.searchsorted(lv[order], zz), 0, len(lv) - 1)]
        key = (iz * ((d.ny + 3) // 4) + iy) * ((d.nx + 3) // 4) + ix
        _np_from_ptr(perm, n, np.int32)[:] = np.argsort(key, kind='stable').astype(np.int32)
        return 0

    def od_sync(self, ctx):
        return 0

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

    def od_bookkeeping(self, ctx, args):
        self.calls.append('od_bookkeeping')
        return self.shim.hs_bookkeeping(args)

    def od_vertical_buoyancy(self, ctx, args):
        self.calls.append('od_vertical_buoyancy')
        return self.shim.hs_vertical_buoyancy(args)

    def od_coastline(self, ctx, args):
        self.calls.append('od_coastline')
        return self.shim.hs2_coastline(args)

    def od_store_previous(self, ctx, n, lon, lat, ids, id_base, n_total, prev_lon, prev_lat):
        self.shim.hs2_store_previous.argtypes = [C.c_int64, _P, _P, _P, C.c_int32, C.c_int64, _P, _P]
        return self.shim.hs2_store_previous(n, lon, lat, ids, id_base, n_total, prev_lon, prev_lat)

    def od_last_error(self, ctx):
        return b'hostshim call failed'


class HostEngine:
    def __init__(self):
        import torch
        self.torch = torch
        self.device = torch.device('cpu')
        self.lib = _HostLib(common.hostshim())
        self.ctx = 1               # truthy: Engine.free_group checks it
        self.math_mode = _lib.OD_MATH_SERIES
        self.groups = weakref.WeakValueDictionary()
        self.dist = None
        self.direction = 1

    # no streams on the host: no prefetch
    def begin_copy_stream(self):
        return False

    def order_after_copies(self):
        pass

    def wait_event(self, ev):
        pass

    enable_distributed = Engine.enable_distributed
    touch = Engine.touch

    def slot_tensor(self, group, slot, comp):
        """The host slab of a ring slot as a torch tensor that shares its memory (target of a gloo broadcast)."""
        g = self.lib.groups[group.gid]
        d = g.desc
        if (slot, comp) not in g.slabs:
            g.slabs[(slot, comp)] = np.zeros((d.nz, d.ny, d.nx), dtype=np.float32)
        return self.torch.from_numpy(g.slabs[(slot, comp)].reshape(-1))

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
    bookkeeping = Engine.bookkeeping
    bbox = Engine.bbox
    vertical_buoyancy = Engine.vertical_buoyancy
    coastline = Engine.coastline
    store_previous = Engine.store_previous
    # gridded readers: Engine's own group management and call wrappers
    add_group = Engine.add_group
    free_group = Engine.free_group
    upload = Engine.upload
    fill_nan = Engine.fill_nan
    interp = Engine.interp
    geod_fwd = Engine.geod_fwd
    _advect_args = Engine._advect_args
    advect_current = Engine.advect_current
    _step_args = Engine._step_args
    step_oceandrift = Engine.step_oceandrift
    leeway_step = Engine.leeway_step
    stokes_drift = Engine.stokes_drift
    vertical_mixing = Engine.vertical_mixing
    partition_active = Engine.partition_active
    permute = Engine.permute
    sort_by_cell = Engine.sort_by_cell
    PROFILES = Engine.PROFILES
    MIX_MODELS = Engine.MIX_MODELS
