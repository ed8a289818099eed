"""The run loop's output buffer on the device (state_to_buffer, basemodel/__init__.py:2384-2499):
`gpu:history = device` keeps the [trajectory, time] block in HBM, one scatter launch per output step, one read-back per
export_buffer_length output steps.  CPU: the scatter of csrc/od_history.cuh compiled for the host, directly and under the
drop-in model (tests/hostengine.py); the result must equal the host-side buffer element for element."""
import ctypes as C

import numpy as np
import pytest

import common
import gyre_common as gc
from opendrift_b200 import _lib


def test_scatter_by_id_host_build():
    lib = common.hostshim()
    lib.hs_history_scatter.restype = C.c_int
    rng = np.random.default_rng(0)
    n_total, ncols, n = 50, 4, 20
    ids = rng.permutation(n_total)[:n].astype(np.int32)
    ids[3] = -1                                   # out-of-range IDs are skipped, not written
    ids[7] = n_total
    lon, lat = rng.uniform(-180, 180, n), rng.uniform(-90, 90, n)
    status = rng.integers(0, 3, n).astype(np.int32)
    for z in (rng.uniform(-50, 0, n).astype(np.float32), rng.uniform(-50, 0, n)):
        bl, ba, bz = (np.full((n_total, ncols), np.nan, np.float32) for _ in range(3))
        bs = np.full((n_total, ncols), -1, np.int32)
        a = _lib.HistoryArgs()
        a.n, a.n_total, a.col, a.ncols, a.z_f64 = n, n_total, 2, ncols, 1 if z.dtype == np.float64 else 0
        a.d_ids, a.d_lon, a.d_lat, a.d_z, a.d_status = (x.ctypes.data for x in (ids, lon, lat, z, status))
        a.d_buf_lon, a.d_buf_lat, a.d_buf_z, a.d_buf_status = (x.ctypes.data for x in (bl, ba, bz, bs))
        assert lib.hs_history_scatter(C.byref(a)) == 0
        ok = (ids >= 0) & (ids < n_total)
        el, ea, ez = (np.full((n_total, ncols), np.nan, np.float32) for _ in range(3))
        es = np.full((n_total, ncols), -1, np.int32)
        el[ids[ok], 2], ea[ids[ok], 2], ez[ids[ok], 2], es[ids[ok], 2] = lon[ok], lat[ok], z[ok], status[ok]
        for got, exp in ((bl, el), (ba, ea), (bz, ez), (bs, es)):
            assert np.array_equal(got, exp, equal_nan=True)
        a.col = ncols
        assert lib.hs_history_scatter(C.byref(a)) == -1


@pytest.mark.parametrize('length', [100, 7, 1])
def test_device_history_equals_host_history(length):
    from hostengine import HostEngine
    from opendrift_b200.models.oceandrift import OceanDrift
    fx = gc.GyreFixture('gyre_rk4')
    hist = {}
    for mode in ('host', 'device'):
        eng = HostEngine()
        o = OceanDrift(loglevel=50, engine=eng)
        o.set_config('environment:fallback:land_binary_mask', 0)
        o.set_config('drift:advection_scheme', fx.scheme)
        o.set_config('gpu:history', mode)
        rd = fx.product_reader()
        o.add_reader(rd)
        # two releases: the second half of the elements starts one second later (rows stay at their fill value until then)
        half = fx.n // 2
        from datetime import timedelta
        o.seed_elements(fx.seed_lon[:half], fx.seed_lat[:half], time=rd.initial_time)
        o.seed_elements(fx.seed_lon[half:], fx.seed_lat[half:], time=rd.initial_time + timedelta(seconds=1))
        o.run(steps=20, time_step=fx.dt, time_step_output=2 * fx.dt, export_buffer_length=length)
        hist[mode] = o.history
        # one housekeeping launch per calculation step (outside + output column + age in one pass) and one for the final state
        assert eng.lib.calls.count('od_bookkeeping') == 21 and eng.lib.calls.count('od_history_scatter') == 0
    h, d = hist['host'], hist['device']
    assert h['time'] == d['time'] and len(h['time']) == 11
    for k in ('lon', 'lat', 'z', 'status'):
        a, b = np.array(h[k]), np.array(d[k])
        assert a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True), k
    assert np.isnan(np.array(d['lon'])[0, fx.n // 2:]).all() and not np.isnan(np.array(d['lon'])[-1]).any()
