"""GPU: the device-resident output buffer (`gpu:history = device`, od_history_scatter) against the host-side buffer, on a
gridded fixture run through the drop-in classes and on the double gyre.  (Sorts last: written after the round's GPU
minutes were spent; host-verified in tests/test_history.py.)"""
import numpy as np
import pytest

import common
import gyre_common as gc

pytestmark = pytest.mark.gpu


def _histories(make, run_kw):
    out = {}
    for mode in ('host', 'device'):
        o = make()
        o.set_config('gpu:history', mode)
        o.run(**run_kw)
        out[mode] = (o.history, np.asarray(o.elements.lon), np.asarray(o.elements.lat))
    return out


@pytest.mark.parametrize('length', [100, 3])
def test_device_history_gridded_fixture(length):
    from test_gpu_dropin import _model
    fx = common.Fixture('rk4_3d_full')
    res = _histories(lambda: _model(fx), dict(steps=fx.steps, time_step=fx.dt, time_step_output=fx.dt, export_buffer_length=length))
    (h, hl, ha), (d, dl, da) = res['host'], res['device']
    assert np.array_equal(hl, dl) and np.array_equal(ha, da)              # the buffer mode does not touch the physics
    assert h['time'] == d['time'] and len(d['time']) == fx.steps + 1
    for k in ('lon', 'lat', 'z', 'status'):
        a, b = np.array(h[k]), np.array(d[k])
        assert a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True), k


def test_device_history_double_gyre_and_errors():
    from opendrift_b200.models.oceandrift import OceanDrift
    fx = gc.GyreFixture('gyre_rk4')

    def make():
        o = OceanDrift(loglevel=50)
        o.set_config('environment:fallback:land_binary_mask', 0)
        o.set_config('drift:advection_scheme', fx.scheme)
        rd = fx.product_reader()
        o.add_reader(rd)
        o.seed_elements(fx.seed_lon, fx.seed_lat, time=rd.initial_time)
        return o
    res = _histories(make, dict(steps=20, time_step=fx.dt, time_step_output=2 * fx.dt, export_buffer_length=4))
    (h, _, _), (d, _, _) = res['host'], res['device']
    for k in ('lon', 'lat', 'z', 'status'):
        assert np.array_equal(np.array(h[k]), np.array(d[k]), equal_nan=True), k
    from opendrift_b200.engine import default_engine
    eng = default_engine()
    torch = eng.torch
    ids = torch.arange(4, dtype=torch.int32, device=eng.device)
    pos = torch.zeros(4, dtype=torch.float64, device=eng.device)
    z = torch.zeros(4, dtype=torch.float32, device=eng.device)
    bufs = tuple(torch.zeros((4, 2), dtype=torch.float32, device=eng.device) for _ in range(3)) + (
        torch.zeros((4, 2), dtype=torch.int32, device=eng.device),)
    with pytest.raises(RuntimeError, match='bad sizes'):
        eng.history_scatter(ids, pos, pos, z, ids, bufs, 2)


def test_page_locked_output_buffer_equals_the_pageable_one():
    """More output columns than one device block: blocks travel asynchronously into a page-locked buffer, two device blocks in
    turn; the result is what the synchronous, pageable path gives."""
    from test_gpu_dropin import _model
    fx = common.Fixture('rk4_3d_full')
    out = {}
    for name, pinned in (('pinned', 16 * 2 ** 30), ('pageable', 0)):
        o = _model(fx)
        o.set_config('gpu:history_pinned_bytes', pinned)
        o.run(steps=fx.steps, time_step=fx.dt, time_step_output=fx.dt, export_buffer_length=3)
        assert (o._hist_pinned is not None) == (name == 'pinned')
        out[name] = o.history
    a, b = out['pinned'], out['pageable']
    assert a['time'] == b['time'] and len(a['time']) == fx.steps + 1
    for k in ('lon', 'lat', 'z', 'status'):
        assert np.array_equal(np.array(a[k]), np.array(b[k]), equal_nan=True), k
    assert np.isfinite(np.array(a['lon'])).all()
