"""GPU: the specialised step kernel (csrc/od_spec.cuh, OD_OPT_SPEC) against the general step kernel -- the same runs with the
option on and off agree BIT FOR BIT (positions and depths), on the benchmarked configurations at 1e5 particles and on the
edge cases of tests/test_spec_host.py (particles on the rim of the block, the antimeridian, polar moves)."""
import numpy as np
import pytest

import bigcases as bc
from test_spec_host import _edge_model

pytestmark = pytest.mark.gpu


def _run(make, on):
    from opendrift_b200.engine import default_engine
    eng = default_engine()
    eng.set_spec(on)
    try:
        o = make()
        o.run(**o._test_run_args)
        return (np.asarray(o.elements.lon).copy(), np.asarray(o.elements.lat).copy(), np.asarray(o.elements.z).copy(),
                np.asarray(o.elements.ID).copy())
    finally:
        eng.set_spec(True)


def _same(a, b):
    assert np.array_equal(a[3], b[3])
    for x, y in zip(a[:3], b[:3]):
        assert np.array_equal(x, y, equal_nan=True)


@pytest.mark.parametrize('kind,sort', [('cfg2', 0), ('cfg2', 20), ('cfg4', 0)])
def test_specialised_kernel_equals_general_kernel_on_benchmarked_configurations(kind, sort):
    if kind not in bc.KINDS:
        pytest.skip('fixture not generated')
    c = bc.BigCase(kind)

    def make():
        o = c.model(**{'gpu:sort_interval_steps': sort})
        o._test_run_args = dict(steps=c.steps, time_step=c.dt, time_step_output=c.steps * c.dt)
        return o

    a, b = _run(make, True), _run(make, False)
    _same(a, b)


@pytest.mark.parametrize('pm180,polar,seed', [(False, False, 1), (True, False, 2), (False, True, 3)])
def test_specialised_kernel_equals_general_kernel_on_edges(pm180, polar, seed):
    a = _run(lambda: _edge_model(pm180, seed, polar), True)
    b = _run(lambda: _edge_model(pm180, seed, polar), False)
    _same(a, b)
