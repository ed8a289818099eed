"""ContinuousReader subclasses (reader_oscillating, a user-written analytical reader) under the drop-in OceanDrift on the host build of
the device sources, against runs of the unmodified reference with the same reader classes (tests/contcases.py,
tests/golden/cont_ref.npz)."""
from datetime import timedelta

import numpy as np
import pytest

import contcases as cc
from hostengine import HostEngine


@pytest.mark.parametrize('case', list(cc.CASES))
def test_continuous_readers_equal_the_reference(case):
    o = cc.run_product(case, engine=HostEngine())
    cc.check(o, case)


@pytest.mark.parametrize('case', list(cc.C2D_CASES))
def test_constant_2d_reader_equals_the_reference(case):
    o = cc.run_c2d_product(case, engine=HostEngine())
    n_act, n_deact = cc.check_c2d(o, case)
    assert n_deact >= 20


def test_reader_queries_equal_the_reference():
    from opendrift_b200.readers import reader_oscillating
    from opendrift_b200.readers.continuous import ContinuousReader
    got = cc.reader_queries(reader_oscillating, ContinuousReader)
    ref = np.load(cc.GOLDEN)
    for k, v in got.items():
        r = ref['query__' + k]
        assert v.shape == r.shape, k
        # (the reference's reader hands back what get_variables returned; the product's values are float32, as Environment stores them)
        assert np.allclose(v, r, rtol=1e-6, atol=1e-7, equal_nan=True), k
        assert np.array_equal(np.isnan(v), np.isnan(r)), k


def test_reference_test_previous_sea_surface_height_reader():
    """tests/models/test_environment.py:66-73: reader_oscillating takes its place in a run (the sea surface height itself is not used
    by this run: a reader for it is accepted only when nothing on the GPU path would read it)."""
    from opendrift_b200.readers import reader_oscillating
    from datetime import datetime
    t0 = datetime(2024, 1, 1)
    r = reader_oscillating.Reader('sea_surface_height', amplitude=1, period=timedelta(hours=6), phase=0, zero_time=t0)
    env, _ = r.get_variables_interpolated(['sea_surface_height'], time=t0 + timedelta(minutes=30), lon=np.array([3.0]), lat=np.array([60.0]))
    assert float(env['sea_surface_height'][0]) == pytest.approx(0.2588, abs=1e-4)     # (the value the reference's test expects)
