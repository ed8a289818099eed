"""GPU: ContinuousReader subclasses (reader_oscillating, a user-written analytical reader; values computed on the host, one round trip
per Runge-Kutta stage) against runs of the unmodified reference (tests/contcases.py).  Added after the GPU minutes of round 2 were
spent -- verified on the host build of the device sources (tests/test_continuous_host.py); it runs after the other GPU tests."""
import pytest

import contcases as cc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('case', list(cc.CASES))
def test_continuous_readers_equal_the_reference(case):
    o = cc.run_product(case)
    print(case, cc.check(o, case))


@pytest.mark.parametrize('case', list(cc.C2D_CASES))
def test_constant_2d_reader_equals_the_reference(case):
    o = cc.run_c2d_product(case)
    print(case, cc.check_c2d(o, case))
