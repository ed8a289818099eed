"""GPU: variables without a fallback value -- elements leaving the readers' coverage ('missing_data' at the top of the loop, undefined
positions after Runge-Kutta mid-points outside the coverage) against runs of the unmodified reference (tests/missingcases.py).
Added after the GPU minutes of round 2 were spent -- verified on the host build of the device sources
(tests/test_missing_host.py); it runs after the other GPU tests."""
import pytest

import missingcases as mc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('case', list(mc.CASES))
def test_missing_data_equals_the_reference(case):
    o = mc.run_product(case)
    n_act, n_deact, n_undefined, cats = mc.check(o, case)
    print(case, n_act, n_deact, n_undefined, cats)
    assert n_deact >= 9 and 'missing_data' in cats
