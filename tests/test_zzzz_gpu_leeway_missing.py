"""GPU: Leeway elements leaving the wind reader's coverage (tests/leewaymissing.py) against runs of the unmodified reference.  Added
after the GPU minutes of round 2 were spent -- verified on the host build of the device sources
(tests/test_leeway_missing_host.py); it runs after the other GPU tests."""
import pytest

import leewaymissing as lm

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('case', list(lm.CASES))
def test_leeway_missing_forcing_equals_the_reference(case):
    o = lm.run_product(case)
    n_act, n_deact, cats = lm.check(o, case)
    print(case, n_act, n_deact, cats)
    assert n_deact > 50 and 'missing_data' in cats
