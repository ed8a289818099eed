"""Variables without a fallback value: elements leaving the readers' coverage are taken out as 'missing_data' at the top of the loop, and
Runge-Kutta mid-points outside the coverage leave undefined positions behind, as in the reference -- the drop-in OceanDrift on the
host build of the device sources against runs of the unmodified reference (tests/missingcases.py, tests/golden/missing_ref.npz)."""
import pytest

import missingcases as mc
from hostengine import HostEngine


@pytest.mark.parametrize('case', list(mc.CASES))
def test_missing_data_equals_the_reference(case):
    o = mc.run_product(case, engine=HostEngine())
    n_act, n_deact, n_undefined, cats = mc.check(o, case)
    assert n_deact >= 9 and 'missing_data' in cats
    if 'rk' in case and 'wind' not in case:
        assert n_undefined > 0          # (mid-points outside the coverage)
