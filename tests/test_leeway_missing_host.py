"""Leeway elements leaving the wind reader's coverage are taken out at the top of the loop, before the step's draws, as the reference
does -- the drop-in class on the host build of the device sources against runs of the unmodified reference
(tests/leewaymissing.py, tests/golden/leeway_missing_ref.npz)."""
import pytest

import leewaymissing as lm
from hostengine import HostEngine


@pytest.mark.parametrize('case', list(lm.CASES))
def test_leeway_missing_forcing_equals_the_reference(case):
    o = lm.run_product(case, engine=HostEngine())
    n_act, n_deact, cats = lm.check(o, case)
    assert n_deact > 50 and 'missing_data' in cats
