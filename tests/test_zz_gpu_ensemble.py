"""GPU: ensemble blocks (tests/enscases.py) against runs of the unmodified reference."""
import pytest

import enscases as ec

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('case', list(ec.CASES))
def test_ensemble_case_equals_the_reference(case):
    o = ec.run_product(case)
    e, spread = ec.check(o, case)
    print(case, 'err deg', e, 'member spread', spread)
    assert e < 5e-8 and spread > 100 * 5e-8
