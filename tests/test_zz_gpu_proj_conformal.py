"""GPU: gridded readers on Mercator / Lambert-conformal-conic planes against runs of the unmodified reference (tests/projcases.py)."""
import pytest

import projcases as pc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('case', list(pc.CASES))
def test_projected_reader_case_equals_the_reference(case):
    o = pc.run_product(case)
    e, dz, moved = pc.check(o, case)
    print(case, 'err deg', e, 'dz', dz, 'moved', moved)
    assert moved > 5e-3
    assert e < 5e-8 and dz <= (1e-9 if 'mixing' in case else 1e-5), (e, dz)
