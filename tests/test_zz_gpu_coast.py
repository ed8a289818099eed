"""GPU: coastline interaction against the land_binary_mask of a gridded reader (tests/coastcases.py: runs of the unmodified
reference), and Reader.get_variables_interpolated at the reader's own precision, bit-equal to the reference's reader
(tests/readercases.py)."""
import pytest

import coastcases as cc
import readercases as rc

pytestmark = pytest.mark.gpu


# (the Leeway case was added after the round's GPU minutes were spent: it runs last, in tests/test_zzzz_gpu_leeway_coast.py)
@pytest.mark.parametrize('case', [c for c in cc.CASES if not c.startswith('leeway')])
def test_coastline_case_equals_the_reference(case):
    o = cc.run_product(case)
    n_act, n_deact, cats = cc.check(o, case)
    assert case in ('seafloor_previous', 'previous_ocean_only') or (n_deact > 0 and len(cats) > 1)
    print(case, n_act, n_deact, cats)


def test_reader_output_equals_the_reference_reader_bit_for_bit():
    from opendrift_b200.engine import default_engine
    assert rc.check(default_engine()) == 27
