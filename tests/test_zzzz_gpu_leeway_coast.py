"""GPU: Leeway with coastline stranding against a run of the unmodified reference (tests/coastcases.py).  Added after the GPU
minutes of round 2 were spent -- verified on the host build of the device sources (tests/test_coast_host.py); it runs after the
other GPU tests."""
import pytest

import coastcases as cc

pytestmark = pytest.mark.gpu


def test_leeway_with_coastline_stranding_equals_the_reference():
    o = cc.run_product('leeway_stranding')
    n_act, n_deact, cats = cc.check(o, 'leeway_stranding')
    print('leeway_stranding', n_act, n_deact, cats)
    assert n_deact > 0 and 'stranded' in cats
