"""GPU: Environment.get_environment with `profiles=` against the unmodified reference (tests/envprofcases.py).  Added after the GPU
minutes of round 2 were spent -- verified on the host build of the device sources (tests/test_envprof_host.py); it runs after the
other GPU tests."""
import numpy as np
import pytest

import envprofcases as ec

pytestmark = pytest.mark.gpu


def test_get_environment_profiles_equal_the_reference():
    from opendrift_b200.models.oceandrift import OceanDrift
    from opendrift_b200.readers import reader_regular_grid
    got = ec.run_all(OceanDrift, lambda lon, lat, z, t, f, name: reader_regular_grid.Reader(lon, lat, z, t, f, name=name))
    ref = np.load(ec.GOLDEN)
    for k in sorted(got):
        assert np.array_equal(got[k], ref[k], equal_nan=True), k
