"""Environment.get_environment with `profiles=` on the host build of the device sources against the unmodified reference
(tests/envprofcases.py, tests/golden/envprof_ref.npz): recarray, missing mask and profiles bit for bit."""
import numpy as np

import envprofcases as ec
from hostengine import HostEngine


def test_get_environment_profiles_equal_the_reference():
    from opendrift_b200.models.oceandrift import OceanDrift
    from opendrift_b200.readers import reader_regular_grid
    got = ec.run_all(OceanDrift, lambda lon, lat, z, t, f, name: reader_regular_grid.Reader(lon, lat, z, t, f, name=name), engine=HostEngine())
    ref = np.load(ec.GOLDEN)
    assert set(got) == set(ref.files)
    for k in sorted(got):
        assert got[k].shape == ref[k].shape, k
        assert np.array_equal(got[k], ref[k], equal_nan=True), (k, float(np.nanmax(np.abs(got[k].astype(float) - ref[k].astype(float)))))
