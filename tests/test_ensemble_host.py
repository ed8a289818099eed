"""Ensemble blocks (tests/enscases.py) on the host build of the device sources: the drop-in OceanDrift against runs of the unmodified
reference on the same ensemble reader -- member i % n_members per element, also when a second reader serves part of the elements;
and the reference's own known answer for ReaderBlock with ensemble members (tests/readers/test_interpolation.py:231-259)."""
import numpy as np
import pytest

import enscases as ec
from hostengine import HostEngine


@pytest.fixture()
def host_engine(monkeypatch):
    eng = HostEngine()
    import opendrift_b200.engine as E
    import opendrift_b200.models.basemodel as B
    monkeypatch.setattr(E, 'default_engine', lambda device=None: eng)
    monkeypatch.setattr(B, 'default_engine', lambda device=None: eng)
    yield eng


@pytest.mark.parametrize('case', list(ec.CASES))
def test_ensemble_case_equals_the_reference(case, host_engine):
    o = ec.run_product(case)
    e, spread = ec.check(o, case)
    assert e < 5e-8, e
    assert spread > 100 * 5e-8          # the members differ by far more than the tolerance


def test_reference_known_answer_for_ensemble_members(host_engine):
    """test_interpolation_ensemble: members 1, 2, 3 (2-D) and 31, 32, 33 (3-D): element 0 -> member 0, 1 -> 1, 3 -> 0."""
    from datetime import datetime
    lon, lat, z = np.linspace(2, 6, 20), np.linspace(58, 61, 16), np.array([0.0, -10.0, -30.0])
    t0 = datetime(2024, 1, 1)
    one2, one3 = np.ones((1, 16, 20), dtype=np.float32), np.ones((1, 3, 16, 20), dtype=np.float32)
    r2 = ec.product_ensemble_reader(lon, lat, None, [t0], {'x_wind': [one2 * 1, one2 * 2, one2 * 3]}, 'e2')
    r3 = ec.product_ensemble_reader(lon, lat, z, [t0], {'sea_water_temperature': [one3 * 31, one3 * 32, one3 * 33]}, 'e3')
    r2.bind(host_engine)
    r3.bind(host_engine)
    x, y, zz = np.linspace(2.5, 5.5, 15), np.linspace(58.5, 60.5, 15), -np.linspace(0, 25, 15)
    v2 = r2.get_variables_interpolated(['x_wind'], time=t0, lon=x, lat=y, z=zz)[0]['x_wind']
    v3 = r3.get_variables_interpolated(['sea_water_temperature'], time=t0, lon=x, lat=y, z=zz)[0]['sea_water_temperature']
    assert (v2[0], v2[1], v2[3]) == (1, 2, 1)
    assert v3[0] == 31 and v3[1] == 32 and abs(v3[3] - 31) < 1e-12
    assert list(v2) == [1 + (i % 3) for i in range(15)]
