"""Gridded readers on Mercator / Lambert-conformal-conic planes (tests/projcases.py) on the host build of the device sources: the
drop-in model classes against runs of the unmodified reference on readers of the same projection; the device projection code
against the oracle; the descriptor's error cases."""
import ctypes as C

import numpy as np
import pytest

import common
import projcases as pc
from hostengine import HostEngine
from test_oracle_proj_conformal import MERC, LCC, STERE_E


@pytest.fixture()
def host_engine(monkeypatch):
    eng = HostEngine()
    import opendrift_b200.engine as E
    import opendrift_b200.models.basemodel as B
    monkeypatch.setattr(E, 'default_engine', lambda device=None: eng)
    monkeypatch.setattr(B, 'default_engine', lambda device=None: eng)
    yield eng


@pytest.mark.parametrize('case', list(pc.CASES))
def test_projected_reader_case_equals_the_reference(case, host_engine):
    o = pc.run_product(case)
    e, dz, moved = pc.check(o, case)
    assert moved > 5e-3
    assert e < 5e-8 and dz <= (1e-9 if 'mixing' in case else 1e-5), (e, dz)


@pytest.mark.parametrize('proj4', MERC + LCC + STERE_E)
def test_device_projection_equals_the_oracle(proj4):
    from opendrift_b200.readers.projection import make_projection
    from oracle.proj_conformal import make
    lib = common.hostshim()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)                   # noqa: E731
    P, O = make_projection(proj4), make(proj4)
    rng = np.random.default_rng(1)
    n = 20000
    lon = O.lon_0 + rng.uniform(-60, 60, n)
    lat = np.clip(getattr(O, 'lat_0', 0.0) + rng.uniform(-40, 40, n), -85, 85)
    if abs(getattr(O, 'lat_0', 0.0)) == 90:
        lat = np.sign(O.lat_0) * rng.uniform(35, 89.5, n)
    ox, oy = O.forward(lon, lat)
    px, py = P(lon, lat)
    assert np.array_equal(px, ox) and np.array_equal(py, oy)      # the product's host-side projection (seeding helper)
    d = P.desc()
    hx, hy, hl, ha = (np.empty(n) for _ in range(4))
    assert lib.hs_proj(C.byref(d), 0, C.c_int64(n), vp(lon), vp(lat), vp(hx), vp(hy)) == 0
    assert lib.hs_proj(C.byref(d), 1, C.c_int64(n), vp(ox), vp(oy), vp(hl), vp(ha)) == 0
    ol, oa = O.inverse(ox, oy)
    assert max(np.max(np.abs(hx - ox)), np.max(np.abs(hy - oy))) < 1e-7           # metres (libm against NumPy)
    assert max(np.max(np.abs(hl - ol)), np.max(np.abs(ha - oa))) < 1e-12          # degrees
    assert np.max(np.abs(ha - lat)) < 1e-11 and np.max(np.abs((hl - lon + 180) % 360 - 180)) < 1e-11


def test_projection_descriptor_errors():
    from opendrift_b200 import _lib
    from opendrift_b200.readers.projection import make_projection
    lib = common.hostshim()
    z = np.zeros(1)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)                   # noqa: E731
    d = make_projection(LCC[1]).desc()
    d.lat_2 = -d.lat_1                                            # a cone with opposite standard parallels
    assert lib.hs_proj(C.byref(d), 0, C.c_int64(1), vp(z), vp(z), vp(z), vp(z)) == 3
    d = make_projection(MERC[0]).desc()
    d.es = 1.5
    assert lib.hs_proj(C.byref(d), 0, C.c_int64(1), vp(z), vp(z), vp(z), vp(z)) == 3
    d.kind = 9
    assert lib.hs_proj(C.byref(d), 0, C.c_int64(1), vp(z), vp(z), vp(z), vp(z)) == 2
    with pytest.raises(NotImplementedError):
        make_projection('+proj=lcc +lat_1=60 +lon_0=0 +units=m')          # no ellipsoid named
    with pytest.raises(NotImplementedError):
        make_projection('+proj=utm +zone=32 +ellps=WGS84')
    assert _lib.OD_PROJ_MERC == 2 and _lib.OD_PROJ_LCC == 3
