"""Coastline interaction (tests/coastcases.py) on the host build of the device sources: the drop-in OceanDrift with a gridded
land_binary_mask reader against runs of the unmodified reference -- surviving elements, deactivated elements in the reference's
order, their status categories ('stranded', 'seeded_on_land', 'missing_data') and positions."""
import numpy as np
import pytest

import coastcases as cc
from hostengine import HostEngine


@pytest.fixture()
def host_engine(monkeypatch):
    eng = HostEngine()
    import opendrift_b200.engine as E
    import opendrift_b200.models.basemodel as B
    monkeypatch.setattr(E, 'default_engine', lambda device=None: eng)
    monkeypatch.setattr(B, 'default_engine', lambda device=None: eng)
    yield eng


@pytest.mark.parametrize('case', list(cc.CASES))
def test_coastline_case_equals_the_reference(case, host_engine):
    o = cc.run_product(case)
    n_act, n_deact, cats = cc.check(o, case)
    assert case in ('seafloor_previous', 'previous_ocean_only') or (n_deact > 0 and len(cats) > 1)
    assert 'od_coastline' in host_engine.lib.calls


def test_nearest_sampling_equals_the_reference_interpolator(host_engine):
    """od_interp + OD_INTERP_NEAREST against Nearest2DInterpolator's arithmetic (interpolators.py:26-40) restated in NumPy, float64
    and float32 positions, incl. points on and beyond the last grid point."""
    import torch
    from opendrift_b200.readers import reader_regular_grid
    from datetime import datetime
    rng = np.random.default_rng(3)
    lon, lat = np.linspace(2.0, 5.0, 23), np.linspace(58.0, 60.0, 17)
    mask = (rng.random((17, 23)) < 0.4).astype(np.float32)
    t0 = datetime(2024, 1, 1)
    r = reader_regular_grid.Reader(lon, lat, None, [t0], {'land_binary_mask': mask[None]}, name='mask')
    r.bind(host_engine)
    g, _ = r.group_of('land_binary_mask')
    n = 5000
    x = rng.uniform(2.0, 5.0, n)
    y = rng.uniform(58.0, 60.0, n)
    x[:50], y[50:100] = 5.0, 60.0
    x[100:150], y[150:200] = 2.0, 58.0
    for f32 in (False, True):
        xx = x.astype(np.float32) if f32 else x
        yy = y.astype(np.float32) if f32 else y
        got = host_engine.interp(g, t0, host_engine.to_device(xx.astype(np.float64)), host_engine.to_device(yy.astype(np.float64)),
                                 None, pos_f32=f32, raw=True, nearest=True)[0].cpu().numpy()
        xg, yg = np.asarray(r.lon_grid if hasattr(r, 'lon_grid') else lon, dtype=np.float32), np.asarray(lat, dtype=np.float32)
        xi = np.round((xx - xg.min()) / (xg.max() - xg.min()) * len(xg)).astype(np.uint32)
        yi = np.round((yy - yg.min()) / (yg.max() - yg.min()) * len(yg)).astype(np.uint32)
        xi[xi >= len(xg)] = len(xg) - 1
        yi[yi >= len(yg)] = len(yg) - 1
        cov = (xx >= xg.min()) & (xx <= xg.max()) & (yy >= yg.min()) & (yy <= yg.max())
        want = np.where(cov, mask[yi, xi], np.nan)
        assert np.array_equal(got, want, equal_nan=True)
