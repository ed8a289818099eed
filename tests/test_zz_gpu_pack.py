"""GPU: od_pack_by_owner / od_unpack_records (the particle exchange of the spatial-tile mode) against the torch restatement
in opendrift_b200/sharding.py (stable argsort by owner + column concatenation)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('n,world', [(1, 2), (257, 3), (100003, 8), (1000000, 2)])
def test_pack_by_owner_equals_stable_sort_and_concatenation(n, world):
    import torch
    from opendrift_b200 import sharding
    from opendrift_b200.engine import default_engine
    eng = default_engine()
    g = torch.Generator(device='cpu').manual_seed(n)
    lon = (torch.rand(n, generator=g, dtype=torch.float64) * 12.0 - 1.0)
    lon[::97] = float('nan')                                    # NaN positions go to the last strip, like torch.bucketize
    bounds = sharding.strip_bounds(0.0, 10.0, world)
    cols = {'ID': torch.arange(n, dtype=torch.int32), 'lon': lon, 'lat': torch.rand(n, generator=g, dtype=torch.float64),
            'z': torch.rand(n, generator=g, dtype=torch.float32), 'flag': (torch.arange(n) % 251).to(torch.uint8)}
    dcols = {k: v.to(eng.device) for k, v in cols.items()}
    rec, counts, layout, perm = eng.pack_by_owner(dcols['lon'], bounds, dcols, want_perm=True)
    owner = sharding.strip_owner(lon, bounds)
    order = torch.argsort(owner, stable=True)
    ref, _ = sharding._pack(cols, order)
    assert counts == torch.bincount(owner, minlength=world).tolist() and sum(counts) == n
    assert torch.equal(perm.cpu().to(torch.int64), order)
    assert rec.shape == ref.shape and torch.equal(rec.cpu(), ref)
    back = eng.unpack_records(rec, layout)
    for k in cols:
        a, b = back[k].cpu(), cols[k][order]
        assert a.dtype == b.dtype and torch.equal(torch.nan_to_num(a.to(torch.float64), nan=-7.0), torch.nan_to_num(b.to(torch.float64), nan=-7.0)), k
