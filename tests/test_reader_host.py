"""Reader.get_variables_interpolated at the reader's own precision on the host build of the device sources: bit-equal to the
unmodified reference's StructuredReader (tests/readercases.py)."""
import readercases as rc
from hostengine import HostEngine


def test_reader_output_equals_the_reference_reader_bit_for_bit():
    assert rc.check(HostEngine()) == 27


def test_projected_reader_output_beside_the_live_reference():
    """get_variables_interpolated of readers on conic / stereographic / Mercator planes, with and without the rotation to east /
    north (rotate_to_proj): unrotated values bit-equal, rotated ones to 1e-9 (the rotation angle comes from a differently organised
    inverse geodesic) and float64 like the reference's."""
    import numpy as np
    import pytest
    from datetime import timedelta
    import projcases as pc
    from oracle import refrun
    if not refrun.available():
        pytest.skip('reference tree not present (GPU box)')
    refrun.setup()
    from opendrift_b200.readers import reader_regular_grid
    eng = HostEngine()
    for case in ('lcc_wgs84_rk4_3d_w', 'stere_wgs84_polar_rk4_3d_w', 'merc_sphere_mixing', 'lcc_sphere_rk2_wind'):
        c, xs, ys, zs, times, fields, lon0, lat0, z0 = pc.setup(case)
        nm = c['readers'][0]
        three_d = nm.startswith('cur3d')
        f = {k: v for k, v in fields[nm].items() if 'velocity' in k and 'upward' not in k}
        rp = reader_regular_grid.Reader(xs, ys, zs if three_d else None, times, f, name=nm, proj4=c['proj4'])
        rp.bind(eng)
        rr = refrun.make_grid_reader(xs, ys, zs if three_d else None, times, f, name=nm, proj4=c['proj4'])
        t = times[0] + timedelta(seconds=1000)
        lon, lat, z = lon0.astype(np.float64), lat0.astype(np.float64), z0.astype(np.float64)
        for rot in (None, '+proj=latlong'):
            a, _ = rp.get_variables_interpolated(list(f), time=t, lon=lon, lat=lat, z=z, rotate_to_proj=rot)
            b, _ = rr.get_variables_interpolated(list(f), time=t, lon=lon, lat=lat, z=z, rotate_to_proj=rot)
            for k in f:
                x, y = np.ma.filled(np.ma.masked_invalid(a[k]), np.nan), np.ma.filled(np.ma.masked_invalid(b[k]), np.nan)
                assert x.dtype == y.dtype, (case, rot, k, x.dtype, y.dtype)
                if rot is None:
                    assert np.array_equal(x, y, equal_nan=True)
                else:
                    assert np.nanmax(np.abs(x - y)) < 1e-9
