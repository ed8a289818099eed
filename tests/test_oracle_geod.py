"""The geodesic oracle (oracle/geod_karney.py) against the series-free mpmath evaluation of the exact
geodesic integrals (tests/golden/geod_mpmath.npz, written by oracle/geod_exact.py) and against the
reference's own known answers that touch pyproj.Geod.fwd."""
import os

import numpy as np

from oracle import geod_karney as gk
from common import GOLDEN


def _load():
    return np.load(os.path.join(GOLDEN, 'geod_mpmath.npz'))


def test_karney_vs_exact_integrals():
    g = _load()
    lon2, lat2 = gk.direct(g['lon1'], g['lat1'], g['azi1'], g['s12'])
    dlon = (lon2 - g['lon2'] + 180.0) % 360.0 - 180.0
    assert np.abs(lat2 - g['lat2']).max() < 1e-12
    assert np.abs(dlon * np.cos(np.radians(g['lat2']))).max() < 1e-12
    assert np.abs(dlon).max() < 1e-11          # includes a start 0.1 deg from the pole


def test_reference_known_answers():
    # tests/models/test_models.py:44-64: 1 m/s northward for 2 h from (4E, 60N): lat + 0.0646
    lon, lat = gk.direct(4.0, 60.0, 0.0, 7200.0)
    assert abs(lat - 60.0646) < 5e-4 and abs(lon - 4.0) < 1e-12
    # tests/models/test_environment.py:30-49: 1 m/s eastward for 1 h from (3E, 60N): lon ~ 3.0645
    lon, lat = gk.direct(3.0, 60.0, 90.0, 3600.0)
    assert abs(lon - 3.0645) < 3e-3 * 1e-1
    # tests/readers/test_variables.py:107-128: wind 5 m/s towards 225 / 45 deg, wdf 0.02, 15 h
    lon, lat = gk.direct(4.0, 60.0, 225.0, 0.02 * 5 * 15 * 3600)
    assert abs(lon - 3.932) < 1e-3 and abs(lat - 59.966) < 1e-3
    lon, lat = gk.direct(4.0, 60.0, 45.0, 0.02 * 5 * 15 * 3600)
    assert abs(lon - 4.068) < 1e-3 and abs(lat - 60.034) < 1e-3


def test_edge_cases():
    # zero distance is the identity to round-off; negative distance = opposite azimuth
    lon, lat = gk.direct(5.0, 57.0, 123.0, 0.0)
    assert abs(lon - 5.0) < 1e-13 and abs(lat - 57.0) < 1e-13
    a = gk.direct(5.0, 57.0, 30.0, -1000.0)
    b = gk.direct(5.0, 57.0, -150.0, 1000.0)
    assert np.allclose(a, b, atol=1e-13)
    # longitude is returned in [-180, 180]
    lon, _ = gk.direct(179.99, 10.0, 90.0, 5000.0)
    assert -180.0 <= lon < -179.9
    # NaN in, NaN out
    assert np.isnan(gk.direct(np.nan, 10.0, 90.0, 5000.0)[0])
    # fake-pyproj front end: arrays, scalars, back azimuth convention
    lo, la, back = gk.Geod().fwd(np.array([4.0]), np.array([60.0]), np.array([0.0]), np.array([7200.0]))
    assert abs(abs(back[0]) - 180.0) < 1e-9
