"""ORACLE (test infrastructure, never the product path).

CPU restatement, in NumPy float64, of the WGS84 *direct geodesic problem* that the
reference reaches through ``pyproj.Geod(ellps='WGS84').fwd`` at every position
update (reference call sites: opendrift/models/basemodel/__init__.py:4643-4657
``update_positions``; opendrift/models/physics_methods.py:632-635, 649-652,
663-666 RK mid-points; basemodel/__init__.py:1151-1164 seeding radius).

pyproj is a third-party dependency that is absent from /root/reference and from
this image (reference pins ``pyproj>=2.3`` on ``PROJ<9.8`` in pyproject.toml:18-19;
PROJ's src/geodesic.c is GeographicLib-C).  The algorithm restated here is the
published one: C. F. F. Karney, "Algorithms for geodesics", J. Geodesy 87 (2013)
43-55, eqs. 7-21 with the order-6 series in the third flattening n / the
expansion parameter eps (A1, C1, C1', A3, C3), organised as "line initialisation"
+ "position at distance s12".

Pinning: `oracle/geod_exact.py` evaluates the *exact* elliptic integrals of the
same problem with mpmath quadrature (no series) and `tests/test_oracle_geod.py`
compares the two (<= 1e-12 deg) on random and edge cases; the committed fixture
is tests/golden/geod_mpmath.npz.  The reference's own known answers that touch
this function (tests/models/test_models.py:44-64, tests/models/test_environment.py:30-49,
tests/readers/test_variables.py:107-128) pin it to ~1e-3 deg only and are
replayed in tests/test_reference_known_answers.py.
"""
import numpy as np

WGS84_A = 6378137.0
WGS84_F = 1.0 / 298.257223563

_DEG = np.pi / 180.0
_TINY = np.sqrt(np.finfo(np.float64).tiny)

nA1 = nC1 = nC1p = nA3 = nC3 = 6


def _polyval(coeffs, x):
    y = np.zeros_like(x) + coeffs[0]
    for c in coeffs[1:]:
        y = y * x + c
    return y


# ---- series coefficients (Karney 2013, eqs. 17, 18, 21, 24, 25), order 6 ----
_A1M1 = ([1, 4, 64, 0], 256)
_C1 = [  # (numerator poly in eps^2, denominator) for C1[l]/eps^l
    ([-1, 6, -16], 32),
    ([-9, 64, -128], 2048),
    ([9, -16], 768),
    ([3, -5], 512),
    ([-7], 1280),
    ([-7], 2048),
]
_C1P = [
    ([205, -432, 768], 1536),
    ([4005, -4736, 3840], 12288),
    ([-225, 116], 384),
    ([-7173, 2695], 7680),
    ([3467], 7680),
    ([38081], 61440),
]
# A3 = sum_k A3x[k] eps^k, A3x[k] polynomial in n  (listed highest power of eps first)
_A3 = [
    ([-3], 128),          # eps^5
    ([-2, -3], 64),       # eps^4
    ([-1, -3, -1], 16),   # eps^3
    ([3, -1, -2], 8),     # eps^2
    ([1, -1], 2),         # eps^1
    ([1], 1),             # eps^0
]
# C3[l] = sum_{k>=l} C3x[l][k] eps^k (highest power first within each l)
_C3 = [
    [([3], 128), ([2, 5], 128), ([-1, 3, 3], 64), ([-1, 0, 1], 8), ([-1, 1], 4)],  # l=1: eps^5..eps^1
    [([5], 256), ([1, 3], 128), ([-3, -2, 3], 64), ([1, -3, 2], 32)],               # l=2: eps^5..eps^2
    [([7], 512), ([-10, 9], 384), ([5, -9, 5], 192)],                               # l=3
    [([7], 512), ([-14, 7], 512)],                                                  # l=4
    [([21], 2560)],                                                                 # l=5
]


class Ellipsoid:
    def __init__(self, a=WGS84_A, f=WGS84_F):
        self.a = a
        self.f = f
        self.f1 = 1.0 - f
        self.e2 = f * (2.0 - f)
        self.ep2 = self.e2 / (self.f1 * self.f1)
        self.n = f / (2.0 - f)
        self.b = a * self.f1
        n = np.float64(self.n)
        # A3x[k], k = 0..5 (coefficient of eps^k)
        self.A3x = [float(_polyval(np.array(c, dtype=np.float64), n) / d) for c, d in reversed(_A3)]
        # C3x[l][k] for l = 1..5, k = l..5
        self.C3x = {}
        for l, rows in enumerate(_C3, start=1):
            ks = list(range(5, l - 1, -1))
            for k, (c, d) in zip(ks, rows):
                self.C3x[(l, k)] = float(_polyval(np.array(c, dtype=np.float64), n) / d)


WGS84 = Ellipsoid()


def _ang_normalize(x):
    """IEEE remainder(x, 360) with -180 -> 180 (geodesic.c AngNormalize)."""
    y = x - 360.0 * np.rint(x / 360.0)
    return np.where(y == -180.0, 180.0, y)


def _ang_round(x):
    z = 1.0 / 16.0
    y = np.abs(x)
    y = np.where(y < z, z - (z - y), y)
    return np.copysign(y, x)


def _sincosd(x):
    """sin, cos of x in degrees with exact quadrant reduction (geodesic.c sincosdx)."""
    q = np.rint(x / 90.0)
    r = x - 90.0 * q            # exact for |x| < 2^52
    q = q.astype(np.int64) & 3
    r = r * _DEG
    s, c = np.sin(r), np.cos(r)
    sinx = np.where(q == 0, s, np.where(q == 1, c, np.where(q == 2, -s, -c)))
    cosx = np.where(q == 0, c, np.where(q == 1, -s, np.where(q == 2, -c, s)))
    sinx = np.where(x == 0, x, sinx)     # sign of zero
    return sinx, cosx + 0.0


def _sin_series(sinx, cosx, c):
    """sum_{i=1..n} c[i] sin(2 i x) by Clenshaw summation (c is a list of arrays, c[0] <-> i=1)."""
    ar = 2.0 * (cosx - sinx) * (cosx + sinx)
    n = len(c)
    k = n
    if n & 1:
        k -= 1
        y0 = c[k]
    else:
        y0 = np.zeros_like(sinx)
    y1 = np.zeros_like(sinx)
    while k > 0:
        k -= 1
        y1 = ar * y0 - y1 + c[k]
        k -= 1
        y0 = ar * y1 - y0 + c[k]
    return 2.0 * sinx * cosx * y0


def _series_coeffs(table, eps):
    eps2 = eps * eps
    out = []
    d = eps
    for c, den in table:
        out.append(d * _polyval(np.array(c, dtype=np.float64), eps2) / den)
        d = d * eps
    return out


def direct(lon1, lat1, azi1, s12, ell=WGS84, return_azi2=False):
    """WGS84 direct problem, all angles in degrees, distance in metres.

    Returns (lon2, lat2) [, azi2] with lon2 normalised to [-180, 180] as PROJ's
    geod_direct does (no LONG_UNROLL).
    """
    lon1, lat1, azi1, s12 = np.broadcast_arrays(
        *(np.asarray(v, dtype=np.float64) for v in (lon1, lat1, azi1, s12)))
    azi1 = _ang_normalize(azi1)
    salp1, calp1 = _sincosd(_ang_round(azi1))
    lat1 = np.where(np.abs(lat1) > 90.0, np.nan, lat1)       # LatFix
    sbet1, cbet1 = _sincosd(_ang_round(lat1))
    sbet1 = sbet1 * ell.f1
    r = np.hypot(sbet1, cbet1)
    sbet1, cbet1 = sbet1 / r, cbet1 / r
    cbet1 = np.maximum(_TINY, cbet1)

    salp0 = salp1 * cbet1
    calp0 = np.hypot(calp1, salp1 * sbet1)
    ssig1 = sbet1
    somg1 = salp0 * sbet1
    csig1 = np.where((sbet1 != 0) | (calp1 != 0), cbet1 * calp1, 1.0)
    comg1 = csig1
    r = np.hypot(ssig1, csig1)
    ssig1, csig1 = ssig1 / r, csig1 / r

    k2 = calp0 * calp0 * ell.ep2
    eps = k2 / (2.0 * (1.0 + np.sqrt(1.0 + k2)) + k2)
    eps2 = eps * eps

    # A1 - 1
    t = _polyval(np.array(_A1M1[0], dtype=np.float64), eps2) / _A1M1[1]
    A1m1 = (t + eps) / (1.0 - eps)
    C1a = _series_coeffs(_C1, eps)
    B11 = _sin_series(ssig1, csig1, C1a)
    s, c = np.sin(B11), np.cos(B11)
    stau1 = ssig1 * c + csig1 * s
    ctau1 = csig1 * c - ssig1 * s
    C1pa = _series_coeffs(_C1P, eps)

    # A3, C3
    A3 = np.zeros_like(eps)
    for k in range(5, -1, -1):
        A3 = A3 * eps + ell.A3x[k]
    C3a = []
    mult = np.ones_like(eps)
    for l in range(1, 6):
        mult = mult * eps
        p = np.zeros_like(eps)
        for k in range(5, l - 1, -1):
            p = p * eps + ell.C3x[(l, k)]
        C3a.append(mult * p)
    A3c = -ell.f * salp0 * A3
    B31 = _sin_series(ssig1, csig1, C3a)

    # position
    tau12 = s12 / (ell.b * (1.0 + A1m1))
    s, c = np.sin(tau12), np.cos(tau12)
    B12 = -_sin_series(stau1 * c + ctau1 * s, ctau1 * c - stau1 * s, C1pa)
    sig12 = tau12 - (B12 - B11)
    ssig12, csig12 = np.sin(sig12), np.cos(sig12)
    ssig2 = ssig1 * csig12 + csig1 * ssig12
    csig2 = csig1 * csig12 - ssig1 * ssig12
    sbet2 = calp0 * ssig2
    cbet2 = np.hypot(salp0, calp0 * csig2)
    degenerate = cbet2 == 0
    cbet2 = np.where(degenerate, _TINY, cbet2)
    csig2 = np.where(degenerate, _TINY, csig2)
    somg2 = salp0 * ssig2
    comg2 = csig2
    omg12 = np.arctan2(somg2 * comg1 - comg2 * somg1,
                       comg2 * comg1 + somg2 * somg1)
    lam12 = omg12 + A3c * (sig12 + (_sin_series(ssig2, csig2, C3a) - B31))
    lon12 = lam12 / _DEG
    lon2 = _ang_normalize(_ang_normalize(lon1) + _ang_normalize(lon12))
    lat2 = np.arctan2(sbet2, ell.f1 * cbet2) / _DEG
    if return_azi2:
        azi2 = np.arctan2(salp0, calp0 * csig2) / _DEG
        return lon2, lat2, azi2
    return lon2, lat2


def inverse_short(lon1, lat1, lon2, lat2, ell=WGS84, iterations=4):
    """Inverse geodesic problem for SHORT lines (up to a few hundred km, away from the poles): forward azimuth at
    point 1 (degrees) and distance (m).

    The reference needs it in one place on this path: BaseReader.rotate_vectors (readers/basereader/variables.py:59-109)
    asks ``Geod.inv`` for the azimuth of a 10 m (projected readers) or 0.1 deg (rotated-pole readers) line along the
    reader's y axis.  PROJ solves the general inverse problem with Karney's Newton iteration on the azimuth; for a short
    line the same solution is reached by inverting the pinned direct solution (``direct`` above): start from the
    mid-latitude metric, then correct the (north, east) displacement with the miss of the direct solution until it
    vanishes (the miss shrinks by (s/a)^2 per pass; 2 passes reach double round-off for the lines used here).
    """
    lon1 = np.asarray(lon1, dtype=np.float64)
    lat1 = np.asarray(lat1, dtype=np.float64)
    lon2 = np.asarray(lon2, dtype=np.float64)
    lat2 = np.asarray(lat2, dtype=np.float64)

    def metric(lat):
        sp = np.sin(lat * _DEG)
        w2 = 1.0 - ell.e2 * sp * sp
        w = np.sqrt(w2)
        return ell.a * (1.0 - ell.e2) / (w2 * w), ell.a / w * np.cos(lat * _DEG)

    def wrap(d):
        return d - 360.0 * np.round(d / 360.0)

    dlat = lat2 - lat1
    dlon = wrap(lon2 - lon1)
    m_mid, n_mid = metric(0.5 * (lat1 + lat2))
    north = m_mid * dlat * _DEG
    east = n_mid * dlon * _DEG
    # azimuth at point 1 = azimuth at the mid-point minus half the meridian convergence
    conv = dlon * _DEG * np.sin(0.5 * (lat1 + lat2) * _DEG)
    az_mid = np.arctan2(east, north)
    s = np.hypot(east, north)
    az = az_mid - 0.5 * conv
    north, east = s * np.cos(az), s * np.sin(az)
    m2, n2 = metric(lat2)
    for _ in range(iterations):
        s = np.hypot(east, north)
        az = np.arctan2(east, north)
        lo, la, az2 = direct(lon1, lat1, az / _DEG, s, ell, return_azi2=True)
        # miss at point 2, expressed in the local frame there and turned back to the frame of point 1
        rn = m2 * (lat2 - la) * _DEG
        re = n2 * wrap(lon2 - lo) * _DEG
        turn = (az2 * _DEG) - az
        c, sn = np.cos(turn), np.sin(turn)
        north = north + (rn * c + re * sn)
        east = east + (-rn * sn + re * c)
    s = np.hypot(east, north)
    az = np.arctan2(east, north)
    lo, la, az2 = direct(lon1, lat1, az / _DEG, s, ell, return_azi2=True)
    return az / _DEG, az2, s


class Geod:
    """Minimal stand-in for ``pyproj.Geod`` (fwd only + short-line inv) used when the
    reference is imported in this container (oracle/refrun.py)."""

    def __init__(self, ellps='WGS84', **kw):
        if ellps != 'WGS84':
            raise NotImplementedError(ellps)
        self.ell = WGS84

    def fwd(self, lons, lats, az, dist, radians=False):
        scalar = np.isscalar(lons)
        lons = np.asarray(lons, dtype=np.float64)
        lats = np.asarray(lats, dtype=np.float64)
        az = np.asarray(az, dtype=np.float64)
        dist = np.asarray(dist, dtype=np.float64)
        if radians:
            lons, lats, az = np.degrees(lons), np.degrees(lats), np.degrees(az)
        lon2, lat2, azi2 = direct(lons, lats, az, dist, self.ell, return_azi2=True)
        back = np.where(azi2 > 0, azi2 - 180.0, azi2 + 180.0)
        if radians:
            lon2, lat2, back = np.radians(lon2), np.radians(lat2), np.radians(back)
        if scalar:
            return float(lon2), float(lat2), float(back)
        return lon2, lat2, back

    def inv(self, lons1, lats1, lons2, lats2, radians=False):
        """Forward azimuth, back azimuth, distance (pyproj.Geod.inv) -- short lines only, see inverse_short."""
        scalar = np.isscalar(lons1)
        a = [np.asarray(v, dtype=np.float64) for v in (lons1, lats1, lons2, lats2)]
        if radians:
            a = [np.degrees(v) for v in a]
        az1, az2, s = inverse_short(*a, ell=self.ell)
        back = np.where(az2 > 0, az2 - 180.0, az2 + 180.0)
        if radians:
            az1, back = np.radians(az1), np.radians(back)
        if scalar:
            return float(az1), float(back), float(s)
        return az1, back, s
