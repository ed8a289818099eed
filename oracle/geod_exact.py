"""ORACLE VALIDATOR (test infrastructure).

Series-free evaluation of the ellipsoidal direct geodesic problem with mpmath
(40 digits): the exact integrals of Karney (2013) eqs. 7-8,

    s/b   = int_0^sigma sqrt(1 + k^2 sin^2 s) ds
    lam   = omega - f sin(alp0) int_0^sigma (2-f) / (1 + (1-f) sqrt(1 + k^2 sin^2 s)) ds

with sigma2 found by Newton iteration.  It shares no series coefficient with
oracle/geod_karney.py, so agreement to 1e-12 deg pins both the algorithm and the
coefficients typed there.  Running this file regenerates tests/golden/geod_mpmath.npz.
"""
import numpy as np
import mpmath as mp

mp.mp.dps = 40
A = mp.mpf(6378137)
F = 1 / mp.mpf('298.257223563')


def direct_exact(lon1, lat1, azi1, s12):
    f = F
    b = A * (1 - f)
    e2 = f * (2 - f)
    ep2 = e2 / (1 - f) ** 2
    d = mp.pi / 180
    phi1, alp1 = mp.mpf(lat1) * d, mp.mpf(azi1) * d
    bet1 = mp.atan((1 - f) * mp.tan(phi1)) if abs(lat1) != 90 else mp.sign(lat1) * mp.pi / 2
    sb1, cb1 = mp.sin(bet1), mp.cos(bet1)
    sa1, ca1 = mp.sin(alp1), mp.cos(alp1)
    sa0 = sa1 * cb1
    ca0 = mp.hypot(ca1, sa1 * sb1)
    sig1 = mp.atan2(sb1, ca1 * cb1) if (sb1 != 0 or ca1 != 0) else mp.mpf(0)
    omg1 = mp.atan2(sa0 * mp.sin(sig1), mp.cos(sig1))
    k2 = ep2 * ca0 ** 2

    def g(s):
        return mp.sqrt(1 + k2 * mp.sin(s) ** 2)

    target = mp.mpf(s12) / b
    sig2 = sig1 + target           # first guess
    for _ in range(60):
        val = mp.quad(g, [sig1, sig2]) - target
        step = val / g(sig2)
        sig2 -= step
        if abs(step) < mp.mpf(10) ** (-35):
            break
    sb2 = ca0 * mp.sin(sig2)
    cb2 = mp.hypot(sa0, ca0 * mp.cos(sig2))
    lat2 = mp.atan2(sb2, (1 - f) * cb2) / d
    omg2 = mp.atan2(sa0 * mp.sin(sig2), mp.cos(sig2))
    # multiples of 2*pi in omg12 are irrelevant: lon2 is reduced mod 360 below
    omg12 = omg2 - omg1
    I3 = mp.quad(lambda s: (2 - f) / (1 + (1 - f) * g(s)), [sig1, sig2])
    lam12 = omg12 - f * sa0 * I3
    lon2 = mp.mpf(lon1) + lam12 / d
    lon2 = lon2 - 360 * mp.floor((lon2 + 180) / 360)      # [-180, 180)
    return float(lon2), float(lat2)


def make_cases(seed=20260924, n_random=1500):
    rng = np.random.default_rng(seed)
    lon = rng.uniform(-180, 180, n_random)
    lat = rng.uniform(-89, 89, n_random)
    azi = rng.uniform(-180, 180, n_random)
    # mix of step-sized (0..5 km), regional (0..200 km) and long (up to 5000 km) lines
    s = np.concatenate([rng.uniform(0, 5e3, n_random // 3),
                        rng.uniform(0, 2e5, n_random // 3),
                        rng.uniform(0, 5e6, n_random - 2 * (n_random // 3))])
    s[::7] *= -1.0                                       # backward runs use negative distance
    edge = np.array([
        # lon, lat, azi, s
        [4.0, 60.0, 0.0, 7200.0],        # tests/models/test_models.py:44-64 (1 m/s north, 2 h)
        [3.0, 60.0, 90.0, 3600.0],       # tests/models/test_environment.py:30-49
        [4.0, 60.0, 225.0, 5400.0],      # tests/readers/test_variables.py:107-128 (0.02*5 m/s*15 h)
        [4.0, 60.0, 45.0, 5400.0],
        [0.0, 0.0, 90.0, 1000.0],        # equatorial
        [0.0, 0.0, 0.0, 1000.0],         # meridional from equator
        [10.0, 89.9, 30.0, 50000.0],     # over the pole region
        [179.99, 10.0, 90.0, 5000.0],    # across the dateline
        [-179.99, -10.0, -90.0, 5000.0],
        [5.0, 57.0, 123.0, 0.0],         # zero distance
        [5.0, 57.0, 180.0, 300.0],
        [5.0, 57.0, -180.0, 300.0],
        [5.0, -57.0, 1e-9, 300.0],
    ])
    lon = np.concatenate([lon, edge[:, 0]])
    lat = np.concatenate([lat, edge[:, 1]])
    azi = np.concatenate([azi, edge[:, 2]])
    s = np.concatenate([s, edge[:, 3]])
    return lon, lat, azi, s


if __name__ == '__main__':
    import os
    lon, lat, azi, s = make_cases()
    out = np.array([direct_exact(*c) for c in zip(lon, lat, azi, s)])
    path = os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'geod_mpmath.npz')
    np.savez_compressed(path, lon1=lon, lat1=lat, azi1=azi, s12=s, lon2=out[:, 0], lat2=out[:, 1])
    print('wrote', path, len(lon), 'cases')
