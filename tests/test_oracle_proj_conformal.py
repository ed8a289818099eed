"""oracle/proj_conformal.py (the fake pyproj's Mercator and Lambert conformal conic) against the published closed forms evaluated
with mpmath at 40 digits (Snyder 1987: Mercator 7-7 / 7-8, Lambert conformal conic 15-7 .. 15-10 with 14-15 and 15-9; the
spherical forms are their e = 0 limits), and round trips."""
import numpy as np
import pytest

MERC = ['+proj=merc +lon_0=5 +lat_ts=60 +ellps=WGS84 +units=m +no_defs',
        '+proj=merc +lon_0=-20 +k_0=0.9996 +x_0=3000 +y_0=-1000 +R=6371000 +units=m +no_defs',
        '+proj=merc +lon_0=0 +a=6378137 +rf=298.257222101 +units=m +no_defs']
LCC = ['+proj=lcc +lat_1=63 +lat_2=63 +lat_0=63 +lon_0=15 +R=6371000 +units=m +no_defs',
       '+proj=lcc +lat_1=30 +lat_2=60 +lat_0=45 +lon_0=10 +x_0=2000 +y_0=7000 +ellps=WGS84 +units=m +no_defs',
       '+proj=lcc +lat_1=66.3 +lon_0=-34 +k_0=0.9999 +ellps=GRS80 +units=m +no_defs',
       '+proj=lcc +lat_1=-20 +lat_2=-50 +lat_0=-35 +lon_0=140 +a=6378137 +es=0.00669438 +units=m +no_defs']


def _mp_setup(proj4):
    import mpmath as mp
    from oracle.proj_conformal import make
    mp.mp.dps = 40
    P = make(proj4)
    a, e = mp.mpf(P.a), mp.sqrt(mp.mpf(P.es))
    t = lambda phi: mp.tan(mp.pi / 4 - phi / 2) / ((1 - e * mp.sin(phi)) / (1 + e * mp.sin(phi))) ** (e / 2)      # noqa: E731  (15-9)
    m = lambda phi: mp.cos(phi) / mp.sqrt(1 - e * e * mp.sin(phi) ** 2)                                            # noqa: E731  (14-15)
    return mp, P, a, e, t, m


@pytest.mark.parametrize('proj4', MERC)
def test_mercator_oracle_against_mpmath(proj4):
    mp, P, a, e, t, m = _mp_setup(proj4)
    p = P.p
    k0 = m(mp.radians(abs(mp.mpf(p['lat_ts'])))) if 'lat_ts' in p else mp.mpf(p.get('k_0', 1.0))
    rng = np.random.default_rng(5)
    lon = P.lon_0 + rng.uniform(-60, 60, 60)
    lat = rng.uniform(-80, 80, 60)
    x, y = P.forward(lon, lat)
    for i in range(len(lon)):
        phi, dl = mp.radians(mp.mpf(float(lat[i]))), mp.radians(mp.mpf(float(lon[i])) - mp.mpf(P.lon_0))
        ex = a * k0 * dl + mp.mpf(P.x_0)
        ey = -a * k0 * mp.log(t(phi)) + mp.mpf(P.y_0)                         # 7-7 in the t form: y = -a k0 ln t
        assert abs(float(ex) - x[i]) < 3e-8 and abs(float(ey) - y[i]) < 3e-8, (i, float(ex), x[i], float(ey), y[i])
    lo, la = P.inverse(x, y)
    assert np.max(np.abs((lo - lon + 180) % 360 - 180)) < 1e-12 and np.max(np.abs(la - lat)) < 1e-12


@pytest.mark.parametrize('proj4', LCC)
def test_lambert_conformal_conic_oracle_against_mpmath(proj4):
    mp, P, a, e, t, m = _mp_setup(proj4)
    phi1, phi2, phi0 = (mp.radians(mp.mpf(v)) for v in (P.lat_1, P.lat_2, P.lat_0))
    if abs(P.lat_1 - P.lat_2) < 1e-12:
        n = mp.sin(phi1)                                                      # 15-8 with one standard parallel
    else:
        n = (mp.log(m(phi1)) - mp.log(m(phi2))) / (mp.log(t(phi1)) - mp.log(t(phi2)))
    F = m(phi1) / (n * t(phi1) ** n)                                           # 15-10
    rho0 = a * F * t(phi0) ** n
    k0 = mp.mpf(P.k_0)
    rng = np.random.default_rng(6)
    lon = P.lon_0 + rng.uniform(-50, 50, 60)
    lat = np.clip(P.lat_0 + rng.uniform(-30, 30, 60), -85, 85)
    x, y = P.forward(lon, lat)
    for i in range(len(lon)):
        phi, dl = mp.radians(mp.mpf(float(lat[i]))), mp.radians(mp.mpf(float(lon[i])) - mp.mpf(P.lon_0))
        rho = a * F * t(phi) ** n                                              # 15-7
        ex = k0 * rho * mp.sin(n * dl) + mp.mpf(P.x_0)
        ey = k0 * (rho0 - rho * mp.cos(n * dl)) + mp.mpf(P.y_0)
        assert abs(float(ex) - x[i]) < 3e-8 and abs(float(ey) - y[i]) < 3e-8, (i, float(ex), x[i], float(ey), y[i])
    lo, la = P.inverse(x, y)
    assert np.max(np.abs((lo - lon + 180) % 360 - 180)) < 1e-12 and np.max(np.abs(la - lat)) < 1e-12


STERE_E = ['+proj=stere +lat_0=90 +lon_0=-45 +lat_ts=70 +ellps=WGS84 +units=m +no_defs',
           '+proj=stere +lat_0=90 +lon_0=10 +k_0=0.994 +x_0=2000000 +y_0=2000000 +ellps=WGS84 +units=m +no_defs',
           '+proj=stere +lat_0=-90 +lon_0=0 +lat_ts=-71 +ellps=WGS84 +units=m +no_defs',
           '+proj=stere +lat_0=52.2 +lon_0=5.4 +k_0=0.9999079 +x_0=155000 +y_0=463000 +a=6377397.155 +rf=299.1528128 +units=m +no_defs',
           '+proj=stere +lat_0=0 +lon_0=20 +ellps=GRS80 +units=m +no_defs']


@pytest.mark.parametrize('proj4', STERE_E)
def test_ellipsoidal_stereographic_oracle_against_mpmath(proj4):
    """Snyder 21-24 .. 21-40 at 40 digits: conformal latitude chi (3-1), oblique A = 2 a k0 m1 / (cos chi1 (1 + sin chi1 sin chi +
    cos chi1 cos chi cos dlam)); polar rho = 2 a k0 t / sqrt((1+e)^(1+e) (1-e)^(1-e)) or a m_c t / t_c with a latitude of true scale."""
    mp, P, a, e, t, m = _mp_setup(proj4)
    chi = lambda phi: 2 * mp.atan(mp.tan(mp.pi / 4 + phi / 2) * ((1 - e * mp.sin(phi)) / (1 + e * mp.sin(phi))) ** (e / 2)) - mp.pi / 2   # noqa: E731
    phi1 = mp.radians(mp.mpf(P.lat_0))
    k0 = mp.mpf(P.k_0)
    rng = np.random.default_rng(8)
    lon = P.lon_0 + rng.uniform(-70, 70, 60)
    if abs(P.lat_0) == 90:
        lat = np.sign(P.lat_0) * rng.uniform(35, 89.5, 60)
    else:
        lat = np.clip(P.lat_0 + rng.uniform(-40, 40, 60), -85, 85)
    x, y = P.forward(lon, lat)
    for i in range(len(lon)):
        phi, dl = mp.radians(mp.mpf(float(lat[i]))), mp.radians(mp.mpf(float(lon[i])) - mp.mpf(P.lon_0))
        if abs(P.lat_0) == 90:
            sgn = 1 if P.lat_0 > 0 else -1
            tt = t(sgn * phi)
            if 'lat_ts' in P.p:
                pc = mp.radians(abs(mp.mpf(P.p['lat_ts'])))
                rho = a * m(pc) * tt / t(pc)                                   # 21-34
            else:
                rho = 2 * a * k0 * tt / mp.sqrt((1 + e) ** (1 + e) * (1 - e) ** (1 - e))          # 21-33
            ex, ey = rho * mp.sin(dl), -sgn * rho * mp.cos(dl)                  # 21-30, 21-31 (south: signs of phi, lam, x, y reversed)
        else:
            c1, c = chi(phi1), chi(phi)
            A = 2 * a * k0 * m(phi1) / (mp.cos(c1) * (1 + mp.sin(c1) * mp.sin(c) + mp.cos(c1) * mp.cos(c) * mp.cos(dl)))   # 21-27
            ex = A * mp.cos(c) * mp.sin(dl)
            ey = A * (mp.cos(c1) * mp.sin(c) - mp.sin(c1) * mp.cos(c) * mp.cos(dl))
        ex, ey = ex + mp.mpf(P.x_0), ey + mp.mpf(P.y_0)
        assert abs(float(ex) - x[i]) < 5e-8 and abs(float(ey) - y[i]) < 5e-8, (i, float(ex), x[i], float(ey), y[i])
    lo, la = P.inverse(x, y)
    assert np.max(np.abs((lo - lon + 180) % 360 - 180)) < 1e-11 and np.max(np.abs(la - lat)) < 1e-12
