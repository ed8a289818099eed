"""ORACLE (test infrastructure, never the product path).

CPU restatement, in NumPy float64, of three more map projections a reader's grid may lie on -- Mercator ('+proj=merc') and
Lambert conformal conic ('+proj=lcc', one or two standard parallels) on a sphere or an ellipsoid, and the stereographic
projection of the ELLIPSOID ('+proj=stere' with an ellipsoid; the sphere is oracle/proj_stere.py) -- for the fake ``pyproj`` of
oracle/refrun.py: the reference hands every position to ``pyproj.Proj`` (BaseReader.lonlat2xy / xy2lonlat,
readers/basereader/variables.py:114-143) and rotates vector components through ``Transformer`` + ``Geod.inv``
(rotate_vectors, variables.py:59-109).

pyproj / PROJ are third-party and absent from /root/reference and from this image (pyproj>=2.3 on PROJ<9.8, pyproject.toml:18-19).
What is restated is the published algorithm -- J. P. Snyder, "Map Projections: A Working Manual" (USGS PP 1395, 1987): Mercator
eqs. 7-1, 7-2 (sphere), 7-6, 7-7 (ellipsoid; y written with asinh / atanh), inverse by the iteration 7-9; Lambert conformal conic
eqs. 15-1 .. 15-5 (sphere), 15-7 .. 15-11 with 14-15 and 15-9 (ellipsoid), inverse 15-5 / 7-9 -- organised the way PROJ's merc.cpp /
lcc.cpp organise them (k_0 from +lat_ts for Mercator; n, c (= F), rho_0 for the cone; lam = lon - lon_0 reduced to [-pi, pi];
x = a k_0 x' + x_0), with PROJ's helper functions msfn (m of 14-15), tsfn (t of 15-9) and phi2 (7-9).

Pinning: PARITY UNPINNED against a PROJ binary.  tests/test_oracle_proj_conformal.py checks it against the closed forms (Snyder's
equations written independently with mpmath at 40 digits: 7-7, 15-7 .. 15-10, 21-27, 21-33 / 21-34), and round trips.
"""
import numpy as np

from oracle.proj_stere import parse_proj4

_EPS10 = 1e-10
_DEG = np.pi / 180.0
_HALFPI = 0.5 * np.pi
_FORTPI = 0.25 * np.pi

ELLIPSOIDS = {'WGS84': (6378137.0, 298.257223563), 'GRS80': (6378137.0, 298.257222101), 'sphere': (6370997.0, 0.0)}


def ellipsoid(p, proj4):
    """(a, es) from the +R / +a / +b / +rf / +f / +e / +es / +ellps / +datum parameters; an ellipsoid must be named."""
    if 'R' in p:
        return float(p['R']), 0.0
    a = rf = None
    if 'ellps' in p or p.get('datum') == 'WGS84':
        name = p.get('ellps', 'WGS84')
        if name not in ELLIPSOIDS:
            raise NotImplementedError('ellipsoid %s: %s' % (name, proj4))
        a, rf = ELLIPSOIDS[name]
    if 'a' in p:
        a = float(p['a'])
    if a is None:
        raise NotImplementedError('no ellipsoid (+R, +a, +ellps, +datum=WGS84) in %s' % proj4)
    if 'es' in p:
        es = float(p['es'])
    elif 'e' in p:
        es = float(p['e']) ** 2
    elif 'rf' in p:
        f = 1.0 / float(p['rf'])
        es = f * (2.0 - f)
    elif 'f' in p:
        f = float(p['f'])
        es = f * (2.0 - f)
    elif 'b' in p:
        es = 1.0 - (float(p['b']) / a) ** 2
    elif rf:
        f = 1.0 / rf
        es = f * (2.0 - f)
    else:
        es = 0.0
    return a, es


def msfn(sinphi, cosphi, es):
    return cosphi / np.sqrt(1.0 - es * sinphi * sinphi)


def tsfn(phi, sinphi, e):
    return np.tan(0.5 * (_HALFPI - phi)) / np.power((1.0 - e * sinphi) / (1.0 + e * sinphi), 0.5 * e)


def phi2(ts, e):
    """Snyder 7-9: the latitude whose isometric co-latitude function t is ts (iterated to 1e-14)."""
    ts = np.asarray(ts, dtype=np.float64)
    phi = _HALFPI - 2.0 * np.arctan(ts)
    for _ in range(20):
        con = e * np.sin(phi)
        new = _HALFPI - 2.0 * np.arctan(ts * np.power((1.0 - con) / (1.0 + con), 0.5 * e))
        done = np.max(np.abs(new - phi)) if new.size else 0.0
        phi = new
        if done < 1e-14:
            break
    return phi


def adjlon(lam):
    lam = np.asarray(lam, dtype=np.float64)
    return np.where(np.abs(lam) > np.pi, np.mod(lam + np.pi, 2 * np.pi) - np.pi, lam)


class _Base:
    def __init__(self, proj4, name):
        p = parse_proj4(proj4)
        if p.get('proj') != name:
            raise NotImplementedError(proj4)
        if p.get('units', 'm') != 'm':
            raise NotImplementedError('projection units other than metres: %s' % proj4)
        self.proj4 = str(proj4)
        self.p = p
        self.a, self.es = ellipsoid(p, proj4)
        self.e = float(np.sqrt(self.es))
        self.lon_0 = float(p.get('lon_0', 0.0))
        self.lam0 = self.lon_0 * _DEG
        self.k_0 = float(p.get('k_0', p.get('k', 1.0)))
        self.x_0, self.y_0 = float(p.get('x_0', 0.0)), float(p.get('y_0', 0.0))

    def forward(self, lon, lat):
        lon = np.atleast_1d(np.asarray(lon, dtype=np.float64))
        lat = np.atleast_1d(np.asarray(lat, dtype=np.float64))
        with np.errstate(all='ignore'):
            x, y = self._fwd(adjlon(lon * _DEG - self.lam0), lat * _DEG)
        return self.a * x + self.x_0, self.a * y + self.y_0

    def inverse(self, x, y):
        x = (np.atleast_1d(np.asarray(x, dtype=np.float64)) - self.x_0) / self.a
        y = (np.atleast_1d(np.asarray(y, dtype=np.float64)) - self.y_0) / self.a
        with np.errstate(all='ignore'):
            lam, phi = self._inv(x, y)
        return adjlon(lam + self.lam0) / _DEG, phi / _DEG


class Merc(_Base):
    """Mercator (forward: degrees -> metres, inverse: metres -> degrees)."""

    def __init__(self, proj4):
        super().__init__(proj4, 'merc')
        if 'lat_ts' in self.p:
            phits = abs(float(self.p['lat_ts'])) * _DEG
            if phits >= _HALFPI:
                raise ValueError('lat_ts >= 90')
            self.k_0 = float(msfn(np.sin(phits), np.cos(phits), self.es))      # (= cos(lat_ts) on a sphere)
        self.desc_extra = {}

    def _fwd(self, lam, phi):
        y = np.arcsinh(np.tan(phi)) - self.e * np.arctanh(self.e * np.sin(phi))
        return self.k_0 * lam, self.k_0 * y

    def _inv(self, x, y):
        ts = np.exp(-y / self.k_0)
        phi = phi2(ts, self.e) if self.es != 0.0 else _HALFPI - 2.0 * np.arctan(ts)
        return x / self.k_0, phi


class Lcc(_Base):
    """Lambert conformal conic, one (+lat_1) or two (+lat_1 +lat_2) standard parallels."""

    def __init__(self, proj4):
        super().__init__(proj4, 'lcc')
        p = self.p
        if 'lat_1' not in p:
            raise NotImplementedError('+proj=lcc needs +lat_1: %s' % proj4)
        self.lat_1 = float(p['lat_1'])
        self.lat_2 = float(p.get('lat_2', self.lat_1))
        self.lat_0 = float(p.get('lat_0', self.lat_1 if 'lat_2' not in p else 0.0))
        phi1, phi2_, phi0 = self.lat_1 * _DEG, self.lat_2 * _DEG, self.lat_0 * _DEG
        if abs(phi1 + phi2_) < _EPS10:
            raise ValueError('lat_1 = -lat_2')
        sinphi, cosphi = np.sin(phi1), np.cos(phi1)
        n = sinphi
        secant = abs(phi1 - phi2_) >= _EPS10
        e, es = self.e, self.es
        if es != 0.0:
            m1 = msfn(sinphi, cosphi, es)
            ml1 = tsfn(phi1, sinphi, e)
            if secant:
                s2 = np.sin(phi2_)
                n = np.log(m1 / msfn(s2, np.cos(phi2_), es)) / np.log(ml1 / tsfn(phi2_, s2, e))
            c = m1 * np.power(ml1, -n) / n
            rho0 = 0.0 if abs(abs(phi0) - _HALFPI) < _EPS10 else c * np.power(tsfn(phi0, np.sin(phi0), e), n)
        else:
            if secant:
                n = np.log(cosphi / np.cos(phi2_)) / np.log(np.tan(_FORTPI + 0.5 * phi2_) / np.tan(_FORTPI + 0.5 * phi1))
            c = cosphi * np.power(np.tan(_FORTPI + 0.5 * phi1), n) / n
            rho0 = 0.0 if abs(abs(phi0) - _HALFPI) < _EPS10 else c * np.power(np.tan(_FORTPI + 0.5 * phi0), -n)
        self.n, self.c, self.rho0 = float(n), float(c), float(rho0)

    def _fwd(self, lam, phi):
        if self.es != 0.0:
            rho = self.c * np.power(tsfn(phi, np.sin(phi), self.e), self.n)
        else:
            rho = self.c * np.power(np.tan(_FORTPI + 0.5 * phi), -self.n)
        rho = np.where(np.abs(np.abs(phi) - _HALFPI) < _EPS10, np.where(phi * self.n > 0, 0.0, np.nan), rho)
        lam = lam * self.n
        return self.k_0 * rho * np.sin(lam), self.k_0 * (self.rho0 - rho * np.cos(lam))

    def _inv(self, x, y):
        x = x / self.k_0
        y = self.rho0 - y / self.k_0
        rho = np.hypot(x, y)
        if self.n < 0:
            rho, x, y = -rho, -x, -y
        safe = np.where(rho != 0, rho, 1.0)
        if self.es != 0.0:
            phi = phi2(np.power(safe / self.c, 1.0 / self.n), self.e)
        else:
            phi = 2.0 * np.arctan(np.power(self.c / safe, 1.0 / self.n)) - _HALFPI
        lam = np.arctan2(x, y) / self.n
        pole = _HALFPI if self.n > 0 else -_HALFPI
        return np.where(rho != 0, lam, 0.0), np.where(rho != 0, phi, pole)


class StereEllipsoid(_Base):
    """Stereographic projection of the ELLIPSOID (Snyder eqs. 21-24 .. 21-40, inverse 21-15, 21-36 .. 21-38, 20-14 .. 20-16 and the
    iteration 3-4 / 7-9), organised like the ellipsoidal half of PROJ's stere.cpp: conformal latitude chi through
    ssfn(phi) = tan(pi/4 + phi/2) ((1 - e sin phi) / (1 + e sin phi))^(e/2); oblique / equatorial aspects with
    akm1 = 2 k_0 cos(phi_0) / sqrt(1 - e^2 sin^2 phi_0); polar aspects with akm1 = 2 k_0 / sqrt((1+e)^(1+e) (1-e)^(1-e)), or, with a
    latitude of true scale, akm1 = m(lat_ts) / t(lat_ts)."""

    def __init__(self, proj4):
        super().__init__(proj4, 'stere')
        p = self.p
        if self.es == 0.0:
            raise NotImplementedError('sphere: oracle/proj_stere.py')
        self.lat_0 = float(p.get('lat_0', 0.0))
        self.has_lat_ts = 'lat_ts' in p
        phits = abs(float(p['lat_ts'])) * _DEG if self.has_lat_ts else _HALFPI
        phi0 = self.lat_0 * _DEG
        self.phi0 = phi0
        t = abs(phi0)
        if abs(t - _HALFPI) < _EPS10:
            self.mode = 'S_POLE' if phi0 < 0 else 'N_POLE'
        else:
            self.mode = 'OBLIQ' if t > _EPS10 else 'EQUIT'
        e = self.e
        if self.mode in ('N_POLE', 'S_POLE'):
            if abs(phits - _HALFPI) < _EPS10:
                self.akm1 = 2.0 * self.k_0 / np.sqrt(np.power(1 + e, 1 + e) * np.power(1 - e, 1 - e))
            else:
                sp = np.sin(phits)
                self.akm1 = float(msfn(sp, np.cos(phits), self.es) / tsfn(phits, sp, e))
            self.sinX1 = self.cosX1 = 0.0
        else:
            sp = np.sin(phi0)
            X = 2.0 * np.arctan(self._ssfn(phi0)) - _HALFPI
            self.akm1 = 2.0 * self.k_0 * np.cos(phi0) / np.sqrt(1.0 - self.es * sp * sp)
            self.sinX1, self.cosX1 = float(np.sin(X)), float(np.cos(X))

    def _ssfn(self, phi):
        s = self.e * np.sin(phi)
        return np.tan(0.5 * (_HALFPI + phi)) * np.power((1.0 - s) / (1.0 + s), 0.5 * self.e)

    def _fwd(self, lam, phi):
        sl, cl = np.sin(lam), np.cos(lam)
        if self.mode in ('OBLIQ', 'EQUIT'):
            X = 2.0 * np.arctan(self._ssfn(phi)) - _HALFPI
            sX, cX = np.sin(X), np.cos(X)
            if self.mode == 'OBLIQ':
                A = self.akm1 / (self.cosX1 * (1.0 + self.sinX1 * sX + self.cosX1 * cX * cl))
                return A * cX * sl, A * (self.cosX1 * sX - self.sinX1 * cX * cl)
            A = self.akm1 / (1.0 + cX * cl)
            return A * cX * sl, A * sX
        if self.mode == 'S_POLE':
            phi, cl = -phi, -cl
        rho = self.akm1 * tsfn(phi, np.sin(phi), self.e)
        return rho * sl, -rho * cl

    def _inv(self, x, y):
        rho = np.hypot(x, y)
        e = self.e
        if self.mode in ('OBLIQ', 'EQUIT'):
            tp = 2.0 * np.arctan2(rho * self.cosX1, self.akm1)
            ct, st = np.cos(tp), np.sin(tp)
            safe = np.where(rho != 0, rho, 1.0)
            phi_l = np.where(rho == 0.0, np.arcsin(ct * self.sinX1), np.arcsin(ct * self.sinX1 + y * st * self.cosX1 / safe))
            tp = np.tan(0.5 * (_HALFPI + phi_l))
            xx = x * st
            yy = rho * self.cosX1 * ct - y * self.sinX1 * st
            halfpi, halfe = _HALFPI, 0.5 * e
        else:
            if self.mode == 'N_POLE':
                y = -y
            tp = -rho / self.akm1
            phi_l = _HALFPI - 2.0 * np.arctan(-tp)
            xx, yy = x, y
            halfpi, halfe = -_HALFPI, -0.5 * e
        phi = phi_l
        for _ in range(20):
            sp = e * np.sin(phi_l)
            phi = 2.0 * np.arctan(tp * np.power((1.0 + sp) / (1.0 - sp), halfe)) - halfpi
            d = np.max(np.abs(phi_l - phi)) if np.size(phi) else 0.0
            phi_l = phi
            if d < 1e-14:
                break
        if self.mode == 'S_POLE':
            phi = -phi
        lam = np.where((xx == 0.0) & (yy == 0.0), 0.0, np.arctan2(xx, yy))
        return lam, phi


def make(proj4):
    p = parse_proj4(str(proj4))
    if p.get('proj') == 'merc':
        return Merc(proj4)
    if p.get('proj') == 'lcc':
        return Lcc(proj4)
    if p.get('proj') == 'stere':
        return StereEllipsoid(proj4)
    raise NotImplementedError(proj4)
