"""Host-side map projection of a reader's native plane, for the handful of points user scripts convert
(`reader.xy2lonlat(x, y)` to place seeds, `reader.lonlat2xy(lon, lat)` to read results back: BaseReader.xy2lonlat /
lonlat2xy, opendrift/readers/basereader/variables.py:114-143, which wrap pyproj.Proj).  The per-particle projection of
the hot path runs on the device (csrc/od_analytic.cuh: stere_forward / stere_inverse); this module only parses the
proj4 string into the library's od_proj_desc and mirrors the same closed forms in NumPy for those few host points.

Supported: '+proj=stere' on a sphere ('+R=...' or '+a=... +e=0'), all four aspects (Snyder 1987, ch. 21).
"""
import re

import numpy as np

from .. import _lib

_DEG = np.pi / 180.0


def parse_proj4(s):
    out = {}
    for m in re.finditer(r'\+([A-Za-z_0-9]+)(?:=(\S+))?', str(s)):
        k, v = m.group(1), m.group(2)
        if v is None:
            out[k] = True
        else:
            try:
                out[k] = float(v)
            except ValueError:
                out[k] = v
    return out


def is_geographic(proj4):
    return any(k in str(proj4) for k in ('latlong', 'longlat', 'lonlat', 'latlon'))


class SphericalStereographic:
    def __init__(self, proj4):
        p = parse_proj4(proj4)
        if p.get('proj') != 'stere':
            raise NotImplementedError('projected readers on the GPU path: spherical +proj=stere only; got %s' % proj4)
        if 'R' in p:
            a = float(p['R'])
        else:
            a = float(p.get('a', 6378137.0))
            es = float(p['e']) ** 2 if 'e' in p else float(p.get('es', 0.0))
            if es != 0.0 or any(k in p for k in ('rf', 'f', 'ellps', 'datum')) or ('b' in p and float(p['b']) != a):
                raise NotImplementedError('ellipsoidal stereographic is not on the GPU path: %s' % proj4)
        if p.get('units', 'm') != 'm':
            raise NotImplementedError('projection units other than metres: %s' % proj4)
        self.proj4 = str(proj4)
        self.a = a
        self.lat_0, self.lon_0 = float(p.get('lat_0', 0.0)), float(p.get('lon_0', 0.0))
        self.has_lat_ts = 'lat_ts' in p
        self.lat_ts = float(p.get('lat_ts', 90.0))
        self.k_0 = float(p.get('k_0', p.get('k', 1.0)))
        self.x_0, self.y_0 = float(p.get('x_0', 0.0)), float(p.get('y_0', 0.0))
        phi0 = self.lat_0 * _DEG
        t = abs(phi0)
        if abs(t - np.pi / 2) < 1e-10:
            self.mode = 'S_POLE' if phi0 < 0 else 'N_POLE'
        else:
            self.mode = 'OBLIQ' if t > 1e-10 else 'EQUIT'
        self.phi0, self.lam0 = phi0, self.lon_0 * _DEG
        self.sinX1, self.cosX1 = np.sin(phi0), np.cos(phi0)
        phits = abs(self.lat_ts * _DEG) if self.has_lat_ts else np.pi / 2
        if self.mode in ('OBLIQ', 'EQUIT') or abs(phits - np.pi / 2) < 1e-10:
            self.akm1 = 2.0 * self.k_0
        else:
            self.akm1 = np.cos(phits) / np.tan(np.pi / 4 - 0.5 * phits)

    def desc(self):
        d = _lib.ProjDesc()
        d.kind = _lib.OD_PROJ_STERE_SPHERE
        d.has_lat_ts = 1 if self.has_lat_ts else 0
        d.a, d.lat_0, d.lon_0, d.lat_ts = self.a, self.lat_0, self.lon_0, self.lat_ts
        d.k_0, d.x_0, d.y_0 = self.k_0, self.x_0, self.y_0
        return d

    @staticmethod
    def _adjlon(lam):
        lam = np.asarray(lam, dtype=np.float64)
        return np.where(np.abs(lam) > np.pi, np.mod(lam + np.pi, 2 * np.pi) - np.pi, lam)

    def __call__(self, a, b, inverse=False):
        scalar = np.isscalar(a)
        a = np.atleast_1d(np.asarray(a, dtype=np.float64))
        b = np.atleast_1d(np.asarray(b, dtype=np.float64))
        ra, rb = self._inverse(a, b) if inverse else self._forward(a, b)
        if scalar:
            return float(ra[0]), float(rb[0])
        return ra, rb

    def _forward(self, lon, lat):
        lam = self._adjlon(lon * _DEG - self.lam0)
        phi = lat * _DEG
        sp, cp, sl, cl = np.sin(phi), np.cos(phi), np.sin(lam), np.cos(lam)
        with np.errstate(all='ignore'):
            if self.mode == 'EQUIT':
                k = self.akm1 / (1.0 + cp * cl)
                x, y = k * cp * sl, k * sp
            elif self.mode == 'OBLIQ':
                k = self.akm1 / (1.0 + self.sinX1 * sp + self.cosX1 * cp * cl)
                x, y = k * cp * sl, k * (self.cosX1 * sp - self.sinX1 * cp * cl)
            else:
                if self.mode == 'N_POLE':
                    cl, phi = -cl, -phi
                y = self.akm1 * np.tan(np.pi / 4 + 0.5 * phi)
                x, y = sl * y, y * cl
        return self.a * x + self.x_0, self.a * y + self.y_0

    def _inverse(self, x, y):
        x = (x - self.x_0) / self.a
        y = (y - self.y_0) / self.a
        rh = np.hypot(x, y)
        c = 2.0 * np.arctan(rh / self.akm1)
        sc, cc = np.sin(c), np.cos(c)
        small = rh <= 1e-10
        rhs = np.where(small, 1.0, rh)
        with np.errstate(all='ignore'):
            if self.mode == 'EQUIT':
                phi = np.where(small, 0.0, np.arcsin(np.clip(y * sc / rhs, -1, 1)))
                lam = np.where((cc != 0) | (x != 0), np.arctan2(x * sc, cc * rh), 0.0)
            elif self.mode == 'OBLIQ':
                phi = np.where(small, self.phi0, np.arcsin(np.clip(cc * self.sinX1 + y * sc * self.cosX1 / rhs, -1, 1)))
                d = cc - self.sinX1 * np.sin(phi)
                lam = np.where((d != 0) | (x != 0), np.arctan2(x * sc * self.cosX1, d * rh), 0.0)
            else:
                if self.mode == 'N_POLE':
                    y = -y
                phi = np.where(small, self.phi0, np.arcsin(-cc if self.mode == 'S_POLE' else cc))
                lam = np.where((x == 0) & (y == 0), 0.0, np.arctan2(x, y))
        return self._adjlon(lam + self.lam0) / _DEG, phi / _DEG
