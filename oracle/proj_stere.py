"""ORACLE (test infrastructure, never the product path).

CPU restatement, in NumPy float64, of the one map projection the reference's own CPU-runnable case needs
(BASELINE.json configs[0]: examples/example_double_gyre_advection_schemes.py): the SPHERICAL stereographic projection
that opendrift/readers/reader_double_gyre.py:30-31 asks pyproj for,

    '+proj=stere +lat_0=0 +lon_0=0 +lat_ts=0 +units=m +a=6.371e+06 +e=0 +no_defs'

Reference call sites: BaseReader.lonlat2xy / xy2lonlat (readers/basereader/variables.py:114-143: ``self.proj(lon, lat,
inverse=False|True)``) on every get_variables_interpolated call, and ``pyproj.Transformer.from_proj(proj_from,
proj_to).transform`` inside rotate_vectors (variables.py:59-109).

pyproj / PROJ are third-party and absent from /root/reference and from this image (pyproj>=2.3 on PROJ<9.8,
pyproject.toml:18-19).  What is restated is the published algorithm -- J. P. Snyder, "Map Projections: A Working
Manual" (USGS PP 1395, 1987), eqs. 21-2 .. 21-4 (forward) and 20-14, 20-15, 20-18, 21-15 (inverse) -- organised the
way PROJ's stere.cpp organises the spherical case (four aspects: equatorial, oblique, north / south polar; scale
constant akm1 = 2 k0, or cos(lat_ts) / tan(pi/4 - lat_ts/2) for the polar aspects), wrapped in PROJ's generic steps
(lam = lon - lon_0 reduced to [-pi, pi]; x = a * x' + x_0).

Pinning: PARITY UNPINNED against a PROJ binary.  tests/test_oracle_proj.py checks it against the closed-form
projection evaluated with mpmath at 40 digits, round trips, and the reference's own numbers for this reader
(the example's seed point x = 0.6 m, y = 0.3 m).
"""
import re

import numpy as np

_EPS10 = 1e-10
_DEG = np.pi / 180.0            # PROJ DEG_TO_RAD
_RAD2DEG = 180.0 / np.pi        # PROJ RAD_TO_DEG
_HALFPI = 0.5 * np.pi
_FORTPI = 0.25 * np.pi


def parse_proj4(s):
    """{'proj': 'stere', 'lat_0': 0.0, ...} from a '+key=value +flag' string (values as float where they parse)."""
    out = {}
    for m in re.finditer(r'\+([A-Za-z_0-9]+)(?:=(\S+))?', s):
        k, v = m.group(1), m.group(2)
        if v is None:
            out[k] = True
        else:
            try:
                out[k] = float(v)
            except ValueError:
                out[k] = v
    return out


class Stere:
    """Spherical stereographic projection (forward: degrees -> metres, inverse: metres -> degrees)."""

    def __init__(self, proj4):
        p = parse_proj4(proj4)
        if p.get('proj') != 'stere':
            raise NotImplementedError(proj4)
        if 'R' in p:
            self.a = float(p['R'])
        else:
            self.a = float(p.get('a', 6378137.0))
            es = float(p.get('e', 0.0)) ** 2 if 'e' in p else float(p.get('es', 0.0))
            if 'b' in p and float(p['b']) != self.a or 'rf' in p or 'f' in p or 'ellps' in p:
                raise NotImplementedError('ellipsoidal stereographic: ' + proj4)
            if es != 0.0:
                raise NotImplementedError('ellipsoidal stereographic: ' + proj4)
        if p.get('units', 'm') != 'm':
            raise NotImplementedError(proj4)
        self.proj4 = proj4
        self.phi0 = float(p.get('lat_0', 0.0)) * _DEG
        self.lam0 = float(p.get('lon_0', 0.0)) * _DEG
        self.k0 = float(p.get('k_0', p.get('k', 1.0)))
        self.x0 = float(p.get('x_0', 0.0))
        self.y0 = float(p.get('y_0', 0.0))
        phits = float(p['lat_ts']) * _DEG if 'lat_ts' in p else _HALFPI
        t = abs(self.phi0)
        if abs(t - _HALFPI) < _EPS10:
            self.mode = 'S_POLE' if self.phi0 < 0 else 'N_POLE'
        else:
            self.mode = 'OBLIQ' if t > _EPS10 else 'EQUIT'
        phits = abs(phits)
        self.sinX1 = np.sin(self.phi0)
        self.cosX1 = np.cos(self.phi0)
        if self.mode in ('OBLIQ', 'EQUIT'):
            self.akm1 = 2.0 * self.k0
        else:
            self.akm1 = (np.cos(phits) / np.tan(_FORTPI - 0.5 * phits)) if abs(phits - _HALFPI) >= _EPS10 else 2.0 * self.k0

    @staticmethod
    def _adjlon(lam):
        """Reduce to [-pi, pi] (PROJ adjlon: values already inside are left untouched)."""
        lam = np.asarray(lam, dtype=np.float64)
        out = lam.copy()
        big = np.abs(lam) > np.pi
        if np.any(big):
            t = lam[big] + np.pi
            t = t - 2.0 * np.pi * np.floor(t / (2.0 * np.pi))
            out[big] = t - np.pi
        return out

    def forward(self, lon, lat):
        lon = np.asarray(lon, dtype=np.float64)
        lat = np.asarray(lat, dtype=np.float64)
        lam = self._adjlon(lon * _DEG - self.lam0)
        phi = lat * _DEG
        sinphi, cosphi = np.sin(phi), np.cos(phi)
        sinlam, coslam = np.sin(lam), np.cos(lam)
        with np.errstate(all='ignore'):
            if self.mode in ('EQUIT', 'OBLIQ'):
                if self.mode == 'EQUIT':
                    d = 1.0 + cosphi * coslam
                else:
                    d = 1.0 + self.sinX1 * sinphi + self.cosX1 * cosphi * coslam
                k = self.akm1 / d
                x = k * cosphi * sinlam
                if self.mode == 'EQUIT':
                    y = k * sinphi
                else:
                    y = k * (self.cosX1 * sinphi - self.sinX1 * cosphi * coslam)
                bad = d <= _EPS10
            else:
                if self.mode == 'N_POLE':
                    coslam = -coslam
                    phi = -phi
                bad = np.abs(phi - _HALFPI) < 1e-8
                y = self.akm1 * np.tan(_FORTPI + 0.5 * phi)
                x = sinlam * y
                y = y * coslam
        x = self.a * x + self.x0
        y = self.a * y + self.y0
        x = np.where(bad, np.inf, x)
        y = np.where(bad, np.inf, y)
        return x, y

    def inverse(self, x, y):
        ra = 1.0 / self.a
        x = (np.asarray(x, dtype=np.float64) - self.x0) * ra
        y = (np.asarray(y, dtype=np.float64) - self.y0) * ra
        rh = np.hypot(x, y)
        c = 2.0 * np.arctan(rh / self.akm1)
        sinc, cosc = np.sin(c), np.cos(c)
        small = np.abs(rh) <= _EPS10
        rhs = np.where(small, 1.0, rh)
        with np.errstate(all='ignore'):
            if self.mode == 'EQUIT':
                phi = np.where(small, 0.0, np.arcsin(np.clip(y * sinc / rhs, -1.0, 1.0)))
                lam = np.where((cosc != 0.0) | (x != 0.0), np.arctan2(x * sinc, cosc * rh), 0.0)
            elif self.mode == 'OBLIQ':
                phi = np.where(small, self.phi0, np.arcsin(np.clip(cosc * self.sinX1 + y * sinc * self.cosX1 / rhs, -1.0, 1.0)))
                cc = cosc - self.sinX1 * np.sin(phi)
                lam = np.where((cc != 0.0) | (x != 0.0), np.arctan2(x * sinc * self.cosX1, cc * rh), 0.0)
            else:
                if self.mode == 'N_POLE':
                    y = -y
                phi = np.where(small, self.phi0, np.arcsin(-cosc if self.mode == 'S_POLE' else cosc))
                lam = np.where((x == 0.0) & (y == 0.0), 0.0, np.arctan2(x, y))
        lam = self._adjlon(lam + self.lam0)
        return lam * _RAD2DEG, phi * _RAD2DEG
