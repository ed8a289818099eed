"""Gridded readers on a Mercator or Lambert-conformal-conic plane (sphere and ellipsoid): the drop-in model classes against runs of
the UNMODIFIED reference on readers of the same projection (tests/golden/proj2_ref.npz, written by `python tests/projcases.py` in
the build container; the reference's pyproj is the stand-in of oracle/refrun.py backed by oracle/proj_conformal.py, which
tests/test_oracle_proj_conformal.py pins to the closed forms at 40 digits).  Same construction as the stereographic cases of
tests/bookkeeping.py (PROJ_CASES)."""
import os

import numpy as np

import common
import bookkeeping as bk

GOLDEN = os.path.join(common.GOLDEN, 'proj2_ref.npz')
MERC_WGS84 = '+proj=merc +lon_0=0 +lat_ts=60 +ellps=WGS84 +units=m +no_defs'
MERC_SPHERE = '+proj=merc +lon_0=10 +k_0=0.9 +x_0=100000 +y_0=-50000 +R=6371000 +units=m +no_defs'
LCC_SPHERE = '+proj=lcc +lat_0=63.3 +lon_0=15 +lat_1=63.3 +lat_2=63.3 +R=6371000 +units=m +no_defs'           # (the MetCoOp / MEPS grid)
LCC_WGS84 = '+proj=lcc +lat_1=52 +lat_2=68 +lat_0=60 +lon_0=8 +x_0=400000 +y_0=200000 +ellps=WGS84 +units=m +no_defs'
STERE_WGS84_POLAR = '+proj=stere +lat_0=90 +lon_0=-10 +lat_ts=70 +x_0=1000000 +y_0=2500000 +ellps=WGS84 +units=m +no_defs'
STERE_GRS80_OBLIQUE = '+proj=stere +lat_0=61 +lon_0=4.5 +k_0=0.9999 +ellps=GRS80 +units=m +no_defs'
CASES = {
    'stere_wgs84_polar_rk4_3d_w': dict(proj4=STERE_WGS84_POLAR, model='OceanDrift', readers=('cur3d',), steps=8, dt=600,
                                       cfg={'drift:advection_scheme': 'runge-kutta4'}),
    'stere_grs80_oblique_leeway': dict(proj4=STERE_GRS80_OBLIQUE, model='Leeway', readers=('cur2d', 'wind'), steps=6, dt=600, cfg={}),
    'merc_wgs84_rk4_3d_w': dict(proj4=MERC_WGS84, model='OceanDrift', readers=('cur3d',), steps=8, dt=600,
                                cfg={'drift:advection_scheme': 'runge-kutta4'}),
    'lcc_sphere_rk2_wind': dict(proj4=LCC_SPHERE, model='OceanDrift', readers=('cur2d', 'wind'), steps=6, dt=900,
                                cfg={'drift:advection_scheme': 'runge-kutta', 'drift:vertical_advection': False}, seed={'z': 0.0}),
    'merc_sphere_mixing': dict(proj4=MERC_SPHERE, model='OceanDrift', readers=('cur3d_k',), steps=3, dt=600,
                               cfg={'drift:vertical_mixing': True, 'drift:vertical_advection': False, 'vertical_mixing:timestep': 60.0}),
    'lcc_wgs84_leeway': dict(proj4=LCC_WGS84, model='Leeway', readers=('cur2d', 'wind'), steps=6, dt=600, cfg={}),
    'lcc_wgs84_rk4_3d_w': dict(proj4=LCC_WGS84, model='OceanDrift', readers=('cur3d',), steps=8, dt=600,
                               cfg={'drift:advection_scheme': 'runge-kutta4'}),
}


_bk_proj_setup = bk.proj_setup


def setup(case):
    """bookkeeping.proj_setup with the projection of this module's case."""
    from oracle.proj_conformal import make
    saved = bk.PROJ_CASES.get(case)
    bk.PROJ_CASES[case] = CASES[case]
    import oracle.proj_stere as ps
    orig = ps.Stere
    ps.Stere = make                                  # (proj_setup builds its projection through this name)
    try:
        return _bk_proj_setup(case)
    finally:
        ps.Stere = orig
        if saved is None:
            del bk.PROJ_CASES[case]


def run_case(case, classes, make_reader, **model_kw):
    saved_setup = bk.proj_setup
    bk.proj_setup = setup
    try:
        return bk.run_proj_case(case, classes, make_reader, **model_kw)
    finally:
        bk.proj_setup = saved_setup


def run_product(case, **model_kw):
    from opendrift_b200.models.oceandrift import OceanDrift
    from opendrift_b200.models.leeway import Leeway
    from opendrift_b200.readers import reader_regular_grid
    return run_case(case, {'OceanDrift': OceanDrift, 'Leeway': Leeway},
                    lambda x, y, z, t, f, name, proj4: reader_regular_grid.Reader(x, y, z, t, f, name=name, proj4=proj4), **model_kw)


def check(o, case):
    ref = np.load(GOLDEN)
    g = lambda k: ref['%s__%s' % (case, k)]                 # noqa: E731
    e = max(common.max_err_deg(np.asarray(o.elements.lon), np.asarray(o.elements.lat), g('lon'), g('lat')))
    dz = float(np.abs(np.asarray(o.elements.z, dtype=np.float64) - g('z')).max())
    moved = float(np.abs(g('lon') - g('lon0')).max())
    return e, dz, moved


if __name__ == '__main__':
    from oracle import refrun
    refrun.setup()
    from opendrift.models.oceandrift import OceanDrift as RefOD
    from opendrift.models.leeway import Leeway as RefLW
    out = {}
    for case in CASES:
        ro = run_case(case, {'OceanDrift': RefOD, 'Leeway': RefLW},
                      lambda x, y, z, t, f, name, proj4: refrun.make_grid_reader(x, y, z, t, f, name=name, proj4=proj4), logfile='/tmp/od_proj2.log')
        lon0 = setup(case)[6]
        out.update({'%s__lon' % case: np.asarray(ro.elements.lon, dtype=np.float64), '%s__lat' % case: np.asarray(ro.elements.lat, dtype=np.float64),
                    '%s__z' % case: np.asarray(ro.elements.z, dtype=np.float64), '%s__lon0' % case: lon0.astype(np.float64)})
        print(case, len(ro.elements.lon), 'moved', float(np.abs(np.asarray(ro.elements.lon) - lon0).max()))
    np.savez_compressed(GOLDEN, **out)
    print('wrote', GOLDEN)
