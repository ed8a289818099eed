"""Ensemble blocks (a reader whose get_variables() returns a list of member arrays per variable; element i of a call is served by
member i % n_members, readers/interpolation/structured.py:120-134) -- the drop-in OceanDrift against runs of the UNMODIFIED reference
on the same in-memory ensemble reader (tests/golden/ens_ref.npz, written by `python tests/enscases.py` in the build container)."""
import os

import numpy as np

import common

GOLDEN = os.path.join(common.GOLDEN, 'ens_ref.npz')
N, N_ENS = 400, 3
CASES = {
    'ens_euler_2d': ('rk4_2d', {'drift:advection_scheme': 'euler'}, 8, False),
    'ens_rk4_3d': ('rk4_3d', {'drift:advection_scheme': 'runge-kutta4'}, 8, False),
    # the ensemble reader covers only the western part: the rest comes from a second, plain reader -- the member of an element is its
    # rank among the elements the ensemble reader actually serves
    'ens_rk2_3d_partial': ('rk4_3d', {'drift:advection_scheme': 'runge-kutta'}, 6, True),
}


def members(fx, partial):
    """Three members: the fixture's current scaled and rotated a little differently per member."""
    nx = len(fx.grid_lon)
    cut = int(0.6 * nx) if partial else nx
    out_u, out_v = [], []
    for m in range(N_ENS):
        a, b = 1.0 + 0.4 * m, 0.25 * m
        out_u.append((a * fx.u - b * fx.v)[..., :cut].astype(np.float32))
        out_v.append((a * fx.v + b * fx.u)[..., :cut].astype(np.float32))
    return fx.grid_lon[:cut], out_u, out_v


def run_case(case, Model, make_ens_reader, make_reader, **model_kw):
    fxname, cfg, steps, partial = CASES[case]
    fx = common.Fixture(fxname)
    elon, eu, ev = members(fx, partial)
    o = Model(loglevel=50, **model_kw)
    o.add_reader(make_ens_reader(elon, fx.grid_lat, fx.grid_z, fx.times, {common.CUR[0]: eu, common.CUR[1]: ev}, 'ensemble'))
    if partial:
        o.add_reader(make_reader(fx.grid_lon, fx.grid_lat, fx.grid_z, fx.times, {common.CUR[0]: (0.5 * fx.u).astype(np.float32),
                                                                                 common.CUR[1]: (-0.5 * fx.v).astype(np.float32)}, 'plain'))
    for k, v in {'general:use_auto_landmask': False, 'general:coastline_action': 'none', 'drift:vertical_advection': False, **cfg}.items():
        o.set_config(k, v)
    if 'environment:constant:land_binary_mask' in getattr(o, '_config', {}):
        o.set_config('environment:constant:land_binary_mask', 0)
    z = fx.z0[:N] if fx.grid_z is not None else 0.0
    o.seed_elements(lon=fx.lon0[:N], lat=fx.lat0[:N], z=z, time=fx.start)
    o.run(steps=steps, time_step=fx.dt, time_step_output=fx.dt)
    return o


def product_ensemble_reader(lon, lat, z, times, fields, name):
    """reader_regular_grid.Reader whose fields are lists of member arrays: get_variables() hands out the list."""
    from opendrift_b200.readers import reader_regular_grid

    class EnsReader(reader_regular_grid.Reader):
        def get_variables(self, requested_variables, time=None, x=None, y=None, z=None):
            ti = self.times.index(time) if time is not None else 0
            out = {'x': self.lon, 'y': self.lat, 'time': time}
            three_d = False
            for v in requested_variables:
                out[v] = [np.array(m[ti], dtype=np.float32, copy=True) for m in self.fields[v]]
                three_d |= out[v][0].ndim == 3
            out['z'] = self.zlev if three_d else 0
            return out
    return EnsReader(lon, lat, z, times, fields, name=name)


def reference_ensemble_reader(lon, lat, z, times, fields, name):
    from oracle import refrun
    base = refrun.make_grid_reader(lon, lat, z, times, {v: m[0] for v, m in fields.items()}, name=name)

    def get_variables(requested_variables, time=None, x=None, y=None, z=None, _r=base):
        it = _r.times.index(time)
        out = {'x': _r.block_x, 'y': _r.lat, 'time': time}
        three_d = False
        for v in requested_variables:
            out[v] = [np.array(m[it], dtype=np.float32, copy=True) for m in fields[v]]
            three_d |= out[v][0].ndim == 3
        out['z'] = _r.zlev if three_d else 0
        return out
    base.get_variables = get_variables
    return base


def run_product(case, **model_kw):
    from opendrift_b200.models.oceandrift import OceanDrift
    from opendrift_b200.readers import reader_regular_grid
    return run_case(case, OceanDrift, product_ensemble_reader,
                    lambda lon, lat, z, t, f, name: reader_regular_grid.Reader(lon, lat, z, t, f, name=name), **model_kw)


def check(o, case):
    ref = np.load(GOLDEN)
    e = max(common.max_err_deg(np.asarray(o.elements.lon), np.asarray(o.elements.lat), ref[case + '__lon'], ref[case + '__lat']))
    spread = float(ref[case + '__spread'])
    return e, spread


if __name__ == '__main__':
    from oracle import refrun
    refrun.setup()
    from opendrift.models.oceandrift import OceanDrift as RefOD
    out = {}
    for case in CASES:
        ro = run_case(case, RefOD, reference_ensemble_reader, lambda lon, lat, z, t, f, name: refrun.make_grid_reader(lon, lat, z, t, f, name=name),
                      logfile='/tmp/od_ens.log')
        lon, lat = np.asarray(ro.elements.lon, dtype=np.float64), np.asarray(ro.elements.lat, dtype=np.float64)
        # how far the members drive neighbouring elements apart: the test is only meaningful if that is far above the tolerance
        fx = common.Fixture(CASES[case][0])
        d = np.abs(lon - fx.lon0[:N])
        spread = float(np.abs(d[0::3].mean() - d[2::3].mean()))
        out.update({case + '__lon': lon, case + '__lat': lat, case + '__spread': np.float64(spread)})
        print(case, len(lon), 'member spread (deg)', spread)
    np.savez_compressed(GOLDEN, **out)
    print('wrote', GOLDEN)
