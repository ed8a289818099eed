"""Environment.get_environment(..., profiles=[...]) -- the host face of the environment for callers outside the step kernels
(environment.py:499-923): the recarray, the missing mask and the vertical profiles `(levels, N)` with their `z`.
Expected results from the UNMODIFIED reference: tests/golden/envprof_ref.npz, written by `python tests/envprofcases.py` in the
build container."""
import os
from datetime import timedelta

import numpy as np

import common

GOLDEN = os.path.join(common.GOLDEN, 'envprof_ref.npz')
N = 120
VARIABLES = ['x_sea_water_velocity', 'y_sea_water_velocity', 'ocean_vertical_diffusivity', 'x_wind']
# name -> (fraction of the grid's columns the reader covers, profiles, profiles_depth, seconds after the first slab)
QUERIES = {
    'k_between_slabs': (1.0, ['ocean_vertical_diffusivity'], 30.0, 900),
    'k_and_u_on_a_slab': (1.0, ['ocean_vertical_diffusivity', 'x_sea_water_velocity'], None, 3600),
    'k_partial_coverage': (0.7, ['ocean_vertical_diffusivity'], 50.0, 1500),
    'wind_constant': (1.0, ['x_wind'], 20.0, 900),
}


def model(Model, make_reader, cut, **model_kw):
    fx = common.Fixture('rk4_3d_cfg4')
    o = Model(loglevel=50, **model_kw)
    k = int(len(fx.grid_lon) * cut)
    c = np.ascontiguousarray
    o.add_reader(make_reader(fx.grid_lon[:k], fx.grid_lat, fx.grid_z, fx.times,
                             {common.CUR[0]: c(fx.u[..., :k]), common.CUR[1]: c(fx.v[..., :k]), 'ocean_vertical_diffusivity': c(fx.kdiff[..., :k])}, 'cur'))
    for key, val in {'general:use_auto_landmask': False, 'environment:constant:land_binary_mask': 0, 'general:coastline_action': 'none',
                     'environment:constant:x_wind': 5.0, 'drift:vertical_mixing': True,
                     'environment:fallback:ocean_vertical_diffusivity': 0.02}.items():
        o.set_config(key, val)
    o.seed_elements(lon=fx.lon0[:N], lat=fx.lat0[:N], z=fx.z0[:N], time=fx.start)
    o.run(steps=1, time_step=fx.dt)             # (finalises the environment)
    return o, fx


def query(o, fx, name):
    cut, profiles, depth, secs = QUERIES[name]
    lon, lat, z = fx.lon0[:N].astype(np.float64) + 0.01, fx.lat0[:N].astype(np.float64), fx.z0[:N]
    env, prof, missing = o.env.get_environment(VARIABLES, fx.start + timedelta(seconds=secs), lon, lat, z, profiles=profiles, profiles_depth=depth)
    out = {'missing': np.asarray(missing, dtype=bool), 'z': np.asarray(prof['z'], dtype=np.float64)}
    for v in VARIABLES:
        out['env_' + v] = np.asarray(env[v], dtype=np.float64)
    for v in profiles:
        out['prof_' + v] = np.asarray(prof[v], dtype=np.float64)
    return out


def run_all(Model, make_reader, **model_kw):
    out, models = {}, {}
    for name, (cut, *_rest) in QUERIES.items():
        if cut not in models:
            models[cut] = model(Model, make_reader, cut, **model_kw)
        for k, v in query(*models[cut], name).items():
            out['%s__%s' % (name, k)] = v
    return out


if __name__ == '__main__':
    from oracle import refrun
    refrun.setup()
    from opendrift.models.oceandrift import OceanDrift as RefOD
    out = run_all(RefOD, lambda lon, lat, z, t, f, name: refrun.make_grid_reader(lon, lat, z, t, f, name=name), logfile='/tmp/od_envprof.log')
    for k, v in out.items():
        print(k, v.shape, v.dtype)
    np.savez_compressed(GOLDEN, **out)
    print('wrote', GOLDEN)
