"""Reader.get_variables_interpolated at the reader's own precision (float64 for 3-D blocks, float32 for 2-D ones, NaN-masked where
not covered; land_binary_mask from the nearest grid point): the product's reader against the UNMODIFIED reference's
StructuredReader on the same in-memory slabs.  tests/golden/reader_ref.npz is written by `python tests/readercases.py` in the build
container; the tests (host engine and GPU) require BIT equality."""
import os
from datetime import timedelta

import numpy as np

import common

GOLDEN = os.path.join(common.GOLDEN, 'reader_ref.npz')
N = 2000


def inputs():
    fx = common.Fixture('rk4_3d_full')
    rng = np.random.default_rng(7)
    pad = 0.02
    lon = rng.uniform(fx.grid_lon[0] - pad, fx.grid_lon[-1] + pad, N)
    lat = rng.uniform(fx.grid_lat[0] - pad, fx.grid_lat[-1] + pad, N)
    z = -rng.uniform(0, 1.2 * abs(float(fx.grid_z[-1])), N)
    lon[:20], lat[20:40] = fx.grid_lon[-1], fx.grid_lat[-1]          # on the last column / row
    mask = (rng.random((len(fx.grid_lat), len(fx.grid_lon))) < 0.3).astype(np.float32)
    f3 = {common.CUR[0]: fx.u, common.CUR[1]: fx.v}
    f2 = {'x_wind': fx.u[:, 0] * 7, 'y_wind': fx.v[:, 0] * 7, 'land_binary_mask': np.repeat(mask[None], len(fx.times), axis=0)}
    times = [fx.times[0], fx.times[0] + timedelta(seconds=1234), fx.times[1]]
    return fx, f3, f2, lon, lat, z, times


CALLS = [('cur64', [common.CUR[0], common.CUR[1]], 'f64'), ('cur32', [common.CUR[0], common.CUR[1]], 'f32'), ('wind', ['x_wind', 'y_wind'], 'f64'),
         ('mask', ['land_binary_mask'], 'f64')]


def evaluate(r3, r2):
    fx, f3, f2, lon, lat, z, times = inputs()
    out = {}
    for ti, t in enumerate(times):
        for name, variables, zt in CALLS:
            r = r3 if name.startswith('cur') else r2
            zz = z.astype(np.float32) if zt == 'f32' else z
            env, _ = r.get_variables_interpolated(variables, time=t, lon=lon, lat=lat, z=zz)
            for v in variables:
                a = np.ma.masked_invalid(env[v])
                out['%s__%d__%s' % (name, ti, v)] = np.ma.filled(a, np.nan)
    # vertical profiles (profiles=[...]): every layer of the block at the elements, lerped in time
    for ti, t in enumerate(times):
        env, prof = r3.get_variables_interpolated([common.CUR[0], common.CUR[1]], profiles=[common.CUR[0]], profiles_depth=50.0,
                                                  time=t, lon=lon[:300], lat=lat[:300], z=z[:300])
        out['prof__%d__z' % ti] = np.asarray(prof['z'], dtype=np.float64)
        out['prof__%d__%s' % (ti, common.CUR[0])] = np.ma.filled(np.ma.masked_invalid(prof[common.CUR[0]]), np.nan)
    return out


def product_readers(engine):
    from opendrift_b200.readers import reader_regular_grid
    fx, f3, f2, *_ = inputs()
    r3 = reader_regular_grid.Reader(fx.grid_lon, fx.grid_lat, fx.grid_z, fx.times, f3, name='cur')
    r2 = reader_regular_grid.Reader(fx.grid_lon, fx.grid_lat, None, fx.times, f2, name='surface')
    r3.bind(engine)
    r2.bind(engine)
    return r3, r2


def check(engine):
    ref = np.load(GOLDEN)
    got = evaluate(*product_readers(engine))
    assert set(got) == set(ref.files)
    for k in got:
        assert got[k].dtype == ref[k].dtype, (k, got[k].dtype, ref[k].dtype)
        assert np.array_equal(got[k], ref[k], equal_nan=True), (k, np.nanmax(np.abs(got[k] - ref[k])))
    return len(got)


if __name__ == '__main__':
    from oracle import refrun
    fx, f3, f2, *_ = inputs()
    out = evaluate(refrun.make_grid_reader(fx.grid_lon, fx.grid_lat, fx.grid_z, fx.times, f3, name='cur'),
                   refrun.make_grid_reader(fx.grid_lon, fx.grid_lat, None, fx.times, f2, name='surface'))
    for k, v in out.items():
        print(k, v.dtype, int(np.isnan(v).sum()))
    np.savez_compressed(GOLDEN, **out)
