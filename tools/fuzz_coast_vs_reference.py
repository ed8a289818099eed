"""Random coastline scenarios (random land mask on its own grid, stranding / previous, scheme, direction, release interval, output
interval, sea floor reader with a random action, Leeway or OceanDrift) -- the drop-in classes on the host build of the device sources
beside the UNMODIFIED reference.  Build container only: python tools/fuzz_coast_vs_reference.py FIRST_SEED LAST_SEED"""
import os
import sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np                                   # noqa: E402
import common                                        # noqa: E402
import coastcases as cc                              # noqa: E402
from datetime import timedelta                       # noqa: E402
from hostengine import HostEngine                    # noqa: E402
import opendrift_b200.engine as E                    # noqa: E402
import opendrift_b200.models.basemodel as B          # noqa: E402
eng = HostEngine()
E.default_engine = lambda device=None: eng
B.default_engine = lambda device=None: eng
from oracle import refrun                            # noqa: E402
refrun.setup()
from opendrift.models.oceandrift import OceanDrift as RefOD          # noqa: E402
from opendrift.models.leeway import Leeway as RefLW                  # noqa: E402
from opendrift_b200.models.oceandrift import OceanDrift              # noqa: E402
from opendrift_b200.models.leeway import Leeway                      # noqa: E402
from opendrift_b200.readers import reader_regular_grid               # noqa: E402


def config(seed):
    rng = np.random.default_rng(seed)
    c = dict(seed=seed, leeway=bool(rng.random() < 0.25), action=str(rng.choice(['stranding', 'previous'])),
             scheme=str(rng.choice(['euler', 'runge-kutta', 'runge-kutta4'])), sign=int(rng.choice([1, 1, -1])),
             release=int(rng.choice([0, 0, 2, 5])), out_every=int(rng.choice([1, 1, 3])), speed=float(rng.choice([3.0, 6.0, 9.0])),
             three_d=bool(rng.random() < 0.5), floor=str(rng.choice(['none', 'none', 'lift_to_seafloor', 'previous', 'deactivate'])),
             ocean_only=bool(rng.random() < 0.3), land_frac=float(rng.uniform(0.05, 0.3)), steps=int(rng.integers(5, 11)))
    if c['leeway']:
        c.update(action='stranding', three_d=False, floor='none', scheme='euler')
    return c


def run(kind, c):
    Ref = kind == 'ref'
    fx = common.Fixture('rk4_3d' if c['three_d'] else 'rk4_2d')
    rng = np.random.default_rng(c['seed'] + 7)
    mk = (lambda lon, lat, z, t, f, name: refrun.make_grid_reader(lon, lat, z, t, f, name=name)) if Ref else \
         (lambda lon, lat, z, t, f, name: reader_regular_grid.Reader(lon, lat, z, t, f, name=name))
    Model = (RefLW if Ref else Leeway) if c['leeway'] else (RefOD if Ref else OceanDrift)
    kw = {'logfile': '/tmp/fz_coast.log'} if Ref else {}
    o = Model(loglevel=50, seed=0, **kw)
    u, v = (c['speed'] * fx.u).astype(np.float32), (c['speed'] * fx.v).astype(np.float32)
    o.add_reader(mk(fx.grid_lon, fx.grid_lat, fx.grid_z, fx.times, {common.CUR[0]: u, common.CUR[1]: v}, 'current'))
    nx, ny = int(rng.integers(15, 40)), int(rng.integers(15, 40))
    mlon = np.linspace(float(fx.grid_lon[0]) + 0.003, float(fx.grid_lon[-1]) - 0.002, nx)
    mlat = np.linspace(float(fx.grid_lat[0]) + 0.002, float(fx.grid_lat[-1]) - 0.003, ny)
    mask = (rng.random((ny, nx)) < c['land_frac']).astype(np.float32)
    o.add_reader(mk(mlon, mlat, None, fx.times, {'land_binary_mask': np.repeat(mask[None], len(fx.times), axis=0)}, 'mask'))
    if c['leeway']:
        X, Y = np.meshgrid(np.linspace(0, 1, len(fx.grid_lon)), np.linspace(0, 1, len(fx.grid_lat)))
        wx = np.stack([7.0 * np.cos(0.5 * k) * (1 + 0.3 * X) for k in range(len(fx.times))]).astype(np.float32)
        wy = np.stack([7.0 * np.sin(0.5 * k) * (1 + 0.3 * Y) for k in range(len(fx.times))]).astype(np.float32)
        o.add_reader(mk(fx.grid_lon, fx.grid_lat, None, fx.times, {'x_wind': wx, 'y_wind': wy}, 'wind'))
    seedkw = {}
    if c['floor'] != 'none':
        XX, YY = np.meshgrid(mlon, mlat)
        floor = (60.0 - 45.0 * (XX - mlon[0]) / (mlon[-1] - mlon[0]) + 4.0 * np.sin(7.0 * YY)).astype(np.float32)
        o.add_reader(mk(mlon, mlat, None, fx.times, {'sea_floor_depth_below_sea_level': np.repeat(floor[None], len(fx.times), axis=0)}, 'floor'))
        seedkw['terminal_velocity'] = -0.02
    cfg = {'general:use_auto_landmask': False, 'environment:constant:land_binary_mask': None, 'general:coastline_approximation_precision': None,
           'seed:ocean_only': c['ocean_only'], 'general:coastline_action': c['action']}
    if not c['leeway']:
        cfg.update({'drift:vertical_advection': False, 'drift:advection_scheme': c['scheme'], 'general:seafloor_action': c['floor'] if c['floor'] != 'none' else 'lift_to_seafloor'})
    for k, val in cfg.items():
        o.set_config(k, val)
    if c['sign'] > 0:
        t = fx.start if not c['release'] else [fx.start, fx.start + timedelta(seconds=c['release'] * fx.dt)]
    else:
        t = fx.times[-1] if not c['release'] else [fx.times[-1] - timedelta(seconds=c['release'] * fx.dt), fx.times[-1]]
    n = 300
    if c['leeway']:
        o.seed_elements(lon=fx.lon0[:n], lat=fx.lat0[:n], time=t, object_type=1)
    else:
        z = fx.z0[:n] if c['three_d'] else np.zeros(n, dtype=np.float32)
        o.seed_elements(lon=fx.lon0[:n], lat=fx.lat0[:n], z=z, time=t, **seedkw)
    o.run(steps=c['steps'], time_step=c['sign'] * fx.dt, time_step_output=c['sign'] * c['out_every'] * fx.dt)
    return cc.summary(o)


if __name__ == '__main__':
    bad = 0
    for seed in range(int(sys.argv[1]), int(sys.argv[2])):
        c = config(seed)
        try:
            r, p = run('ref', c), run('prod', c)
            ok = list(r['cats']) == list(p['cats']) and np.array_equal(r['id'], p['id']) and np.array_equal(r['d_id'], p['d_id']) and \
                np.array_equal(r['d_status'], p['d_status'])
            e = 0.0
            if ok and len(r['id']):
                e = max(common.max_err_deg(p['lon'], p['lat'], r['lon'], r['lat']))
                e = max(e, float(np.max(np.abs(p['z'] - r['z']))) * 1e-3)
            if ok and len(r['d_id']):
                e = max(e, max(common.max_err_deg(p['d_lon'], p['d_lat'], r['d_lon'], r['d_lat'])))
            # 'previous': an element that is moved back lands on the float32 value of its earlier position (the reference keeps the
            # previous positions in float32); where the two float64 positions, 1e-10 deg apart, straddle a float32 rounding boundary
            # the restored positions differ by one float32 ulp (2.4e-7 deg at these longitudes) -- seed 59
            tol = 1e-6 if c['leeway'] else (5e-7 if 'previous' in (c['action'], c['floor']) else 5e-8)
            good = ok and e < tol
            bad += not good
            print(seed, 'OK ' if good else 'BAD', 'err %.1e' % e, 'active', len(r['id']), 'deact', len(r['d_id']), list(r['cats']), '' if good else (c, list(p['cats'])))
        except NotImplementedError as ex:
            print(seed, 'REFUSED', str(ex)[:90])
        except Exception as ex:
            bad += 1
            import traceback
            print(seed, 'EXC', repr(ex)[:200], c)
            traceback.print_exc(limit=3)
    print('bad', bad)
