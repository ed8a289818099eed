"""Fuzzer (build container only: needs /root/reference): draws random OceanDrift option sets -- scheme, forward / backward, release
interval, wind, Stokes profile, vertical advection, mixing model, diffusion, the three uncertainty settings, truncation, wind-drift
depth, relative wind, drift-factor arrays, terminal velocity, a nested current reader -- and runs the UNMODIFIED reference
(oracle/refrun.py) and the drop-in classes (on tests/hostengine.py, the host build of the device code) side by side.

    python tools/fuzz_vs_reference.py FIRST_SEED LAST_SEED

Prints one line per configuration; BAD lines carry the option set.  75 configurations were run in round 1; it found the double
uncertainty draw with analytical diffusivity models."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, ROOT)
import numpy as np, common, traceback
from datetime import timedelta
from oracle import refrun
refrun.setup()
from hostengine import HostEngine
from opendrift_b200.readers import reader_regular_grid
CUR = common.CUR
full = common.Fixture('rk4_3d_cfg4')      # u, v, w, K, wind, stokes
def build(kind, c):
    fx = full
    mk = (lambda lon, lat, z, t, f, name: refrun.make_grid_reader(lon, lat, z, t, f, name=name)) if kind == 'ref' else \
         (lambda lon, lat, z, t, f, name: reader_regular_grid.Reader(lon, lat, z, t, f, name=name))
    if kind == 'ref':
        from opendrift.models.oceandrift import OceanDrift as M
        o = M(loglevel=50, logfile='/tmp/x.log', seed=c['seed'])
        base = {'general:use_auto_landmask': False, 'environment:constant:land_binary_mask': 0, 'general:coastline_action': 'none'}
    else:
        from opendrift_b200.models.oceandrift import OceanDrift as M
        o = M(loglevel=50, seed=c['seed'], engine=HostEngine())
        base = {'general:use_auto_landmask': False}
    nx = len(fx.grid_lon)
    cut = c.get('cut')            # no fallback value for the current (and the wind): the readers cover the western part only
    f3 = {CUR[0]: fx.u, CUR[1]: fx.v}
    if c['w']: f3['upward_sea_water_velocity'] = fx.w
    if c['mixing'] == 'environment': f3['ocean_vertical_diffusivity'] = fx.kdiff
    if c['chain'] == 'nested':
        h = nx // 2
        o.add_reader(mk(fx.grid_lon[:h], fx.grid_lat, fx.grid_z, fx.times, {CUR[0]: (1.3*fx.u[..., :h]).astype(np.float32), CUR[1]: (0.7*fx.v[..., :h]).astype(np.float32)}, 'nested'))
    elif c['chain'] == 'handover':          # a reader that covers only the first hour (forward) / the last hour (backward), in front
        sl = slice(0, 2) if c['dt'] > 0 else slice(len(fx.times) - 2, len(fx.times))
        o.add_reader(mk(fx.grid_lon, fx.grid_lat, fx.grid_z, fx.times[sl], {CUR[0]: (1.3*fx.u[sl]).astype(np.float32), CUR[1]: (0.7*fx.v[sl]).astype(np.float32)}, 'first_hour'))
    if cut:
        kc = int(nx * cut)
        o.add_reader(mk(fx.grid_lon[:kc], fx.grid_lat, fx.grid_z, fx.times, {k: np.ascontiguousarray((c['speed'] * v[..., :kc]).astype(np.float32) if k in CUR else v[..., :kc]) for k, v in f3.items()}, 'cur'))
        base = dict(base, **{'environment:fallback:x_sea_water_velocity': None, 'environment:fallback:y_sea_water_velocity': None})
        if c['wind'] is True and c.get('wind_none'):
            base = dict(base, **{'environment:fallback:x_wind': None, 'environment:fallback:y_wind': None})
    else:
        o.add_reader(mk(fx.grid_lon, fx.grid_lat, fx.grid_z, fx.times, f3, 'cur'))
    if c['wind'] == 'constant':
        base = dict(base, **{'environment:constant:x_wind': 6.0, 'environment:constant:y_wind': -4.0})
    elif c['wind']: o.add_reader(mk(fx.wind_lon, fx.wind_lat, None, fx.times, {'x_wind': fx.x_wind, 'y_wind': fx.y_wind}, 'wind'))
    if c['stokes']: o.add_reader(mk(fx.grid_lon, fx.grid_lat, None, fx.times, dict(fx.stokes), 'waves'))
    cfg = dict(base)
    cfg['drift:advection_scheme'] = c['scheme']
    cfg['drift:vertical_advection'] = bool(c['w'])
    cfg['drift:stokes_drift'] = bool(c['stokes'])
    if c['stokes']: cfg['drift:stokes_drift_profile'] = c['stokes']
    if c['mixing']:
        cfg['drift:vertical_mixing'] = True
        cfg['vertical_mixing:timestep'] = c['dt_mix']
        cfg['vertical_mixing:diffusivitymodel'] = c['mixing']
    if c['D']: cfg['environment:constant:horizontal_diffusivity'] = c['D']
    for k in ('current_uncertainty', 'current_uncertainty_uniform', 'wind_uncertainty'):
        if c.get(k): cfg['drift:' + k] = c[k]
    if c['truncate']: cfg['drift:truncate_ocean_model_below_m'] = c['truncate']
    if c['wdd'] is not None: cfg['drift:wind_drift_depth'] = c['wdd']
    if c['relative_wind']: cfg['drift:relative_wind'] = True
    if c.get('max_age'): cfg['drift:max_age_seconds'] = c['max_age']
    if c.get('deact'): cfg['drift:deactivate_east_of'] = c['deact']
    for k, v in cfg.items(): o.set_config(k, v)
    n = c['n']
    dt = c['dt']
    t = fx.start if dt > 0 else fx.times[-1] - timedelta(seconds=600)
    if c['release']:
        t = [t, t + timedelta(seconds=abs(dt) * 3)] if dt > 0 else [t - timedelta(seconds=abs(dt) * 3), t]
    kw = dict(lon=fx.lon0[:n], lat=fx.lat0[:n], z=c['z'](fx, n), time=t)
    if c['cdf'] is not None: kw['current_drift_factor'] = c['cdf']
    if c['wdf'] is not None: kw['wind_drift_factor'] = c['wdf']
    if c['tv'] is not None: kw['terminal_velocity'] = c['tv']
    np.random.seed(c['seed'])
    o.seed_elements(**kw)
    o.run(steps=c['steps'], time_step=dt, time_step_output=dt)
    return o
def draw(seed):
    r = np.random.default_rng(seed)
    n = 250
    c = dict(seed=int(seed), n=n, scheme=r.choice(['euler', 'runge-kutta', 'runge-kutta4']), w=bool(r.integers(2)), wind=bool(r.integers(2)),
             stokes=r.choice([None, None, 'Phillips', 'exponential', 'monochromatic']), mixing=r.choice([None, None, 'environment', 'windspeed_Sundby1983', 'windspeed_Large1994', 'constant']),
             dt_mix=float(r.choice([60.0, 100.0, 45.0])), D=float(r.choice([0, 0, 5.0])), truncate=r.choice([None, None, 30.0]),
             wdd=r.choice([None, 0, 0.5]), relative_wind=bool(r.integers(4) == 0), chain=r.choice([None, None, 'nested', 'handover']), release=bool(r.integers(3) == 0),
             dt=float(r.choice([600, 900, -600])), steps=int(r.integers(3, 7)))
    if r.integers(3) == 0: c['current_uncertainty'] = 0.1
    if r.integers(4) == 0: c['current_uncertainty_uniform'] = 0.05
    if c['wind'] and r.integers(4) == 0: c['wind_uncertainty'] = 1.0
    if not c['wind']: c['relative_wind'] = False
    if c['wind'] and r.integers(5) == 0: c['wind'] = 'constant'
    if c['stokes'] and not c['wind']: c['wind'] = True
    if c['mixing'] in ('windspeed_Sundby1983', 'windspeed_Large1994') and not c['wind']: c['wind'] = True
    if r.integers(5) == 0: c['max_age'] = float(abs(c['dt']) * 2.5)
    if r.integers(5) == 0: c['deact'] = float(np.percentile(full.lon0[:n], 70))
    zsel = int(r.integers(3))
    c['z'] = (lambda fx, n: fx.z0[:n]) if zsel == 0 else ((lambda fx, n: 0.0) if zsel == 1 else (lambda fx, n: np.where(np.arange(n) % 2 == 0, 0.0, fx.z0[:n])))
    c['cdf'] = None if r.integers(3) else np.linspace(0.5, 1.0, n).astype(np.float32)
    c['wdf'] = None if r.integers(3) else (0.03 if r.integers(2) else np.linspace(0, 0.04, n).astype(np.float32))
    c['tv'] = None if (not c['mixing'] or r.integers(2)) else 0.001
    if seed >= 420 and r.integers(3) == 0:        # (seeds below 420 were run before this option existed)
        c['cut'], c['speed'], c['wind_none'] = float(r.uniform(0.6, 0.8)), float(r.choice([1.0, 5.0])), bool(r.integers(2))
        if c['dt'] < 0: c['chain'] = None if c['chain'] == 'handover' else c['chain']
        c['steps'] += 4
    return c
bad = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    c = draw(seed)
    desc = {k: (v if not callable(v) and np.ndim(v) == 0 else '...') for k, v in c.items() if k not in ('seed', 'n')}
    try:
        r = build('ref', c)
    except BaseException as ex:
        print(seed, 'ref failed', repr(ex)[:120]); continue
    try:
        p = build('prod', c)
        rid, pid = np.asarray(r.elements.ID), np.asarray(p.elements.ID)
        if not np.array_equal(rid, pid): bad += 1; print(seed, 'BAD ids', len(rid), len(pid), desc); continue
        rd_ = np.asarray(r.elements_deactivated.ID) if len(r.elements_deactivated) else np.zeros(0)
        pd_ = np.asarray(p.elements_deactivated.ID) if p.num_elements_deactivated() else np.zeros(0)
        if not np.array_equal(rd_, pd_): bad += 1; print(seed, 'BAD deactivated ids', len(rd_), len(pd_), desc); continue
        nan = np.isnan(np.asarray(r.elements.lon, float)) if len(rid) else np.zeros(0, bool)
        if len(rid) and not np.array_equal(nan, np.isnan(np.asarray(p.elements.lon, float))): bad += 1; print(seed, 'BAD undefined positions', desc); continue
        if len(rd_) and not (np.array_equal(np.asarray(r.elements_deactivated.status), np.asarray(p.elements_deactivated.status)) and list(r.status_categories) == list(p.status_categories)):
            bad += 1; print(seed, 'BAD status', list(r.status_categories), list(p.status_categories), desc); continue
        sel = ~nan
        e = max(common.max_err_deg(np.asarray(p.elements.lon)[sel], np.asarray(p.elements.lat)[sel], np.asarray(r.elements.lon)[sel], np.asarray(r.elements.lat)[sel])) if sel.any() else 0
        ez = np.nanmax(np.abs(np.asarray(p.elements.z, float) - np.asarray(r.elements.z, float))) if len(rid) else 0
        # (currents sped up five times: the moves are 4 km long and one last-bit difference of a float32 azimuth -- NumPy's float32 arctan2,
        #  DESIGN.md section 3 -- is 1e-8 deg)
        ok = e < (5e-8 if c.get('speed', 1.0) == 1.0 else 2e-7) and ez < (1e-4 if str(c['mixing']).startswith('windspeed') else 1e-5)      # (torch CPU float32 sqrt: DESIGN section 3)
        bad += not ok
        print(seed, 'OK ' if ok else 'BAD', 'err %.1e z %.1e' % (e, ez), 'cut %s deact %d' % (c.get('cut'), len(rd_)), '' if ok else desc)
    except BaseException as ex:
        bad += 1; print(seed, 'EXC', repr(ex)[:200], desc); traceback.print_exc(limit=3)
print('bad', bad)
