"""Secondary legs of bench.py (N = 1): the same hot path measured through the model classes a user of the reference calls
(`OceanDrift.run()`, `Leeway.run()`), the other BASELINE configurations that fit one GPU (configs[3]: 3-D current + wind + Stokes +
vertical mixing, RK4, 5 M particles; configs[4]: Leeway / Euler, 20 M particles), kernel-alone times with their rooflines, and
parity of each against the CPU oracle.  Nothing here feeds `value` / `e2e` of the headline line; every leg is guarded by the caller.
"""
import time
from datetime import timedelta

import numpy as np

from opendrift_b200 import synthetic as syn

CUR = ['x_sea_water_velocity', 'y_sea_water_velocity']
PERIOD = 10
DT = 600.0
K_M = 6.371e6 * np.pi / 180          # metres per degree (only used to print errors in metres)


class Periodic:
    """fields[var] of a reader: time index -> slab of a field that repeats every PERIOD slabs."""

    def __init__(self, slabs):
        self.slabs = slabs

    def __call__(self, ti):
        return self.slabs[ti % PERIOD]

    def __getitem__(self, ti):          # the oracle port indexes fields[var][ti]
        return self.slabs[ti % PERIOD]


def host_fields(kind, three_d=True):
    """NumPy slabs of one period for a configuration ('cfg2', 'cfg4', 'cfg5'): {reader name: {variable: [PERIOD slabs]}}."""
    g = syn.GridSpec() if three_d else syn.GridSpec(nz=1)
    times = syn.slab_times(PERIOD)
    secs = [(t - syn.T0).total_seconds() for t in times]
    uv = [syn.double_gyre_uv(g, s, three_d=three_d) for s in secs]
    out = {'current': {CUR[0]: [a for a, _ in uv], CUR[1]: [b for _, b in uv]}}
    if kind in ('cfg2', 'cfg4'):
        w = syn.upward_w(g)
        out['current']['upward_sea_water_velocity'] = [w] * PERIOD
    if kind == 'cfg4':
        out['current']['ocean_vertical_diffusivity'] = [syn.vertical_diffusivity(g, s) for s in secs]
    if kind in ('cfg4', 'cfg5'):
        ww = [syn.wind_xy(g, s) for s in secs]
        out['wind'] = {'x_wind': [a for a, _ in ww], 'y_wind': [b for _, b in ww]}
    if kind == 'cfg4':
        st = [syn.stokes_xy(g, s) for s in secs]
        hs = syn.wave_height(g, 0.0)
        out['waves'] = {'sea_surface_wave_stokes_drift_x_velocity': [a for a, _ in st],
                        'sea_surface_wave_stokes_drift_y_velocity': [b for _, b in st],
                        'sea_surface_wave_significant_height': [hs] * PERIOD}
    return g, out


def to_device(fields, eng, torch):
    out = {}
    cache = {}
    for rname, fv in fields.items():
        out[rname] = {}
        for v, slabs in fv.items():
            dev = []
            for a in slabs:
                if id(a) not in cache:
                    cache[id(a)] = torch.from_numpy(np.ascontiguousarray(a)).to(eng.device)
                dev.append(cache[id(a)])
            out[rname][v] = dev
    return out


def product_readers(g, fields, n_times):
    from opendrift_b200.readers import reader_regular_grid
    times = syn.slab_times(n_times)
    return [reader_regular_grid.Reader(g.lon, g.lat, g.z if nm == 'current' else None, times,
                                       {v: Periodic(s) for v, s in fv.items()}, name=nm) for nm, fv in fields.items()]


def port_readers(g, fields, n_times):
    from oracle import advect_port as ap
    times = syn.slab_times(n_times)
    return [ap.GridReader(g.lon, g.lat, g.z if nm == 'current' else None, times, {v: Periodic(s) for v, s in fv.items()})
            for nm, fv in fields.items()]


class StepTimer:
    """CUDA events around the steady-state steps of a model run: wraps one Engine method (the step launch of the model) and
    records an event right before call number `first` and right before the LAST call (`total` calls are expected).  The interval
    holds total - 1 - first complete steps of run() -- step launch, the next step's housekeeping launch, cell sorts, slab
    changes, output columns -- and none of the work run() does once after the last step (restoring the element order, reading the
    output block back).  Pure instrumentation: the wrapped call is unchanged."""

    def __init__(self, eng, torch, method, first, total):
        self.eng, self.torch, self.method, self.first, self.total = eng, torch, method, first, total
        self.calls = 0
        self.e0 = torch.cuda.Event(enable_timing=True)
        self.e1 = torch.cuda.Event(enable_timing=True)
        self.orig = getattr(eng, method)

    def __enter__(self):
        self.stamps = []

        def wrapped(*a, **k):
            self.stamps.append(time.perf_counter())
            if self.calls == self.first:
                self.e0.record()
            if self.calls == self.total - 1:
                self.e1.record()
                self.t_last = time.perf_counter()
            self.calls += 1
            return self.orig(*a, **k)
        setattr(self.eng, self.method, wrapped)
        return self

    def __exit__(self, *exc):
        self.torch.cuda.synchronize()
        self.t1 = time.perf_counter()
        delattr(self.eng, self.method)           # the instance attribute shadows the class method
        return False

    def ms_per_step(self):
        return self.e0.elapsed_time(self.e1) / max(1, self.total - 1 - self.first)

    def host_ms(self):
        """(median, 95th percentile) of the host time between two step launches, steady steps only."""
        d = np.diff(np.array(self.stamps[self.first:])) * 1e3
        return (float(np.median(d)), float(np.percentile(d, 95))) if len(d) else (0.0, 0.0)

    def tail_s(self):
        """Wall time from the launch of the last step to the return of run(): last step + final state + read-back."""
        return self.t1 - self.t_last


def events(torch, fn, reps=5):
    out = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        out.append(a.elapsed_time(b))
    return float(np.median(out))


def _oceandrift(eng, readers, cfg):
    from opendrift_b200.models.oceandrift import OceanDrift
    o = OceanDrift(loglevel=50, seed=0, engine=eng)
    for r in readers:
        o.add_reader(r)
    o.set_config('general:use_auto_landmask', False)
    o.set_config('general:coastline_action', 'none')
    for k, v in cfg.items():
        o.set_config(k, v)
    return o


# ---------------------------------------------------------------------------------------------------------------------------
def leg_api(eng, torch, n, steps, dev_fields, grid, peak):
    """configs[1] through the API the north star names: OceanDrift(...).add_reader(...).seed_elements(...).run(steps=K) --
    release, per-step housekeeping launch, cell sorting, the fused step launch, the output buffer -- with output at the end only and
    with output at every step."""
    cfg = {'drift:advection_scheme': 'runge-kutta4', 'drift:vertical_advection': True, 'drift:stokes_drift': False}
    lon0, lat0, z0 = syn.particle_cloud(n, seed=4242)
    n_times = syn.n_slabs_for(steps + 2, DT) + 1
    out = {}
    for name, every in (('output_at_end', steps), ('output_every_step', 1)):
        k = steps if every != 1 else min(steps, 40)
        o = _oceandrift(eng, product_readers(grid, dev_fields, n_times), cfg)
        w0 = time.perf_counter()
        o.seed_elements(lon=lon0, lat=lat0, z=z0, time=syn.T0)
        w1 = time.perf_counter()
        first = min(3, k - 2)
        l0 = eng.launches()
        with StepTimer(eng, torch, 'step_oceandrift', first, k) as st:
            o.run(steps=k, time_step=DT, time_step_output=every * DT)
        wall = st.t1 - w1
        ms = st.ms_per_step()
        out[name] = {'steps': k, 'ms_per_step_steady': ms, 'particle_steps_per_s_steady': n / (ms * 1e-3),
                     'run_wall_s': wall, 'seed_elements_s': w1 - w0, 'after_last_step_s': st.tail_s(),
                     'particle_steps_per_s_whole_run': n * k / wall,
                     'output_columns': len(o.history['time']), 'gpu_launches_per_step': (eng.launches() - l0) / k,
                     'host_ms_between_launches_p50_p95': st.host_ms()}
        assert o.num_elements_active() == n and st.calls == k
        del o
    out['note'] = ('steady = CUDA events from the launch of step 3 to the launch of the last step inside run(): per step the housekeeping '
                   'launch (outside / output column / age), the fused step launch, the cell sort every 20 steps, slab changes with their '
                   'prefetch, and -- output_every_step -- the read-back of full output blocks; whole_run = wall clock of run() incl. the '
                   'release of the seeded elements to the device (host arrays -> HBM) and, after the last step, the restoring of the element '
                   'order and the read-back of the output block (after_last_step_s)')
    return out


def leg_api_distributed(eng, torch, dist, n, steps, dev_fields, grid, rank, world):
    """configs[2]: OceanDrift.run() under the torch.distributed job -- every rank seeds its own n elements (gpu:shard = none), only
    rank 0's readers hold data, every new slab reaches the other ranks by the broadcast the field group issues on the copy stream.
    Steady-state time per step from CUDA events inside run(), max over ranks."""
    cfg = {'drift:advection_scheme': 'runge-kutta4', 'drift:vertical_advection': True, 'drift:stokes_drift': False, 'gpu:shard': 'none'}
    if dev_fields is None:           # a rank that never reads: one zero slab stands in for the geometry probe of bind()
        z3 = torch.zeros((grid.nz, grid.ny, grid.nx), dtype=torch.float32, device=eng.device)
        dev_fields = {'current': {CUR[0]: [z3] * PERIOD, CUR[1]: [z3] * PERIOD, 'upward_sea_water_velocity': [z3] * PERIOD}}
    lon0, lat0, z0 = syn.particle_cloud(n, seed=5000 + rank)
    n_times = syn.n_slabs_for(steps + 2, DT) + 1
    o = _oceandrift(eng, product_readers(grid, dev_fields, n_times), cfg)
    o.seed_elements(lon=lon0, lat=lat0, z=z0, time=syn.T0)
    b0 = eng.dist.slabs_broadcast
    with StepTimer(eng, torch, 'step_oceandrift', min(3, steps - 2), steps) as st:
        o.run(steps=steps, time_step=DT, time_step_output=steps * DT)
    assert o.num_elements_active() == n
    t = torch.tensor([st.ms_per_step()], dtype=torch.float64, device=eng.device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    return {'steps': steps, 'ms_per_step_steady_max_over_ranks': ms, 'particle_steps_per_s_steady': n * world / (ms * 1e-3),
            'slab_broadcasts_inside_run': eng.dist.slabs_broadcast - b0, 'elements_per_rank': n,
            'note': 'OceanDrift.run() on every rank of the torch.distributed job (gpu:shard = none: each rank seeded its own elements); forcing '
                    'read by rank 0 only and broadcast into the other ranks\' device ring one slab ahead on the copy stream'}


# ---------------------------------------------------------------------------------------------------------------------------
def parity_port_oceandrift(eng, torch, kind, host, grid, n, steps, cfg, port_kw):
    """A small run of the same model / configuration with the reference's draws (gpu:rng numpy) against the CPU oracle."""
    from oracle import advect_port as ap
    n_times = syn.n_slabs_for(steps + 2, DT) + 1
    lon0, lat0, z0 = syn.particle_cloud(n, seed=777)
    start = syn.T0 + timedelta(seconds=1800)
    o = _oceandrift(eng, product_readers(grid, to_device(host, eng, torch), n_times), dict(cfg, **{'gpu:rng': 'numpy'}))
    o.seed_elements(lon=lon0, lat=lat0, z=z0, time=start)
    o.run(steps=steps, time_step=DT, time_step_output=steps * DT)
    c0 = time.perf_counter()
    pl, pa, pz = ap.run_oceandrift(port_readers(grid, host, n_times), lon0, lat0, z0, start, DT, steps, seed=0, **port_kw)
    cpu_s = time.perf_counter() - c0
    lon, lat, z = np.asarray(o.elements.lon), np.asarray(o.elements.lat), np.asarray(o.elements.z)
    err = float(max(np.abs(lon - pl).max(), np.abs(lat - pa).max()))
    dz = float(np.abs(z.astype(np.float64) - pz.astype(np.float64)).max())
    return {'particles': n, 'steps': steps, 'max_err_deg': err, 'max_err_z_m': dz, 'tolerance_deg': 1e-6, 'ok': bool(err < 5e-8 and dz <= 1e-6),
            'against': 'oracle/advect_port.py (bit-identical to the reference on the committed fixtures), same seeds and draws',
            'cpu_port_particle_steps_per_s': n * steps / cpu_s, 'cpu_cores': 1}


def leg_cfg4(eng, torch, n, steps, peak):
    """BASELINE configs[3]: 3-D current + w + K, wind, Stokes drift; vertical mixing (10 inner iterations) + RK4; device generator."""
    grid, host = host_fields('cfg4')
    dev = to_device(host, eng, torch)
    cfg = {'drift:advection_scheme': 'runge-kutta4', 'drift:vertical_advection': True, 'drift:vertical_mixing': True,
           'vertical_mixing:timestep': 60.0, 'drift:stokes_drift_profile': 'Phillips'}
    n_times = syn.n_slabs_for(steps + 2, DT) + 1
    lon0, lat0, z0 = syn.particle_cloud(n, seed=99)
    o = _oceandrift(eng, product_readers(grid, dev, n_times), dict(cfg, **{'gpu:rng': 'philox'}))
    o.seed_elements(lon=lon0, lat=lat0, z=z0, time=syn.T0)
    l0 = eng.launches()
    with StepTimer(eng, torch, 'step_oceandrift', min(3, steps - 2), steps) as st:
        o.run(steps=steps, time_step=DT, time_step_output=steps * DT)
    launches = eng.launches() - l0
    ms = st.ms_per_step()
    # kernel-alone: the mixing launch and the fused step launch (current + wind + w) on the final state
    el = o.elements
    lon, lat, z = el.dev('lon'), el.dev('lat'), el.dev('z')
    rd = {r.name: r for r in o.env.readers.values()}
    gk = rd['current'].group_of('ocean_vertical_diffusivity')[0]
    guv = rd['current'].group_of(CUR[0])[0]
    gw = rd['current'].group_of('upward_sea_water_velocity')[0]
    gwind = rd['wind'].group_of('x_wind')[0]
    t = o.time - timedelta(seconds=DT)
    mv, ids, tv = el.dev('moving', torch.int32), el.dev('ID', torch.int32), el.dev('terminal_velocity')
    mix_ms = events(torch, lambda: eng.vertical_mixing(gk, t, lon, lat, z, 60.0, 10, moving=mv, terminal_velocity=tv, ids=ids, rand=None,
                                                       seed=0, step_index=1))
    wdf, fac = el.dev('wind_drift_factor'), el.dev('current_drift_factor')

    def step():
        tl, ta, tz = lon.clone(), lat.clone(), z.clone()
        eng.step_oceandrift(guv, 'runge-kutta4', t, timedelta(seconds=DT), tl, ta, tz, factor=fac, moving=mv, wind=gwind, wdf=wdf,
                            w_group=gw, diffusivity=0.0)
    clone_ms = events(torch, lambda: (lon.clone(), lat.clone(), z.clone()))
    step_ms = events(torch, step) - clone_ms
    cells = grid.nx * grid.ny * grid.nz * 4
    mix_bytes = n * (16 + 8 + 8 + 4 + 8 + 4) + 2 * cells           # lon, lat, z in (f64 after the first mix), z out, moving, tv, ID + K pair
    step_bytes = n * (32 + 8 + 8 + 4 + 8 + 8 + 8) + 2 * 3 * cells + 2 * 2 * grid.nx * grid.ny * 4
    par = parity_port_oceandrift(eng, torch, 'cfg4', host, grid, 20000, 2, cfg,
                                 dict(scheme='runge-kutta4', vertical_adv=True, wind=True, mixing=True, dt_mix=60.0, stokes='Phillips'))
    del o
    return {'workload': 'OceanDrift, synthetic 512x512x50 u/v/w/K reader + 512x512 wind and Stokes/Hs readers, %d particles, vertical mixing '
                        '(dt 60 s, 10 inner iterations) + RK4 + wind drift + Stokes drift (Phillips) + vertical advection, dt=600 s, through '
                        'OceanDrift.run() (BASELINE configs[3])' % n,
            'value': n / (ms * 1e-3), 'unit': 'particle-steps/s', 'ms_per_step': ms, 'steps': steps, 'rng': 'philox (device, keyed by element ID)',
            'gpu_launches_per_step': launches / steps, 'host_ms_between_launches_p50_p95': st.host_ms(),
            'mix_kernel': {'kernel_ms': mix_ms, 'algorithmic_bytes_per_launch': mix_bytes, 'achieved_GBps': mix_bytes / (mix_ms * 1e-3) / 1e9,
                           'frac_of_hbm_peak': mix_bytes / (mix_ms * 1e-3) / 1e9 / peak,
                           'bound': 'instruction issue / latency of the 10 dependent random-walk iterations per particle (each: level '
                                    'search, K and dK/dz from a register window of the column, Philox draw, sqrt)'},
            'step_kernel_all_extras': {'kernel_ms': step_ms, 'algorithmic_bytes_per_launch': step_bytes,
                                       'achieved_GBps': step_bytes / (step_ms * 1e-3) / 1e9,
                                       'frac_of_hbm_peak': step_bytes / (step_ms * 1e-3) / 1e9 / peak,
                                       'kernel': 'step_spec_kernel<RK4, F64, EXTRAS=1> (current + wind move + vertical advection; csrc/od_spec.cuh)'},
            'parity': par}


def leg_cfg5(eng, torch, n, steps, peak):
    """BASELINE configs[4]: Leeway (Euler), 2-D current + wind readers; device generator for the jibing draws."""
    from opendrift_b200.models.leeway import Leeway
    from oracle import leeway_port as lp
    grid, host = host_fields('cfg5', three_d=False)
    dev = to_device(host, eng, torch)
    n_times = syn.n_slabs_for(steps + 2, DT) + 1

    def model(rng):
        o = Leeway(loglevel=50, seed=0, engine=eng)
        for r in product_readers(grid, dev, n_times):
            o.add_reader(r)
        o.set_config('general:use_auto_landmask', False)
        o.set_config('gpu:rng', rng)
        return o
    lon0, lat0, _ = syn.particle_cloud(n, seed=31, three_d=False)
    o = model('philox')
    w0 = time.perf_counter()
    o.seed_elements(lon=lon0, lat=lat0, time=syn.T0, object_type=1)
    seed_s = time.perf_counter() - w0
    l0 = eng.launches()
    with StepTimer(eng, torch, 'leeway_step', min(3, steps - 2), steps) as st:
        o.run(steps=steps, time_step=DT, time_step_output=steps * DT)
    launches = eng.launches() - l0
    ms = st.ms_per_step()
    # kernel-alone
    el = o.elements
    rd = {r.name: r for r in o.env.readers.values()}
    gw, gc = rd['wind'].group_of('x_wind')[0], rd['current'].group_of(CUR[0])[0]
    cols = {'dw_slope': 'downwind_slope', 'dw_offset': 'downwind_offset', 'dw_eps': 'downwind_eps',
            'cw_slope': 'crosswind_slope', 'cw_offset': 'crosswind_offset', 'cw_eps': 'crosswind_eps'}
    d = {k: el.dev(v, torch.float32) for k, v in cols.items()}
    d['orientation'], d['capsized'], d['jibe_probability'] = el.dev('orientation', torch.uint8), el.dev('capsized', torch.uint8), el.dev('jibe_probability')
    t = o.time - timedelta(seconds=DT)
    lon, lat = el.dev('lon'), el.dev('lat')
    mv, stt, ids = el.dev('moving', torch.int32), el.dev('status', torch.int32), el.dev('ID', torch.int32)
    k_ms = events(torch, lambda: eng.leeway_step(gw, gc, t, timedelta(seconds=DT), lon, lat, d, moving=mv, status=stt, ids=ids, rand=None,
                                                 seed=0, step_index=1, missing_code=1))
    kbytes = n * (32 + 24 + 2 + 1 + d['jibe_probability'].element_size() + 4 + 4 + 4) + 2 * 2 * 2 * grid.nx * grid.ny * 4
    del o
    # parity: the reference's draws, against the oracle
    m, ps = 100000, 3
    pl0, pa0, _ = syn.particle_cloud(m, seed=32, three_d=False)
    start = syn.T0 + timedelta(seconds=1800)
    q = model('numpy')
    q.seed_elements(lon=pl0, lat=pa0, time=start, object_type=1)
    q.run(steps=ps, time_step=DT, time_step_output=ps * DT)
    prop = {k: v for k, v in q.leewayprop[1].items() if k not in ('OBJKEY', 'Description')}
    c0 = time.perf_counter()
    rl, ra, rel = lp.run_leeway(port_readers(grid, host, n_times), pl0, pa0, start, DT, ps, prop, seed=0)
    cpu_s = time.perf_counter() - c0
    err = float(max(np.abs(np.asarray(q.elements.lon) - rl).max(), np.abs(np.asarray(q.elements.lat) - ra).max()))
    same = bool(np.array_equal(np.asarray(q.elements.orientation).astype(np.int64), np.asarray(rel['orientation']).astype(np.int64)))
    return {'workload': 'Leeway (object type 1, PIW-1), Euler, synthetic 512x512 current and wind readers, %d particles, dt=600 s, through '
                        'Leeway.run() (BASELINE configs[4] on one GPU)' % n,
            'value': n / (ms * 1e-3), 'unit': 'particle-steps/s', 'ms_per_step': ms, 'steps': steps, 'seed_elements_s': seed_s,
            'rng': 'philox (device, keyed by element ID)', 'gpu_launches_per_step': launches / steps,
            'host_ms_between_launches_p50_p95': st.host_ms(),
            'leeway_kernel': {'kernel_ms': k_ms, 'algorithmic_bytes_per_launch': kbytes, 'achieved_GBps': kbytes / (k_ms * 1e-3) / 1e9,
                              'frac_of_hbm_peak': kbytes / (k_ms * 1e-3) / 1e9 / peak,
                              'bound': 'FP64 issue (two full geodesic moves per element: leeway, then current) over 72 B of state'},
            'parity': {'particles': m, 'steps': ps, 'max_err_deg': err, 'orientation_equal': same, 'tolerance_deg': 1e-6, 'ok': bool(err < 5e-8 and same),
                       'against': 'oracle/leeway_port.py, same seeds and draws', 'cpu_port_particle_steps_per_s': m * ps / cpu_s, 'cpu_cores': 1}}
