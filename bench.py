#!/usr/bin/env python
"""bench.py -- particle-steps/s of the RK4 advection hot path (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

A "step" is one OceanDrift time step (dt = 600 s) of N_local = 10 M particles per GPU: four trilinear +
time interpolations of (u, v) from the synthetic 512x512x50 double-gyre field, three WGS84 geodesic
mid-point moves and the final update_positions -- one launch of step_kernel<RK4> in libodcuda.so.

value  : particle-steps/s with state and forcing slabs resident in HBM (CUDA events, max over ranks)
e2e    : the same steps through the C-ABI with HOST buffers: per step the particle state is copied
         host->device from pinned memory and the new positions device->host; forcing slabs come from
         pinned host memory when the reader time advances
--impl reference : the reference's own CPU algorithm (oracle/advect_port.py, a NumPy/SciPy restatement that
         is bit-identical to the reference on the committed fixtures; /root/reference cannot travel to the
         GPU box) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import time
from datetime import timedelta

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from opendrift_b200 import synthetic as syn  # noqa: E402

DT = 600.0
CUR = ['x_sea_water_velocity', 'y_sea_water_velocity']
METRIC = 'particle-steps/sec (RK4, 10M particles/GPU, 512x512x50 field)'


def measured_peak_hbm():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (one `-lms 100` process, started before the
    warm-up so that the GPU never idles between warm-up and the timed steps; samples are filtered by timestamp)."""
    Q = ('timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.35)          # let the first samples arrive before the timed region starts
        except Exception:
            self.proc = None

    def stop(self, t_begin=None, t_end=None):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable'], 'samples': 0}
        from datetime import datetime as _dt
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            out = ''
        rows = []
        for line in out.strip().splitlines():
            r = [c.strip() for c in line.split(',')]
            try:
                rows.append((_dt.strptime(r[0], '%Y/%m/%d %H:%M:%S.%f'), float(r[1]), float(r[2]), float(r[3]), r[5:9]))
            except Exception:
                continue
        inside = [r for r in rows if t_begin is not None and t_begin <= r[0] <= t_end]
        window = 'timed region' if inside else 'whole run (no sample fell inside the timed region)'
        sm, mx, pw, reasons = [], [], [], set()
        for ts, a, b, c, flags in (inside or rows):
            sm.append(a)
            mx.append(b)
            pw.append(c)
            for name, v in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], flags):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples'], 'samples': 0}
        # samples under load = those above the idle clock
        load = [x for x in sm if x > 0.5 * max(mx)] or sm
        return {'sm_mhz': float(np.median(load)), 'sm_max_mhz': float(np.max(mx)), 'reasons': sorted(reasons),
                'samples': len(sm), 'samples_under_load': len(load), 'power_w_max': float(np.max(pw)), 'window': window}


PERIOD = 10      # the synthetic double gyre repeats every 36000 s = 10 hourly slabs


class PeriodicSlabs:
    """fields[var][time_index] for a field that is periodic in time: PERIOD distinct slabs serve any run length."""

    def __init__(self, slabs, comp):
        self.slabs, self.comp = slabs, comp

    def __getitem__(self, ti):
        return self.slabs[ti % PERIOD][self.comp]


# ------------------------------------------------------------------------------------------------
def cpu_port_rate(n, steps, seed=123):
    """Reference CPU algorithm (NumPy/SciPy port) on a bounded sample; returns (rate, seconds, threads)."""
    from oracle import advect_port as ap
    grid = syn.GridSpec()
    times = syn.slab_times(syn.n_slabs_for(steps, DT))
    slabs = [syn.double_gyre_uv(grid, (t - syn.T0).total_seconds()) for t in times[:PERIOD]]
    w = syn.upward_w(grid)
    fields = {CUR[0]: PeriodicSlabs(slabs, 0), CUR[1]: PeriodicSlabs(slabs, 1),
              'upward_sea_water_velocity': PeriodicSlabs([(w,)] * PERIOD, 0)}
    reader = ap.GridReader(grid.lon, grid.lat, grid.z, times, fields)
    lon, lat, z = syn.particle_cloud(n, seed=seed)
    t0 = time.perf_counter()
    ap.run_oceandrift([reader], lon, lat, z, syn.T0, DT, steps, scheme='runge-kutta4', vertical_adv=True)
    dt = time.perf_counter() - t0
    return n * steps / dt, dt, 1


def run_reference(args):
    """The reference arm: the reference's own CPU algorithm (single process, as the reference runs) with each
    step a bounded sample of the workload, sized so that K + W steps end within ~2 minutes."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n = args.ref_particles
    if n <= 0:
        budget = 3.0e6                                   # particle-steps (about 90 s at ~3.5e4 particle-steps/s)
        n = int(min(100_000, max(2_000, budget / max(1, args.steps + args.warmup))))
    if args.warmup:
        cpu_port_rate(n, args.warmup)
    rate, secs, thr = cpu_port_rate(n, args.steps)
    ms = secs * 1e3 / args.steps
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': 'particle-steps/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'OceanDrift RK4 + vertical advection, synthetic 512x512x50 double-gyre u/v/w reader, dt=600 s (BASELINE configs[1]); '
                               'CPU arm: each step is a bounded sample of %d particles of that workload' % n},
        'cpu_baseline': {'value': rate, 'unit': 'particle-steps/s', 'cores': thr, 'kind': 'port',
                         'sample': '%d particles x %d steps; oracle/advect_port.py = NumPy/SciPy restatement of the reference '
                                   'path, bit-identical to the reference on the committed fixtures; single process, single '
                                   'thread, as the reference runs (/root/reference cannot travel to the GPU box)' % (n, args.steps)},
        'e2e': {'value': rate, 'unit': 'particle-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------

def configs0_double_gyre(eng, torch, n=10_000_000):
    """BASELINE configs[0] -- the reference's analytical double gyre on its stereographic plane -- as one more measured
    kernel: RK4 launches of od_analytic_advect over n particles (CUDA events, median of 5), one step of a subsample
    checked against the oracle port, and the port's own rate on this host.  Guarded by the caller: it never affects the
    headline line."""
    import time
    from datetime import datetime
    from opendrift_b200.readers import reader_double_gyre
    from oracle import advect_port as ap, gyre_port
    t0 = datetime(2000, 1, 1)
    rd = reader_double_gyre.Reader(initial_time=t0, epsilon=0.25, omega=0.628, A=0.25)
    rd.bind(eng, {v: 0.0 for v in rd.variables})
    d = rd.analytic_desc()
    rng = np.random.default_rng(0)
    lon, lat = rd.xy2lonlat(rng.uniform(0.0, 2.0, n), rng.uniform(0.0, 1.0, n))
    dl0, da0 = eng.to_device(lon), eng.to_device(lat)
    out = []
    for _ in range(5):
        tl, ta = dl0.clone(), da0.clone()
        ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ka.record()
        eng.analytic_advect(d, 'runge-kutta4', (0.0, 0.05, 0.1), 0.1, tl, ta)
        kb.record()
        torch.cuda.synchronize()
        out.append(ka.elapsed_time(kb))
    ms = float(np.median(out))
    m = 20000
    pr = gyre_port.DoubleGyreReader(t0, epsilon=0.25, omega=0.628, A=0.25)
    c0 = time.perf_counter()
    env = ap.get_environment([pr], ['x_sea_water_velocity', 'y_sea_water_velocity'], t0, lon[:m], lat[:m], np.zeros(m))
    pl, pa = ap.advect_ocean_current([pr], 'runge-kutta4', t0, 0.1, lon[:m], lat[:m], np.zeros(m), np.ones(m),
                                     np.ones(m, dtype=np.int32), env)
    cpu_s = time.perf_counter() - c0
    k = 6.371e6 * np.pi / 180
    err = float(np.max(np.hypot((tl[:m].cpu().numpy() - pl) * k, (ta[:m].cpu().numpy() - pa) * k)))
    return {'workload': 'reader_double_gyre on +proj=stere sphere, RK4, dt=0.1 s, %d particles (examples/example_double_gyre_advection_schemes.py at scale)' % n,
            'kernel': 'analytic_step_kernel<RK4, F64, SeriesMath>', 'kernel_ms': ms, 'particle_steps_per_s_kernel': n / (ms * 1e-3),
            'algorithmic_bytes_per_launch': 32 * n, 'hbm_GBps_algorithmic': 32 * n / (ms * 1e-3) / 1e9,
            'bound': 'FP64 / transcendental issue (projection, closed-form field and vector rotation evaluated per stage; no field traffic)',
            'max_err_m_vs_port_one_step': err, 'parity_ok': bool(err < 1e-7),
            'cpu_port_particle_steps_per_s': m / cpu_s, 'cpu_cores': 1}


def run_b200(args):
    import torch
    import torch.distributed as dist
    from opendrift_b200.engine import Engine

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    eng = Engine(local)
    dev = eng.device
    n = args.particles
    grid = syn.GridSpec()
    n_times = syn.n_slabs_for(args.warmup + args.steps + 4, DT) + PERIOD
    times = syn.slab_times(n_times)

    # forcing slabs: rank 0 builds them on the host; the other ranks receive them with an NCCL broadcast over
    # NVLink (the once-per-reader-time-step exchange of the multi-GPU design: replicated field, sharded particles)
    host_slabs, dev_slabs = [], []
    t_bcast = 0.0
    for ti in range(PERIOD):
        if rank == 0:
            u, v = syn.double_gyre_uv(grid, (times[ti] - syn.T0).total_seconds())
            hu, hv = torch.from_numpy(u).pin_memory(), torch.from_numpy(v).pin_memory()
            du, dv = hu.to(dev, non_blocking=True), hv.to(dev, non_blocking=True)
        else:
            du = torch.empty((grid.nz, grid.ny, grid.nx), dtype=torch.float32, device=dev)
            dv = torch.empty_like(du)
        if world > 1:
            torch.cuda.synchronize()
            b0 = time.perf_counter()
            dist.broadcast(du, 0)
            dist.broadcast(dv, 0)
            torch.cuda.synchronize()
            t_bcast += time.perf_counter() - b0
            if rank != 0:
                hu, hv = du.cpu().pin_memory(), dv.cpu().pin_memory()
        host_slabs.append((hu, hv))
        dev_slabs.append((du, dv))
    torch.cuda.synchronize()
    resident = {'on': True}

    def supplier(ti, c):
        return dev_slabs[ti % PERIOD][c] if resident['on'] else host_slabs[ti % PERIOD][c]

    grp = eng.add_group(grid.lon, grid.lat, grid.z, 2, times, supplier, (0.0, 0.0), n_slots=3)
    if os.environ.get('OD_BENCH_NOFILL'):
        grp.fill_nan = 0
    # upward_sea_water_velocity of the u/v/w reader (static in time; every rank builds it from the same formula)
    h_w = torch.from_numpy(syn.upward_w(grid)).pin_memory()
    d_w = h_w.to(dev)
    wgrp = eng.add_group(grid.lon, grid.lat, grid.z, 1, times, lambda ti, c: d_w if resident['on'] else h_w, (0.0,), n_slots=3)

    lon0, lat0, z0 = syn.particle_cloud(n, seed=1000 + rank)
    h_lon = torch.from_numpy(lon0.astype(np.float64)).pin_memory()
    h_lat = torch.from_numpy(lat0.astype(np.float64)).pin_memory()
    h_z = torch.from_numpy(z0).pin_memory()
    st = {'lon': h_lon.to(dev), 'lat': h_lat.to(dev), 'z': h_z.to(dev), 't': times[0], 'k': 0}
    dt = timedelta(seconds=DT)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def resort():
        # spatial ordering of the SoA particle arrays (locality of the field gathers); part of the step cost
        perm = eng.sort_by_cell(grp, st['lon'], st['lat'], st['z'])
        for k in ('lon', 'lat', 'z'):
            st[k] = eng.permute(perm, st[k])

    step_events = []
    host_us = []

    def step(record=False):
        if args.sort_every and st['k'] % args.sort_every == 0:
            resort()
        if record:
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record()
            h0 = time.perf_counter()
        eng.step_oceandrift(grp, 'runge-kutta4', st['t'], dt, st['lon'], st['lat'], st['z'], w_group=wgrp, pos_f32=(st['k'] == 0))
        if record:
            host_us.append((time.perf_counter() - h0) * 1e6)
            eb.record()
            step_events.append((ea, eb))
        st['t'] += dt
        st['k'] += 1

    # ---- resident run: state and slabs in HBM -----------------------------------------------------------
    # clock ramp: a fresh process on an idle GPU runs its first second ~20 % slow (power state / clock ramp), far
    # longer than W steps of 2 ms; spin the same kernel on scratch copies (simulation state untouched) first
    sampler = ClockSampler(local) if rank == 0 and not os.environ.get('OD_BENCH_NOSMI') else None
    if sampler:
        sampler.start()
    ramp_t0 = time.perf_counter()
    tl, ta, tz = st['lon'].clone(), st['lat'].clone(), st['z'].clone()
    while time.perf_counter() - ramp_t0 < args.ramp_seconds:
        for _ in range(20):
            eng.step_oceandrift(grp, 'runge-kutta4', st['t'], dt, tl, ta, tz, w_group=wgrp)
        torch.cuda.synchronize()
    del tl, ta, tz
    for _ in range(args.warmup):
        step()
    barrier()
    l0 = eng.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    from datetime import datetime as _dt
    import gc
    gc.collect()
    gc.disable()                       # no collector pauses between launches of the timed steps
    wall0 = _dt.now()
    e0.record()
    for _ in range(args.steps):
        step(record=True)
    e1.record()
    barrier()
    gc.enable()
    wall1 = _dt.now()
    ms_total = e0.elapsed_time(e1)
    launches = eng.launches() - l0
    loop_kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in step_events]))   # includes slab upload / pair packing
    clocks = sampler.stop(wall0, wall1) if sampler else None
    per_step = np.array([a.elapsed_time(b) for a, b in step_events])

    # dominant kernel alone: CUDA events around single launches of the step kernel on the launching stream
    def kernel_alone(fn):
        out = []
        for _ in range(5):
            ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            tl, ta, tz = st['lon'].clone(), st['lat'].clone(), st['z'].clone()
            torch.cuda.synchronize()
            ka.record()
            fn(tl, ta, tz)
            kb.record()
            torch.cuda.synchronize()
            out.append(ka.elapsed_time(kb))
        return float(np.median(out))

    t_now = st['t']
    kernel_ms = kernel_alone(lambda tl, ta, tz: eng.step_oceandrift(grp, 'runge-kutta4', t_now, dt, tl, ta, tz, w_group=wgrp))
    sort_ms = None
    if args.sort_every:
        sa, sb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        sa.record()
        resort()
        sb.record()
        torch.cuda.synchronize()
        sort_ms = sa.elapsed_time(sb)
    # the operation-by-operation replay of the reference (float32 mid-point azimuths, full Karney geodesic), same launch
    exact_kernel_ms = kernel_alone(lambda tl, ta, tz: eng.step_oceandrift(grp, 'runge-kutta4', t_now, dt, tl, ta, tz, w_group=wgrp, fast=0))
    # the opt-in fast arithmetic (float32 sampling, first-order mid-points), same launch
    fast_kernel_ms = kernel_alone(lambda tl, ta, tz: eng.step_oceandrift(grp, 'runge-kutta4', t_now, dt, tl, ta, tz, w_group=wgrp, fast=1))
    # current advection only (no vertical advection): the plain kernel and its TMA-staged variant (shared-memory field boxes)
    uv_kernel_ms = kernel_alone(lambda tl, ta, tz: eng.advect_current(grp, 'runge-kutta4', t_now, dt, tl, ta, tz))
    eng.set_tile(True)
    tile_kernel_ms = kernel_alone(lambda tl, ta, tz: eng.advect_current(grp, 'runge-kutta4', t_now, dt, tl, ta, tz))
    eng.set_tile(False)

    # ---- end-to-end: HOST buffers through Engine.advect_current_host, copies inside the timed region -----
    resident['on'] = False
    grp.resident = [None] * grp.n_slots                     # forcing slabs come from pinned host memory again
    wgrp.resident = [None] * wgrp.n_slots
    o_lon, o_lat, o_z = torch.empty_like(h_lon).pin_memory(), torch.empty_like(h_lat).pin_memory(), torch.empty_like(h_z).pin_memory()
    e2e_steps = max(3, min(args.steps, 20))
    t_e2e = times[0]
    bufs = [(h_lon, h_lat, h_z), (o_lon, o_lat, o_z)]

    def e2e_step(i, t):
        src, dst = bufs[i % 2], bufs[(i + 1) % 2]
        eng.step_oceandrift_host(grp, 'runge-kutta4', t, dt, src[0], src[1], src[2], dst[0], dst[1], dst[2], w_group=wgrp,
                                 chunks=args.e2e_chunks)

    for i in range(2):
        e2e_step(i, t_e2e)
        t_e2e += dt
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for i in range(e2e_steps):
        e2e_step(i, t_e2e)
        t_e2e += dt
    g1.record()
    barrier()
    e2e_ms = g0.elapsed_time(g1)

    # the link the end-to-end number lives on: this step's bytes (20 B in, 20 B out per particle) as two plain pinned
    # copies running concurrently on two streams
    pcie = None
    if rank == 0:
        try:
            s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
            d_in, d_out = torch.empty_like(h_lon, device=dev), torch.empty_like(h_lat, device=dev)
            d_in2, d_z2 = torch.empty_like(h_lat, device=dev), torch.empty_like(h_z, device=dev)
            torch.cuda.synchronize()
            pa, pb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 4
            pa.record()
            for _ in range(reps):
                with torch.cuda.stream(s_in):
                    s_in.wait_event(pa)
                    d_in.copy_(h_lon, non_blocking=True)
                    d_in2.copy_(h_lat, non_blocking=True)
                    d_z2.copy_(h_z, non_blocking=True)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(pa)
                    o_lon.copy_(d_out, non_blocking=True)
                    o_lat.copy_(d_in, non_blocking=True)
                    o_z.copy_(d_z2, non_blocking=True)
            torch.cuda.current_stream().wait_stream(s_in)
            torch.cuda.current_stream().wait_stream(s_out)
            pb.record()
            torch.cuda.synchronize()
            pms = pa.elapsed_time(pb) / reps
            pcie = {'ms_per_step_copies_only': pms, 'particle_steps_per_s_ceiling': n / (pms * 1e-3),
                    'GBps_both_directions': n * 40 / pms / 1e6}
            del d_in, d_out, d_in2, d_z2
        except Exception as exc:        # the probe is informative only
            pcie = {'error': str(exc)}

    if world > 1:       # device-timed, max over ranks
        tt = torch.tensor([ms_total, e2e_ms, kernel_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total, e2e_ms, kernel_ms = [float(x) for x in tt.tolist()]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = n * world * args.steps / (ms_total * 1e-3)
    e2e_value = n * world * e2e_steps / (e2e_ms * 1e-3)
    peak, peak_src = measured_peak_hbm()
    field_bytes = 2 * 3 * grid.nx * grid.ny * grid.nz * 4           # two time slabs x (u, v, w) float32
    state_bytes = 52                                                  # lon, lat read + written (32), z, moving, factor (12), z updated (8)
    b_alg = state_bytes * n + field_bytes                             # 83.5 B per particle-step at 10 M with w
    achieved = b_alg / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, 'profiles', 'step_kernel_traffic.json')
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get('dram_bytes_per_launch')
        except Exception:
            pass
    cpu = None
    if not args.no_cpu:
        rate, secs, thr = cpu_port_rate(args.cpu_particles, args.cpu_steps)
        cpu = {'value': rate, 'unit': 'particle-steps/s', 'cores': thr, 'kind': 'port',
               'sample': '%d particles x %d steps (%.1f s) of the same workload through oracle/advect_port.py, the '
                         'NumPy/SciPy restatement of the reference path (bit-identical to the reference on the '
                         'committed fixtures; single process like the reference)' % (args.cpu_particles, args.cpu_steps, secs)}
    gyre = None
    if world == 1 and not os.environ.get('OD_BENCH_NO_GYRE'):
        try:
            gyre = configs0_double_gyre(eng, torch, min(n, 10_000_000))
        except Exception as ex:          # never let the extra measurement touch the headline
            gyre = {'error': repr(ex)[:300]}
    line = {
        'metric': METRIC, 'value': value, 'unit': 'particle-steps/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'OceanDrift RK4 + vertical advection, synthetic 512x512x50 double-gyre u/v/w reader, %d particles per GPU, '
                               'dt=600 s (BASELINE configs[1]%s)' % (n, '; configs[2] sharding' if world > 1 else ''),
                   'particles_per_gpu': n, 'field': '512x512x50 f32 u,v (hourly slabs) + w', 'scheme': 'runge-kutta4',
                   'sort_every': args.sort_every, 'sort_ms': sort_ms, 'clock_ramp_s': args.ramp_seconds, 'mode': 'default arithmetic OD_MATH_SERIES: bit-exact field sampling (float64 index and weight arithmetic of the reference), '
                           'float64 short-arc series geodesic (round-off accurate, full Karney solution beyond its range)',
                   'parallelism': 'particle-index shards x%d, replicated field (NCCL broadcast of slabs: %.1f ms per slab pair)'
                                  % (world, 1e3 * t_bcast / PERIOD) if world > 1 else 'single GPU',
                   'l2': 'inputs larger than L2 (state %.0f MB + forcing %.0f MB per step)' % (n * 20 / 1e6, field_bytes / 1e6)},
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': 'particle-steps/s', 'h2d_bytes_per_step': n * 20, 'd2h_bytes_per_step': n * 20,
                'steps': e2e_steps, 'api': 'od_step_oceandrift_host through Engine.step_oceandrift_host (pinned host lon/lat/z in and out, %d-chunk '
                       'three-stream copy/compute pipeline inside the C-ABI call)' % args.e2e_chunks,
                'pcie_probe': pcie},
        'gpu_launches': launches,
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'traffic': traffic, 'peak_source': peak_src, 'kernel': 'step_kernel<RK4, extras> (current advection + vertical advection in one launch)', 'kernel_ms': kernel_ms, 'kernel_ms_mean_in_timed_loop': loop_kernel_ms,
                     'loop_ms_p50': float(np.median(per_step)), 'loop_ms_p95': float(np.percentile(per_step, 95)),
                     'loop_ms_first20_mean': float(per_step[:20].mean()), 'host_us_per_launch_median': float(np.median(host_us)), 'host_us_per_launch_max': float(np.max(host_us)),
                     'algorithmic_bytes_per_launch': b_alg,
                     'note': 'this float64 kernel is bound by FP64 issue (the reference samples its float32 fields with float64 '
                             'index and weight arithmetic, reproduced bit for bit), not by its 65 algorithmic bytes per '
                             'particle-step; see DESIGN.md and profiles/'},
        'cpu_baseline': cpu,
        'configs0_double_gyre': gyre,
        'current_only': {'kernel_ms': uv_kernel_ms, 'particle_steps_per_s_kernel': n / (uv_kernel_ms * 1e-3),
                         'note': 'od_advect_current alone (u/v sampling and moves, no vertical advection), same particles'},
        'tma_tile': {'kernel_ms': tile_kernel_ms, 'note': 'opt-in OD_OPT_TILE for current_only: one cp.async.bulk.tensor.4d box per block; same bits; '
                                                          'not faster than L1-served gathers on sorted particles (see DESIGN.md)'},
        'exact_replay_mode': {'kernel_ms': exact_kernel_ms, 'particle_steps_per_s_kernel': n / (exact_kernel_ms * 1e-3),
                              'note': 'OD_MATH_EXACT: the reference arithmetic operation by operation (float32 mid-point azimuth and '
                                      'distance, order-6 Karney geodesic for every move); ~1e-9 deg per step from the default, the size '
                                      'of the reference\'s own float32 arctan2 noise'},
        'fast_mode': {'kernel_ms': fast_kernel_ms, 'particle_steps_per_s_kernel': n / (fast_kernel_ms * 1e-3),
                      'hbm_frac_algorithmic': b_alg / (fast_kernel_ms * 1e-3) / 1e9 / peak,
                      'note': 'opt-in FastMath (float32 sampling, mid-latitude moves on float64 positions), <= 2e-8 deg from the '
                              'reference on the fixtures; not the headline: value/e2e use the default mode'},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=300)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--particles', type=int, default=10_000_000)
    ap.add_argument('--sort-every', type=int, default=20)
    ap.add_argument('--e2e-chunks', type=int, default=12)
    ap.add_argument('--ramp-seconds', type=float, default=1.5)
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--cpu-particles', type=int, default=50_000)
    ap.add_argument('--cpu-steps', type=int, default=4)
    ap.add_argument('--ref-particles', type=int, default=0, help='0 = sized from --steps')
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'b200':
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
