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
import threading
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


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([c.strip() for c in line.split(',')])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], r[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable'], 'samples': 0}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(np.max(mx)), 'reasons': sorted(reasons),
                'samples': len(sm)}


def field_slabs(grid, n_slabs):
    times = syn.slab_times(n_slabs)
    slabs = [syn.double_gyre_uv(grid, (t - syn.T0).total_seconds()) for t in times]
    return times, slabs


# ------------------------------------------------------------------------------------------------
def cpu_port_rate(n, steps, seed=123):
    """Reference CPU algorithm (NumPy/SciPy port) on a bounded sample; returns (rate, seconds, threads)."""
    from oracle import advect_port as ap
    grid = syn.GridSpec()
    times, slabs = field_slabs(grid, syn.n_slabs_for(steps, DT))
    fields = {CUR[0]: np.stack([s[0] for s in slabs]), CUR[1]: np.stack([s[1] for s in slabs])}
    reader = ap.GridReader(grid.lon, grid.lat, grid.z, times, fields)
    lon, lat, z = syn.particle_cloud(n, seed=seed)
    t0 = time.perf_counter()
    ap.run_oceandrift([reader], lon, lat, z, syn.T0, DT, steps, scheme='runge-kutta4')
    dt = time.perf_counter() - t0
    return n * steps / dt, dt, 1


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n = args.ref_particles
    rates = []
    for _ in range(args.warmup):
        cpu_port_rate(max(1000, n // 10), 1)
    t0 = time.perf_counter()
    rate, secs, thr = cpu_port_rate(n, args.steps)
    ms = secs * 1e3 / args.steps
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': 'particle-steps/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'OceanDrift RK4, synthetic 512x512x50 double-gyre u/v, dt=600 s (configs[1]); '
                               'CPU arm on a bounded sample of %d particles per step' % n},
        'cpu_baseline': {'value': rate, 'unit': 'particle-steps/s', 'cores': thr, 'kind': 'port',
                         'sample': '%d particles x %d steps, NumPy/SciPy restatement of the reference path '
                                   '(single process, as the reference runs)' % (n, args.steps)},
        'e2e': {'value': rate, 'unit': 'particle-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
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
    steps_total = args.warmup + args.steps
    grid = syn.GridSpec()
    n_slabs = syn.n_slabs_for(2 * steps_total + 2, DT)
    times = syn.slab_times(n_slabs)

    # forcing slabs: rank 0 builds them on the host; other ranks receive them with an NCCL broadcast
    # (the once-per-reader-time-step exchange of the multi-GPU design; replicated field, sharded particles)
    host_slabs = []
    dev_slabs = []
    for ti in range(n_slabs):
        if rank == 0:
            u, v = syn.double_gyre_uv(grid, (times[ti] - syn.T0).total_seconds())
            hu, hv = torch.from_numpy(u).pin_memory(), torch.from_numpy(v).pin_memory()
            du, dv = hu.to(dev, non_blocking=True), hv.to(dev, non_blocking=True)
        else:
            hu = hv = None
            du = torch.empty((grid.nz, grid.ny, grid.nx), dtype=torch.float32, device=dev)
            dv = torch.empty_like(du)
        if world > 1:
            dist.broadcast(du, 0)
            dist.broadcast(dv, 0)
            if rank != 0:
                hu, hv = du.cpu().pin_memory(), dv.cpu().pin_memory()
        host_slabs.append((hu, hv))
        dev_slabs.append((du, dv))
    torch.cuda.synchronize()

    resident = {'on': True}

    def supplier(ti, c):
        return dev_slabs[ti][c] if resident['on'] else host_slabs[ti][c]

    grp = eng.add_group(grid.lon, grid.lat, grid.z, 2, times, supplier, (0.0, 0.0), n_slots=3)

    lon0, lat0, z0 = syn.particle_cloud(n, seed=1000 + rank)
    h_lon = torch.from_numpy(lon0.astype(np.float64)).pin_memory()
    h_lat = torch.from_numpy(lat0.astype(np.float64)).pin_memory()
    h_z = torch.from_numpy(z0).pin_memory()
    lon, lat, z = h_lon.to(dev), h_lat.to(dev), h_z.to(dev)
    dt = timedelta(seconds=DT)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # optional spatial ordering of the particle arrays (locality of the field gathers)
    def resort():
        nonlocal lon, lat, z
        perm = eng.sort_by_cell(grp, lon, lat, z)
        lon, lat, z = eng.permute(perm, lon), eng.permute(perm, lat), eng.permute(perm, z)

    state = {'t': times[0], 'k': 0}

    def step():
        if args.sort_every and state['k'] % args.sort_every == 0:
            resort()
        eng.advect_current(grp, 'runge-kutta4', state['t'], dt, lon, lat, z, pos_f32=(state['k'] == 0))
        state['t'] += dt
        state['k'] += 1

    # ---- resident run ---------------------------------------------------------------------------
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = eng.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ms = []
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = eng.launches() - l0
    clocks = sampler.stop() if sampler else None

    # dominant kernel alone (CUDA events around single launches of step_kernel<RK4>, no sort / pack)
    for _ in range(3):
        ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tl, ta = lon.clone(), lat.clone()
        torch.cuda.synchronize()
        ka.record()
        eng.advect_current(grp, 'runge-kutta4', state['t'], dt, tl, ta, z)
        kb.record()
        torch.cuda.synchronize()
        kern_ms.append(ka.elapsed_time(kb))
    kernel_ms = float(np.median(kern_ms))

    # ---- end-to-end: host buffers, copies inside the timed region -----------------------------------
    resident['on'] = False
    grp.resident = [None] * grp.n_slots                     # slabs must come from the host again
    o_lon = torch.empty_like(h_lon).pin_memory()
    o_lat = torch.empty_like(h_lat).pin_memory()
    t_e2e = times[0]

    def e2e_step(t):
        lon.copy_(h_lon, non_blocking=True)
        lat.copy_(h_lat, non_blocking=True)
        z.copy_(h_z, non_blocking=True)
        eng.advect_current(grp, 'runge-kutta4', t, dt, lon, lat, z)
        o_lon.copy_(lon, non_blocking=True)
        o_lat.copy_(lat, non_blocking=True)
        h_lon.copy_(o_lon)            # host-side state advance (the caller owns the arrays)
        h_lat.copy_(o_lat)

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        e2e_step(t_e2e)
        torch.cuda.synchronize()
        t_e2e += dt
    barrier()
    w0 = time.perf_counter()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(e2e_steps):
        e2e_step(t_e2e)
        torch.cuda.synchronize()
        t_e2e += dt
    g1.record()
    barrier()
    e2e_ms = max(g0.elapsed_time(g1), (time.perf_counter() - w0) * 1e3 * 0)  # device clock

    # max over ranks
    if world > 1:
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
    field_bytes = 2 * 2 * grid.nx * grid.ny * grid.nz * 4           # two slabs x (u, v)
    state_bytes = 44                                                  # lon,lat r/w (32) + z, moving, cdf (12)
    b_alg = state_bytes * n + field_bytes                             # per launch
    achieved = b_alg / (kernel_ms * 1e-3) / 1e9
    cpu = None
    if not args.no_cpu:
        rate, secs, thr = cpu_port_rate(args.cpu_particles, args.cpu_steps)
        cpu = {'value': rate, 'unit': 'particle-steps/s', 'cores': thr, 'kind': 'port',
               'sample': '%d particles x %d steps (%.1f s) of the same workload through oracle/advect_port.py, the '
                         'NumPy/SciPy restatement of the reference path' % (args.cpu_particles, args.cpu_steps, secs)}
    line = {
        'metric': METRIC, 'value': value, 'unit': 'particle-steps/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'OceanDrift RK4, synthetic 512x512x50 double-gyre u/v reader, %d particles per GPU, '
                               'dt=600 s (BASELINE configs[1]%s)' % (n, '; configs[2] sharding' if world > 1 else ''),
                   'particles_per_gpu': n, 'field': '512x512x50 f32 u,v, hourly slabs', 'scheme': 'runge-kutta4',
                   'sort_every': args.sort_every, 'parallelism': 'particle-index shards x%d, replicated field' % world,
                   'l2': 'inputs larger than L2 (state %.0f MB + forcing %.0f MB per step)' % (n * 20 / 1e6, field_bytes / 1e6)},
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': 'particle-steps/s', 'h2d_bytes_per_step': n * 20, 'd2h_bytes_per_step': n * 16,
                'steps': e2e_steps},
        'gpu_launches': launches,
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'traffic': None, 'peak_source': peak_src, 'kernel': 'step_kernel<RK4>', 'kernel_ms': kernel_ms,
                     'algorithmic_bytes_per_launch': b_alg,
                     'note': 'fp64 RK4 is bound by the FP64 pipe and gather latency, not by algorithmic HBM bytes '
                             '(65 B per particle-step); see DESIGN.md'},
        'cpu_baseline': cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--particles', type=int, default=10_000_000)
    ap.add_argument('--sort-every', type=int, default=0)
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--cpu-particles', type=int, default=50_000)
    ap.add_argument('--cpu-steps', type=int, default=4)
    ap.add_argument('--ref-particles', type=int, default=50_000)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'b200':
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
