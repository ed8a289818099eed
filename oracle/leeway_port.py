"""ORACLE (test infrastructure): NumPy restatement of the reference's Leeway model for the hot path
(opendrift/models/leeway.py): seeding of the per-element leeway coefficients (:292-400) and update() (:430-494) --
wind leeway move, current move, jibing -- on top of the interpolation / geodesic restated in advect_port.py.
Pinned bit-for-bit to the unmodified reference by tests/test_oracle_port.py on tests/golden/ref_leeway_*.npz."""
from datetime import timedelta

import numpy as np

from . import advect_port as ap

RIGHT, LEFT = 0, 1


def seed_coefficients(number, prop):
    """Leeway.seed_elements (:326-372): draws from the legacy global generator in the reference's order."""
    orientation = np.r_[:number] % 2
    ones = np.ones_like(orientation)
    downwind_slope = ones * prop['DWSLOPE']
    downwind_offset = ones * prop['DWOFFSET']
    dwstd = prop['DWSTD']
    rdw = np.zeros(number)
    epsdw = np.zeros(number)
    for i in range(number):
        rdw[i] = np.random.randn(1)[0]
        epsdw[i] = rdw[i] * dwstd
        while downwind_slope[i] + epsdw[i] / 20.0 < 0.0:
            rdw[i] = np.random.randn(1)[0]
            epsdw[i] = rdw[i] * dwstd
    rcw = np.random.randn(number)
    crosswind_slope = np.zeros(number)
    crosswind_offset = np.zeros(number)
    crosswind_eps = np.zeros(number)
    crosswind_slope[orientation == RIGHT] = prop['CWRSLOPE']
    crosswind_slope[orientation == LEFT] = prop['CWLSLOPE']
    crosswind_offset[orientation == RIGHT] = prop['CWROFFSET']
    crosswind_offset[orientation == LEFT] = prop['CWLOFFSET']
    crosswind_eps[orientation == RIGHT] = rcw[orientation == RIGHT] * prop['CWRSTD']
    crosswind_eps[orientation == LEFT] = rcw[orientation == LEFT] * prop['CWLSTD']
    f32 = np.float32                      # LeewayObj casts the seeded arrays to the declared dtypes
    return dict(orientation=np.uint8(orientation), downwind_slope=f32(downwind_slope), downwind_offset=f32(downwind_offset),
                downwind_eps=f32(epsdw), crosswind_slope=f32(crosswind_slope), crosswind_offset=f32(crosswind_offset),
                crosswind_eps=f32(crosswind_eps))


def run_leeway(readers, lon, lat, start_time, dt, steps, prop, seed=0, jibe_probability=0.04, capsize_fraction=0.4,
               capsizing=None):
    """capsizing: None (processes:capsizing off) or (wind_threshold, wind_threshold_sigma) (leeway.py:438-454)."""
    np.random.seed(seed)
    n = len(lon)
    lon = np.asarray(lon, dtype=np.float32)
    lat = np.asarray(lat, dtype=np.float32)
    el = seed_coefficients(n, prop)
    z = np.float32(0) * np.ones(n)
    moving = np.int32(1) * np.ones(n)                    # scalars become float64 arrays on release
    capsized = np.uint8(0) * np.ones(n)
    jp = np.float32(jibe_probability) * np.ones(n)
    variables = ['x_wind', 'y_wind', 'x_sea_water_velocity', 'y_sea_water_velocity']
    fallback = {v: None for v in variables}
    time = start_time
    for _ in range(steps):
        env = ap.get_environment(readers, variables, time, lon, lat, z, fallback=fallback)
        # Leeway.update (:430-494)
        windspeed = np.sqrt(env['x_wind'] ** 2 + env['y_wind'] ** 2)
        winddir = np.arctan2(env['x_wind'], env['y_wind'])
        if capsizing is not None:
            thr, sig = capsizing
            can = np.where(capsized == (0 if dt >= 0 else 1))[0]          # forward runs capsize, backward runs un-capsize
            if len(can) > 0:
                prob = (.5 + .5 * np.tanh((windspeed[can] - thr) / sig)) * np.abs(float(dt)) / 3600
                flip = can[np.where(np.random.rand(len(can)) < prob)[0]]
                capsized[flip] = 1 - capsized[flip]
        downwind = ((el['downwind_slope'] + el['downwind_eps'] / 20.0) * windspeed + el['downwind_offset'] +
                    el['downwind_eps'] / 2.0) * .01
        crosswind = ((el['crosswind_slope'] + el['crosswind_eps'] / 20.0) * windspeed + el['crosswind_offset'] +
                     el['crosswind_eps'] / 2.0) * .01
        sinth, costh = np.sin(winddir), np.cos(winddir)
        y_leeway = downwind * costh + crosswind * sinth
        x_leeway = -downwind * sinth + crosswind * costh
        x_leeway[capsized == 1] *= capsize_fraction
        y_leeway[capsized == 1] *= capsize_fraction
        lon, lat = ap.update_positions(lon, lat, -x_leeway, y_leeway, moving, dt)
        lon, lat = ap.update_positions(lon, lat, env['x_sea_water_velocity'], env['y_sea_water_velocity'], moving, dt)
        jibe_rate = -np.log(1 - jp) / 3600
        jp_step = 1 - np.exp(-jibe_rate * np.abs(dt))
        jib = jp_step > np.random.random(n)
        el['crosswind_slope'][jib] = -el['crosswind_slope'][jib]
        el['orientation'][jib] = 1 - el['orientation'][jib]
        time = time + timedelta(seconds=dt)
    el['capsized'] = capsized
    return lon, lat, el
