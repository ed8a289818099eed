"""ORACLE (test infrastructure, never the product path): NumPy/SciPy restatement of the
reference's per-timestep advection hot path for geographic regular-grid readers.

Each function cites the reference code it restates.  It keeps the reference's
*algorithm and rounding points* (float32 block -> float64 bilinear per layer ->
float32 -> float64 vertical lerp -> time lerp -> float32 environment -> float32
azimuth/speed -> float64 WGS84 geodesic), including the per-layer loop over every
block layer that dominates the reference's run time, so that it can stand in for the
reference as the CPU baseline on the GPU box (where /root/reference does not exist).

Pinned against the real reference (oracle/refrun.py, run in the build container) by
tests/test_oracle_port.py on the committed fixtures tests/golden/ref_*.npz, which
oracle/make_golden.py writes.
"""
from datetime import datetime, timedelta

import numpy as np
from scipy.ndimage import map_coordinates
from scipy.interpolate import interp1d

from . import geod_karney

FALLBACK = {'x_sea_water_velocity': 0.0, 'y_sea_water_velocity': 0.0,
            'upward_sea_water_velocity': 0.0, 'x_wind': 0.0, 'y_wind': 0.0,
            'horizontal_diffusivity': 0.0, 'ocean_vertical_diffusivity': 0.0,
            'sea_surface_wave_stokes_drift_x_velocity': 0.0,
            'sea_surface_wave_stokes_drift_y_velocity': 0.0, 'sea_surface_wave_significant_height': 0.0}


class GridReader:
    """In-memory regular lon/lat(/z) reader: what reference StructuredReader subclasses
    provide through get_variables() (opendrift/readers/basereader/structured.py:125-147),
    always returning the full grid as one block (as reader_constant_2d.py:45-49)."""

    def __init__(self, lon, lat, z, times, fields):
        self.x = np.asarray(lon, dtype=np.float32)
        self.y = np.asarray(lat, dtype=np.float32)
        self.z = None if z is None else np.asarray(z, dtype=np.float64)
        self.times = list(times)
        self.fields = fields            # var -> (nt, [nz,] ny, nx) float32
        self.variables = list(fields.keys())
        self.xmin, self.xmax = float(self.x.min()), float(self.x.max())
        self.ymin, self.ymax = float(self.y.min()), float(self.y.max())
        self.start_time, self.end_time = self.times[0], self.times[-1]
        self.time_step = (self.times[1] - self.times[0]) if len(self.times) > 1 else None
        self._blocks = {}
        # east-west global coverage (variables.py:289-301) of an exactly periodic grid: the block is the full circle
        # plus one wrapped column (see oracle/refrun.py:make_grid_reader)
        dx = float(self.x[1] - self.x[0])
        self.global_coverage = (self.xmin - 2 * dx <= 0 and self.xmax + 2 * dx >= 360) or \
                               (self.xmin - 2 * dx <= -180 and self.xmax + 2 * dx >= 180)
        self.periodic = bool(self.global_coverage) and abs(len(self.x) * dx - 360.0) < 1e-3 * dx
        self.block_x = self.x
        if self.periodic:
            self.block_x = np.append(self.x, np.float32(self.x[-1] + np.float32(dx))).astype(np.float32)

    def block(self, var, it):
        """The ReaderBlock array of one variable and time: a private float32 copy (the NaN fill of
        Linear2DInterpolator mutates it), NaNs filled towards the sea floor at block creation
        (interpolation/structured.py:58-70, interpolators.py:204-212)."""
        key = (var, it)
        if key not in self._blocks:
            a = np.array(self.fields[var][it], dtype=np.float32, copy=True)
            if self.periodic:
                a = np.concatenate([a, a[..., :1]], axis=-1)
            if a.ndim == 3:
                for i in range(1, a.shape[0]):
                    m = np.isnan(a[i])
                    if m.any():
                        a[i][m] = a[i - 1][m]
            self._blocks[key] = a
        return self._blocks[key]

    # opendrift/readers/basereader/variables.py:402-443 (constant time step branch == list lookup here)
    def nearest_time(self, time):
        if self.start_time == self.end_time:
            return self.start_time, None, 0, 0
        indx = (time - self.start_time).total_seconds() / self.time_step.total_seconds()
        ib, ia = int(np.floor(indx)), int(np.ceil(indx))
        return self.times[ib], self.times[ia], ib, ia


def expand_numpy_array(data):
    """interpolators.py:9-20: replace non-finite cells by the maximum of their 3x3 neighbourhood, in place."""
    from scipy.ndimage import grey_dilation
    if not np.isfinite(data).any():
        return
    mask = ~np.isfinite(data)
    minval = np.finfo(data[mask].dtype).min
    data[mask] = minval
    data[mask] = grey_dilation(data, size=3)[mask]
    data[data == minval] = np.nan


def _linear2d(block2d, xi, yi):
    """Linear2DInterpolator.__call__ (opendrift/readers/interpolation/interpolators.py:113-139) including the
    NaN loop: dilate the block IN PLACE and re-interpolate the missing points with mode='nearest', <= 10 times."""
    if not np.isfinite(block2d).any():
        return np.nan * np.ones(len(xi))
    interp = map_coordinates(block2d, [yi, xi], cval=np.nan, order=1)
    missing = np.where(~np.isfinite(interp))[0]
    i = 0
    while len(missing) > 0:
        i += 1
        if i > 10:
            return interp
        expand_numpy_array(block2d)
        interp[missing] = map_coordinates(block2d, [yi[missing], xi[missing]], cval=np.nan, order=1, mode='nearest')
        missing = np.where(~np.isfinite(interp))[0]
    return interp


def block_interpolate(reader, it, variables, x, y, z, profiles=None):
    """ReaderBlock.interpolate (opendrift/readers/interpolation/structured.py:107-146) for the
    default 'linearNDFast' horizontal + 'linear' vertical interpolators."""
    xg, yg = reader.block_x, reader.y
    # interpolators.py:107-111 (float32 grid end points, promoted to float64 by x)
    xi = (x - xg[0]) / (xg[-1] - xg[0]) * (len(xg) - 1)
    yi = (y - yg[0]) / (yg[-1] - yg[0]) * (len(yg) - 1)
    out = {}
    prof = {}
    lin1d = None
    for var in variables:
        data = reader.block(var, it)
        if data.ndim == 2:
            out[var] = _linear2d(data, xi, yi)              # float32 result
            continue
        if lin1d is None:
            # Linear1DInterpolator.__init__ (interpolators.py:174-193); z is a float32 copy
            zg = reader.z
            zc = z.copy()
            zc[zc < zg.min()] = zg.min()
            zc[zc > zg.max()] = zg.max()
            if zg[1] > zg[0]:
                f = interp1d(zg, range(len(zg)))
            else:
                f = interp1d(zg[::-1], range(len(zg))[::-1])
            interp_zi = f(zc)
            ia = np.floor(interp_zi).astype(np.int8)
            ia[ia < 0] = 0
            ib = np.minimum(ia + 1, len(zg) - 1)
            wa = 1 - (interp_zi - ia)
            lin1d = (ia, ib, wa, np.arange(len(zc)))
        # structured.py:148-163: every layer of the block, float32 results into a float64 array
        nl = data.shape[0]
        horiz = np.empty((nl, len(x)))
        for layer in range(nl):
            horiz[layer, :] = _linear2d(data[layer], xi, yi)
        if profiles is not None and var in profiles:
            prof[var] = horiz                                   # structured.py:137-138: not interpolated in z
        ia, ib, wa, rng = lin1d
        out[var] = horiz[ia, rng] * wa + horiz[ib, rng] * (1 - wa)   # interpolators.py:195-197
    if profiles is not None:
        return out, prof
    return out


def reader_interpolate(reader, variables, time, lon, lat, z, profiles=None):
    """Variables.get_variables_interpolated -> get_variables_interpolated_xy ->
    StructuredReader._get_variables_interpolated_ for a '+proj=latlong' reader
    (opendrift/readers/basereader/variables.py:860-920, 709-858; structured.py:202-400)."""
    if hasattr(reader, 'interpolate'):          # analytical reader on a projected plane (oracle/gyre_port.py)
        assert profiles is None
        return reader.interpolate(variables, time, lon, lat, z)
    lon = np.mod(lon, 360) if reader.xmin >= 0 else np.mod(lon + 180, 360) - 180   # variables.py:259-280, structured.py:205-215
    x, y = lon, lat
    if reader.global_coverage:                                                   # variables.py:239-242: north-south only
        covered = np.where((y >= reader.ymin) & (y <= reader.ymax))[0]
    else:
        covered = np.where((x >= reader.xmin) & (x <= reader.xmax) &
                           (y >= reader.ymin) & (y <= reader.ymax))[0]           # variables.py:229-257
    n = len(x)
    if len(covered) == 0:
        return {v: np.full(n, np.nan) for v in variables}
    xc, yc, zc = x[covered], y[covered], z.copy()[covered]
    t_before, t_after, ib, ia = reader.nearest_time(time)
    if time == t_before:
        t_after = None
    if profiles is not None:
        assert len(covered) == n, 'profiles are restated for fully covered particle sets only'
        env_before, prof_before = block_interpolate(reader, ib, variables, xc, yc, zc, profiles)
    else:
        env_before = block_interpolate(reader, ib, variables, xc, yc, zc)
    env_profiles = None
    if t_after is not None:
        if profiles is not None:
            env_after, prof_after = block_interpolate(reader, ia, variables, xc, yc, zc, profiles)
        else:
            env_after = block_interpolate(reader, ia, variables, xc, yc, zc)
        w = (time - t_before).total_seconds() / (t_after - t_before).total_seconds()  # structured.py:353-364
        env = {v: env_before[v] * (1 - w) + env_after[v] * w for v in variables}
        if profiles is not None:                                                      # structured.py:366-383
            env_profiles = {'z': reader.z}
            for v in prof_before:
                env_profiles[v] = prof_before[v] * (1 - w) + prof_after[v] * w
    else:
        env = env_before
        if profiles is not None:
            env_profiles = dict(prof_before, z=reader.z)
    if profiles is not None:
        return env, env_profiles
    if len(covered) != n:                                                        # variables.py:841-853
        for v in variables:
            tmp = np.nan * np.ones(n)
            tmp[covered] = env[v]
            env[v] = tmp
    return env


NOISE = {'current': 0.0, 'current_uniform': 0.0, 'wind': 0.0}     # drift:*_uncertainty of the run being restated


def get_environment(readers, variables, time, lon, lat, z, fallback=FALLBACK, truncate_below=None, profiles=None):
    """Environment.get_environment for one reader per variable group
    (opendrift/models/basemodel/environment.py:499-923): float32 cast at :695-696,
    fallback fill at :782-791."""
    if truncate_below is not None:
        z = z.copy()
        z[z < -truncate_below] = -truncate_below
    env = {}
    remaining = list(variables)
    for reader in readers:
        group = [v for v in remaining if v in reader.variables]
        if not group:
            continue
        env_profiles = None
        if profiles is not None and set(group) & set(profiles):
            tmp, env_profiles = reader_interpolate(reader, group, time, lon, lat, z, [p for p in profiles if p in group])
            # environment.py:697-724: with a single reader the profile block is written back onto itself through
            # a float32 cast for all layers but the last (z_ind = arange(len(z) - 1)); NaNs of the last layer are
            # filled from the layer above.
            prof = {}
            for k, a in env_profiles.items():
                a = np.array(a, dtype=np.float64)
                if k != 'z':
                    nl = a.shape[0]
                    a[0:nl - 1] = a[0:nl - 1].astype(np.float32)
                    if nl > 1:
                        mb = np.isnan(a[-1])
                        a[-1, mb] = a[-2, mb]
                prof[k] = a
            env['__profiles__'] = prof
        else:
            tmp = reader_interpolate(reader, group, time, lon, lat, z)
        for v in group:
            env[v] = np.asarray(tmp[v]).astype(np.float32)
            remaining.remove(v)
    for v in remaining:
        env[v] = np.full(len(lon), np.nan, dtype=np.float32)
    for v in variables:
        fb = fallback.get(v)
        if fb is not None:
            bad = ~np.isfinite(env[v])
            env[v][bad] = fb
    # environment.py:869-891: uncertainty draws from the legacy generator, per call, in this order
    n = len(lon)
    if 'x_sea_water_velocity' in variables and 'y_sea_water_velocity' in variables:
        std = NOISE['current']
        if std > 0:
            env['x_sea_water_velocity'] += np.random.normal(0, std, n)
            env['y_sea_water_velocity'] += np.random.normal(0, std, n)
        std = NOISE['current_uniform']
        if std > 0:
            env['x_sea_water_velocity'] += np.random.uniform(-std, std, n)
            env['y_sea_water_velocity'] += np.random.uniform(-std, std, n)
    if 'x_wind' in variables and 'y_wind' in variables:
        std = NOISE['wind']
        if std > 0:
            env['x_wind'] += np.random.normal(0, std, n)
            env['y_wind'] += np.random.normal(0, std, n)
    if profiles is not None:
        return env, env.pop('__profiles__', None)
    return env


_GEOD = geod_karney.Geod()


def update_positions(lon, lat, x_vel, y_vel, moving, dt):
    """OpenDriftSimulation.update_positions (opendrift/models/basemodel/__init__.py:4630-4669)."""
    azimuth = np.degrees(np.arctan2(x_vel, y_vel))
    velocity = np.sqrt(x_vel ** 2 + y_vel ** 2)
    velocity = velocity * moving
    lon2, lat2, _ = _GEOD.fwd(lon, lat, azimuth, velocity * dt)
    return lon2, lat2


def advect_ocean_current(readers, scheme, time, dt, lon, lat, z, cdf, moving, env, factor=1,
                         truncate_below=None):
    """PhysicsMethods.advect_ocean_current (opendrift/models/physics_methods.py:611-691),
    including the reference's RK4 stage-4 quirk (half step at :660-666, time t+dt at :669)."""
    factor = factor * cdf
    uv = ['x_sea_water_velocity', 'y_sea_water_velocity']
    ts = timedelta(seconds=dt)
    x_vel, y_vel = env[uv[0]], env[uv[1]]
    if scheme == 'euler':
        return update_positions(lon, lat, factor * x_vel, factor * y_vel, moving, dt)

    def mid(xv, yv):
        az = np.degrees(np.arctan2(xv, yv))
        speed = np.sqrt(xv * xv + yv * yv)
        dist = speed * dt * .5
        lo, la, _ = _GEOD.fwd(lon, lat, az, dist, radians=False)
        return lo, la

    mlon, mlat = mid(x_vel, y_vel)
    e2 = get_environment(readers, uv, time + ts / 2, mlon, mlat, z, truncate_below=truncate_below)
    if scheme == 'runge-kutta':
        return update_positions(lon, lat, factor * e2[uv[0]], factor * e2[uv[1]], moving, dt)
    assert scheme == 'runge-kutta4'
    lon2, lat2 = mid(e2[uv[0]], e2[uv[1]])
    e3 = get_environment(readers, uv, time + ts / 2, lon2, lat2, z, truncate_below=truncate_below)
    lon3, lat3 = mid(e3[uv[0]], e3[uv[1]])
    e4 = get_environment(readers, uv, time + ts, lon3, lat3, z, truncate_below=truncate_below)
    u4 = (x_vel + 2 * e2[uv[0]] + 2 * e3[uv[0]] + e4[uv[0]]) / 6.0
    v4 = (y_vel + 2 * e2[uv[1]] + 2 * e3[uv[1]] + e4[uv[1]]) / 6.0
    return update_positions(lon, lat, u4 * factor, v4 * factor, moving, dt)


def advect_wind(lon, lat, z, wdf_in, env, moving, dt, wind_drift_depth=0.1, factor=1):
    """PhysicsMethods.advect_wind (opendrift/models/physics_methods.py:712-791), relative_wind off."""
    n = len(lon)
    wind_drift_factor = wdf_in.copy()
    if wind_drift_depth == 0:
        surface_only = True
        wdd = 0
    else:
        wdd = np.abs(wind_drift_depth) * np.ones(n)
        surface_only = False
    surface = z >= -wdd
    if surface.sum() == 0:
        return lon, lat
    wdf = wind_drift_factor.copy()
    wdf_air = wdf.copy()
    if not surface_only:
        wdf = wdf * (wdd + z) / wdd
        wdf[z > 0] = wdf_air[z > 0]
    wdf[~surface] = 0.0
    x_wind, y_wind = env['x_wind'].copy(), env['y_wind'].copy()
    speed = np.sqrt(x_wind[surface] * x_wind[surface] + y_wind[surface] * y_wind[surface])
    if wdf[surface].max() == 0 or speed.max() == 0:
        return lon, lat
    return update_positions(lon, lat, x_wind * wdf * factor, y_wind * wdf * factor, moving, dt)


def stokes_drift(lon, lat, z, env, moving, dt, profile='Phillips', factor=1):
    """PhysicsMethods.stokes_drift (opendrift/models/physics_methods.py:793-848) for the variables OceanDrift
    requires: surface Stokes drift from a reader, Hs from a reader or else from wind (significant_wave_height,
    :893-906), wave period from wind (wave_period / _wave_frequency, :908-943), profiles :332-416."""
    import scipy.special
    us = env['sea_surface_wave_stokes_drift_x_velocity']
    vs = env['sea_surface_wave_stokes_drift_y_velocity']
    if np.max(np.array(us + vs)) == 0:
        return lon, lat
    wind_speed = np.sqrt(env['x_wind'] ** 2 + env['y_wind'] ** 2)
    hs_env = env['sea_surface_wave_significant_height']
    wave_height = hs_env if hs_env.max() > 0 else 0.0246 * np.power(wind_speed, 2)
    omega = 5 * np.ones(wind_speed.shape)
    omega[wind_speed > 0] = 0.877 * 9.81 / (1.17 * wind_speed[wind_speed > 0])
    wave_period = (2 * np.pi) / omega
    if np.max(np.array(wave_height)) == 0:
        wave_height = 1
    if np.max(np.array(wave_period)) == 0:
        wave_period = 8
    speed = np.sqrt(us ** 2 + vs ** 2)
    transport = (2. * np.pi / wave_period) * np.power(wave_height, 2) / 16
    if profile == 'monochromatic':
        km = speed / (2 * transport)
        unit = np.exp(2 * km * z)
    elif profile == 'exponential':
        km = speed / (2 * transport)
        ke = km / 3
        unit = np.exp(2.0 * ke * z) / (1.0 - 8.0 * ke * z)
    elif profile == 'Phillips':
        beta = 1
        km = speed * (1 - 2 * beta / 3) / (2 * transport)
        unit = (np.exp(2 * km * z) - beta * np.sqrt(2 * np.pi * km * np.abs(z)) *
                scipy.special.erfc(np.sqrt(2 * km * np.abs(z))))
    else:
        raise NotImplementedError(profile)
    su, sv = us * unit, vs * unit
    zero = speed == 0
    su[zero] = 0
    sv[zero] = 0
    return update_positions(lon, lat, su * factor, sv * factor, moving, dt)


def vertical_advection(z, w, moving, dt, at_surface=False):
    """OceanDrift.vertical_advection (opendrift/models/oceandrift.py:315-350), no SSH correction."""
    z = z.copy()
    applicable = np.where(z <= 0)[0] if at_surface else np.where(z < 0)[0]
    if len(applicable) > 0:
        z[applicable] = np.minimum(0, z[applicable] + moving[applicable] * w[applicable] * dt)
    return z


def vertical_mixing(z, moving, Kprofiles, mixing_z, dt, dt_mix=60.0, sea_floor_depth=10000.0, mix_at_surface=False,
                    rng=np.random):
    """OceanDrift.vertical_mixing (opendrift/models/oceandrift.py:397-571) with diffusivity from the environment
    profiles, zero terminal velocity (update_terminal_velocity is a no-op in OceanDrift, :285-291) and the
    stock surface_stick (:370-374).  Visser random walk, `int(dt/dt_mix)` inner iterations, one
    np.random.random(N) draw each (:524)."""
    n = len(z)
    dt_mix = dt_mix * np.sign(dt)
    Zmin = -1. * (np.float32(sea_floor_depth) * np.ones(n, dtype=np.float32) + np.zeros(n, dtype=np.float32))   # :420
    z_i = range(mixing_z.shape[0])
    z_index = interp1d(-mixing_z, z_i, bounds_error=False, fill_value=(0, len(z_i) - 1))    # :483-487
    ntimes_mix = np.abs(int(dt / dt_mix))
    gradK = -np.gradient(Kprofiles, mixing_z, axis=0)                                         # :500-502
    gradK[np.abs(gradK) < 1e-10] = 0
    w = np.float32(0) * np.ones(n)                                                            # terminal_velocity
    for _ in range(ntimes_mix):
        surface = z == 0
        zi = np.round(z_index(-z)).astype(np.uint16)
        Kz = Kprofiles[zi, range(Kprofiles.shape[1])]
        dKdz = gradK[zi, range(Kprofiles.shape[1])]
        R = 2 * rng.random(n) - 1
        r = 1.0 / 3
        z = z - moving * (dKdz * dt_mix - R * np.sqrt((Kz * np.abs(dt_mix) * 2 / r)))
        reflect = np.where(z >= 0)
        if len(reflect[0]) > 0:
            z[reflect] = -z[reflect]
        bottom = np.where(np.logical_and(z < Zmin, moving == 1))
        if len(bottom[0]) > 0:
            z[bottom] = 2 * Zmin[bottom] - z[bottom]
        z = z + w * dt_mix * moving
        if not mix_at_surface:
            z[surface] = 0.
        above = np.where(z > 0)                                                               # surface_stick
        if len(above[0]) > 0:
            z[above] = 0
    return z


def diffusivity_profiles(model, wind_speed, mld, background=0.0):
    """Analytical eddy-diffusivity columns on 1 m levels, as OceanDrift.vertical_mixing builds them when the diffusivity
    does not come from an ocean model (opendrift/models/oceandrift.py:429-453, get_diffusivity_profile :385-395):
    mixing_z = -arange(0, max(MLD) + 2); K from the wind speed with Large et al. (1994) or Sundby (1983)
    (opendrift/models/physics_methods.py:203-249).  wind_speed and mld are the float32 environment arrays; the dtype
    flow (float32 wind stress, float64 depth ratio) is NumPy's own since the same expressions are evaluated."""
    mixing_z = -np.arange(0, mld.max() + 2)
    wind, depth = np.meshgrid(wind_speed, np.abs(mixing_z))
    if model == 'windspeed_Large1994':
        depth = np.abs(depth)
        windstress = wind * wind * 1.25e-3 * 1.22                      # cd (Kara et al. 2007) * air density
        sigma = depth / mld

        def shape(s):                                                    # vertical shape function of the eddy diffusivity
            g = 1. * s + (-2) * s**2 + 1 * s**3
            g[np.where(g >= 1)] = g[np.where(g >= 1)] * 0.
            return g
        K = mld * 0.2 * 0.4 * shape(sigma) * windstress + sigma * background     # 0.2: the stability function
        K[depth >= mld] = background
    elif model == 'windspeed_Sundby1983':
        K = 76.1e-4 + 2.26e-4 * wind * wind * np.ones(np.atleast_1d(depth.shape))
        K[depth > mld - 1] = (K[depth > mld - 1] + background) / 2
        K[depth >= mld] = background
    else:
        raise ValueError('Unknown diffusivity model: ' + model)
    return K, mixing_z


def horizontal_diffusion(lon, lat, D, moving, dt, rng=np.random):
    """OpenDriftSimulation.horizontal_diffusion (opendrift/models/basemodel/__init__.py:1746-1772):
    two normal draws from the legacy global generator, x first."""
    if len(D) == 0 or D.max() == 0:
        return lon, lat
    adt = np.abs(dt)
    n = len(lon)
    x_vel = moving * np.sqrt(2 * D / adt) * rng.normal(scale=1, size=n)
    y_vel = moving * np.sqrt(2 * D / adt) * rng.normal(scale=1, size=n)
    return update_positions(lon, lat, x_vel, y_vel, moving, dt)


def run_oceandrift(readers, lon, lat, z, start_time, dt, steps, scheme='runge-kutta4',
                   vertical_adv=False, wind=False, wind_drift_depth=0.1, wdf=0.02, cdf=1.0,
                   diffusivity=0.0, seed=0, truncate_below=None, mixing=False, dt_mix=60.0, stokes=None, noise=None,
                   w_at_surface=False, diffusivity_model=None, mld=50.0, background_diffusivity=1.2e-5, resume=False):
    """OpenDriftSimulation.run main loop (opendrift/models/basemodel/__init__.py:2193-2304) +
    OceanDrift.update (opendrift/models/oceandrift.py:185-211), restricted to the hot path:
    no stranding, no deactivation (the synthetic box has no normal flow), Stokes off."""
    NOISE.update({'current': 0.0, 'current_uniform': 0.0, 'wind': 0.0})
    NOISE.update(noise or {})
    if NOISE['wind'] > 0:
        wind = True        # OceanDrift always requests the wind: its uncertainty is added to the fallback wind (0) too, and drifts
    np.random.seed(seed)                       # basemodel/__init__.py:326
    n = len(lon)
    # seeding casts to the declared element dtypes (opendrift/elements/elements.py:156-158)
    if resume:          # continue a run from a state after its first update: float64 positions (bench.py's parity legs)
        lon, lat = np.asarray(lon, dtype=np.float64), np.asarray(lat, dtype=np.float64)
        z = np.asarray(z)
    else:
        lon = np.asarray(lon, dtype=np.float32)
        lat = np.asarray(lat, dtype=np.float32)
        z = np.asarray(z, dtype=np.float32) * np.ones(n, dtype=np.float32)
    # Scalar element properties (defaults, or scalars given to seed_elements) become *float64*
    # arrays when the scheduled elements are released: LagrangianArray.move_elements does
    # ``self_var*np.ones(self_len)`` (opendrift/elements/elements.py:213-216).  Arrays given to
    # seed_elements keep their declared dtype (float32 / int32).
    def prop(v, dtype):
        if np.ndim(v) == 0:
            if n == 1:                  # a single element is not "shorter than the array": no promotion (elements.py:213)
                return np.full(1, v, dtype=dtype)
            return dtype(v) * np.ones(n)
        return np.asarray(v, dtype=dtype)
    cdf = prop(cdf, np.float32)
    wdf_arr = prop(wdf, np.float32)
    moving = prop(1, np.int32)
    variables = ['x_sea_water_velocity', 'y_sea_water_velocity']
    if vertical_adv:
        variables.append('upward_sea_water_velocity')
    if wind:
        variables += ['x_wind', 'y_wind']
    analytic = mixing and diffusivity_model not in (None, 'environment')
    if mixing and not analytic:
        variables.append('ocean_vertical_diffusivity')
    if analytic and 'x_wind' not in variables:
        variables += ['x_wind', 'y_wind']          # OceanDrift always requires the wind (fallback 0)
    if stokes:
        variables += ['sea_surface_wave_stokes_drift_x_velocity', 'sea_surface_wave_stokes_drift_y_velocity',
                      'sea_surface_wave_significant_height']
        if 'x_wind' not in variables:
            variables += ['x_wind', 'y_wind']
    time = start_time
    for _ in range(steps):
        if mixing and not analytic:
            env, prof = get_environment(readers, variables, time, lon, lat, z, truncate_below=truncate_below,
                                        profiles=['ocean_vertical_diffusivity'])
        else:
            env = get_environment(readers, variables, time, lon, lat, z, truncate_below=truncate_below)
        lon0, lat0, z0 = lon, lat, z
        lon, lat = advect_ocean_current(readers, scheme, time, dt, lon, lat, z, cdf, moving, env,
                                        truncate_below=truncate_below)
        if wind:
            lon, lat = advect_wind(lon, lat, z, wdf_arr, env, moving, dt, wind_drift_depth)
        if stokes:
            lon, lat = stokes_drift(lon, lat, z, env, moving, dt, profile=stokes)
        if analytic:
            wind_speed = np.sqrt(env['x_wind']**2 + env['y_wind']**2)                # PhysicsMethods.wind_speed
            Kp, mz = diffusivity_profiles(diffusivity_model, wind_speed, np.float32(mld) * np.ones(n, dtype=np.float32),
                                          background_diffusivity)
            z = vertical_mixing(z, moving, Kp, mz, dt, dt_mix)
        elif mixing:
            z = vertical_mixing(z, moving, prof['ocean_vertical_diffusivity'], prof['z'], dt, dt_mix)
        if vertical_adv:
            z = vertical_advection(z, env['upward_sea_water_velocity'], moving, dt, at_surface=w_at_surface)
        if diffusivity > 0:
            D = np.float32(diffusivity) * np.ones(n, dtype=np.float32)
            lon, lat = horizontal_diffusion(lon, lat, D, moving, dt)
        time = time + timedelta(seconds=dt)
    return lon, lat, z
