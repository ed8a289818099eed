"""Leeway (search-and-rescue drift) on the GPU path: the reference's model class (opendrift/models/leeway.py)
with the same element type (LeewayObj :49-134), required variables (:144-171), seeding of the per-element leeway
coefficients from the legacy generator (:292-400) and update() (:430-494), the latter as ONE kernel launch per time
step (od_leeway_step: leeway move + current move + jibing).  Euler only, like the reference.

Object categories: the reference reads ~85 categories from OBJECTPROP.DAT (Allen & Plourde 1999 / Allen 2005).  A
path to such a file can be given as `Leeway(d=path)` exactly like the reference; without it the four person-in-water
categories below are available.  Capsizing (`processes:capsizing`, :438-454) runs inside the same launch.
"""
from collections import OrderedDict

import numpy as np

from ..config import CONFIG_LEVEL_ESSENTIAL, CONFIG_LEVEL_BASIC, CONFIG_LEVEL_ADVANCED
from ..elements import LagrangianArray
from .basemodel import OpenDriftSimulation

RIGHT, LEFT = 0, 1

_COLS = ('DWSLOPE', 'DWOFFSET', 'DWSTD', 'CWRSLOPE', 'CWROFFSET', 'CWRSTD', 'CWLSLOPE', 'CWLOFFSET', 'CWLSTD')
# (key, description, downwind slope [%], offset [cm/s], std; crosswind right slope, offset, std; crosswind left ...)
_BUILTIN = [
    ('PIW-1', 'Person-in-water (PIW), unknown state (mean values)', 0.96, 0.00, 12.00, 0.54, 0.00, 9.40, -0.54, 0.00, 9.40),
    ('PIW-2', '>PIW, vertical PFD type III conscious', 0.48, 0.00, 8.30, 0.15, 0.00, 6.70, -0.15, 0.00, 6.70),
    ('PIW-3', '>PIW, sitting, PFD type I or II', 1.60, -3.98, 2.42, 0.13, 0.33, 2.11, -0.13, -0.33, 2.11),
    ('PIW-4', '>PIW, survival suit (face up)', 1.71, 1.12, 3.93, 1.36, -3.30, 1.71, -0.13, -2.65, 1.62),
]


def read_object_properties(path=None):
    """{number: {OBJKEY, Description, DWSLOPE, ...}} from an OBJECTPROP.DAT-style file (three lines per object:
    key, description, nine numbers; leeway.py:186-209) or the built-in categories."""
    props = OrderedDict()
    if path is None:
        for i, row in enumerate(_BUILTIN, start=1):
            props[i] = dict(OBJKEY=row[0], Description=row[1], **dict(zip(_COLS, row[2:])))
        return props
    lines = open(path).readlines()
    for i in range(len(lines) // 3 + 1):
        if i * 3 >= len(lines) or not lines[i * 3].strip():
            break
        vals = [float(x) for x in lines[i * 3 + 2].split()]
        props[i + 1] = dict(OBJKEY=lines[i * 3].split()[0].strip(), Description=lines[i * 3 + 1].strip(),
                            **dict(zip(_COLS, vals)))
    return props


class LeewayObj(LagrangianArray):
    variables = LagrangianArray.add_variables([
        ('object_type', {'dtype': np.uint16, 'units': '1', 'seed': False, 'default': 0}),
        ('orientation', {'dtype': np.uint8, 'units': '1', 'seed': False, 'default': 1,
                         'description': '0/1 is left/right of downwind. Randomly chosen at seed time'}),
        ('jibe_probability', {'dtype': np.float32, 'units': '1/h', 'default': 0.04,
                              'description': 'Probability per hour that an object may change orientation (jibing)'}),
        ('capsized', {'dtype': np.uint8, 'units': '1', 'seed': True, 'default': 0}),
        ('downwind_slope', {'dtype': np.float32, 'units': '%', 'seed': False, 'default': 1}),
        ('crosswind_slope', {'dtype': np.float32, 'units': '1', 'seed': False, 'default': 1}),
        ('downwind_offset', {'dtype': np.float32, 'units': 'cm/s', 'seed': False, 'default': 0}),
        ('crosswind_offset', {'dtype': np.float32, 'units': 'cm/s', 'seed': False, 'default': 0}),
        ('downwind_eps', {'dtype': np.float32, 'units': 'cm/s', 'seed': False, 'default': 0}),
        ('crosswind_eps', {'dtype': np.float32, 'units': 'cm/s', 'seed': False, 'default': 0}),
        ('current_drift_factor', {'dtype': np.float32, 'units': '1', 'default': 1})])


class Leeway(OpenDriftSimulation):
    ElementType = LeewayObj

    required_variables = {
        'x_wind': {'fallback': None},
        'y_wind': {'fallback': None},
        'x_sea_water_velocity': {'fallback': None},
        'y_sea_water_velocity': {'fallback': None},
        'sea_surface_wave_stokes_drift_x_velocity': {'fallback': 0, 'skip_if': ['drift:stokes_drift', 'is', False]},
        'sea_surface_wave_stokes_drift_y_velocity': {'fallback': 0, 'skip_if': ['drift:stokes_drift', 'is', False]},
        'land_binary_mask': {'fallback': None},
    }

    def __init__(self, d=None, *args, **kwargs):
        self.leewayprop = read_object_properties(d)
        super().__init__(*args, **kwargs)
        descriptions = [p['Description'] for p in self.leewayprop.values()]
        self._add_config({
            'seed:object_type': {'type': 'enum', 'enum': descriptions, 'default': descriptions[0],
                                 'level': CONFIG_LEVEL_ESSENTIAL, 'description': 'Leeway object category for this simulation'},
            'seed:jibe_probability': {'type': 'float', 'default': 0.04, 'min': 0, 'max': 1, 'units': 'probability',
                                      'level': CONFIG_LEVEL_BASIC,
                                      'description': 'Probability per hour for jibing (objects changing orientation)'},
            'processes:capsizing': {'type': 'bool', 'default': False, 'level': CONFIG_LEVEL_BASIC,
                                    'description': 'If True, elements can be capsized when wind exceeds threshold given by '
                                                   'config item capsize:wind_threshold'},
            'capsizing:wind_threshold': {'type': 'float', 'default': 30, 'min': 0, 'max': 50, 'units': 'm/s',
                                         'level': CONFIG_LEVEL_BASIC, 'description':
                                         'Probability of capsizing per hour is: 0.5 + 0.5tanh((windspeed-wind_threshold)/wind_threshold_sigma)'},
            'capsizing:wind_threshold_sigma': {'type': 'float', 'default': 5, 'min': 0, 'max': 20, 'units': 'm/s',
                                               'level': CONFIG_LEVEL_BASIC,
                                               'description': 'Sigma parameter in parameterization of capsize probability'},
            'capsizing:leeway_fraction': {'type': 'float', 'default': 0.4, 'min': 0, 'max': 1, 'units': 'fraction',
                                          'level': CONFIG_LEVEL_BASIC,
                                          'description': 'Leeway coefficients of capsized elements are multiplied by this factor'},
            'drift:stokes_drift': {'type': 'bool', 'default': False, 'level': CONFIG_LEVEL_ADVANCED,
                                   'description': 'Advection elements with surface Stokes drift.'},
            'gpu:rng': {'type': 'enum', 'enum': ['numpy', 'philox'], 'default': 'numpy', 'level': CONFIG_LEVEL_ADVANCED,
                        'description': 'numpy: jibing draws from the legacy generator on the host (bit parity); philox: on device.'},
        })
        self._set_config_default('general:time_step_minutes', 10)
        self._set_config_default('general:time_step_output_minutes', 60)
        self._set_config_default('drift:max_speed', 5)

    def seed_elements(self, lon, lat, object_type=None, **kwargs):
        """leeway.py:292-400 -- same draws from the legacy generator, in the same order."""
        lon = np.atleast_1d(lon).ravel()
        lat = np.atleast_1d(lat).ravel()
        if kwargs.get('number') is not None:
            number = kwargs['number']
        elif len(lon) > 1:
            number = len(lon)
        else:
            number = self.get_config('seed:number')
        if object_type is None:
            name = self.get_config('seed:object_type')
            for object_type, p in self.leewayprop.items():
                if name in (p['OBJKEY'], p['Description']):
                    break
            else:
                raise ValueError('Object %s not available' % name)
        if object_type not in self.leewayprop:
            raise ValueError('Leeway object type %r is not in the property table (%d built-in categories: %s). Pass the full '
                             'table as the reference does: Leeway(d="<path to an OBJECTPROP.DAT>").'
                             % (object_type, len(self.leewayprop), ', '.join(p['OBJKEY'] for p in self.leewayprop.values())))
        prop = self.leewayprop[object_type]
        orientation = np.r_[:number] % 2
        ones = np.ones_like(orientation)
        downwind_slope = ones * prop['DWSLOPE']
        downwind_offset = ones * prop['DWOFFSET']
        epsdw = self._downwind_eps(number, float(prop['DWSLOPE']), float(prop['DWSTD']))
        rcw = np.random.randn(number)
        right, left = orientation == RIGHT, orientation == LEFT
        crosswind_slope = np.where(right, prop['CWRSLOPE'], prop['CWLSLOPE']).astype(float)
        crosswind_offset = np.where(right, prop['CWROFFSET'], prop['CWLOFFSET']).astype(float)
        crosswind_eps = np.where(right, rcw * prop['CWRSTD'], rcw * prop['CWLSTD'])
        return super().seed_elements(lon, lat, orientation=orientation, object_type=object_type,
                                     downwind_slope=downwind_slope, crosswind_slope=crosswind_slope,
                                     downwind_offset=downwind_offset, crosswind_offset=crosswind_offset,
                                     downwind_eps=epsdw, crosswind_eps=crosswind_eps, **kwargs)

    @staticmethod
    def _downwind_eps(number, dwslope, dwstd):
        """leeway.py:340-349: every element draws N(0,1) values ONE AT A TIME from the legacy generator until its downwind slope
        is non-negative.  Single draws consume the generator's stream exactly as a vector draw does (the polar method's spare
        value is cached either way), so the loop equals: take the stream in order, drop the rejected values, give element i the
        i-th accepted one -- and leave the generator just behind the last value used.  Done here in vector form (20 M
        elements seed in a second instead of a minute); the draws and the generator state are the loop's, bit for bit."""
        state = np.random.get_state()
        need, have = number, []
        used = 0
        while need > 0:
            s = np.random.randn(need + (need >> 3) + 16)
            ok = dwslope + (s * dwstd) / 20.0 >= 0.0
            pos = np.flatnonzero(ok)
            if len(pos) >= need:
                cut = pos[need - 1] + 1              # stream values consumed up to and including the last accepted one
                have.append(s[:cut][ok[:cut]])
                used += cut
                need = 0
            else:
                have.append(s[ok])
                used += len(s)
                need -= len(pos)
        np.random.set_state(state)
        np.random.randn(used)                        # position the generator exactly where the element-by-element loop leaves it
        return np.concatenate(have) * dwstd

    def list_object_categories(self, substr=None):
        for i, p in self.leewayprop.items():
            if substr is None or substr.lower() in (p['Description'] + p['OBJKEY']).lower():
                print('%i %s %s' % (i, p['OBJKEY'], p['Description']))

    def report_missing_variables(self):
        """basemodel/__init__.py:2249, 2501-2515: elements whose wind or current is missing leave as 'missing_data' at the top of the
        loop -- before update() draws np.random.random(n) for the jibing (and the capsizing draws), so n and with it the legacy
        generator's stream are the reference's from the step on which an element leaves the readers' coverage.  With the
        reference's draws (gpu:rng = numpy) the two vector pairs are therefore sampled here as well (two launches); the device
        generator is keyed by element ID and does not depend on n: there the step launch flags the elements itself and they leave
        one output step later."""
        self._missing_reported = False
        if type(self).update is not Leeway.update or self.get_config('gpu:rng') != 'numpy' or self.num_elements_active() == 0:
            return
        eng, el, torch = self.engine, self.elements, self.engine.torch
        lon, lat = el.dev('lon', torch.float64), el.dev('lat', torch.float64)
        missing = None
        for xname, yname in (('x_wind', 'y_wind'), ('x_sea_water_velocity', 'y_sea_water_velocity')):
            if self.env.reader_for(xname, self.time) is None:
                continue                                       # constants / fallback values: never missing
            for a in eng.interp(self._pair_group(xname, yname, self.time), self.time, lon, lat, pos_f32=el.positions_f32):
                m = ~torch.isfinite(a)
                missing = m if missing is None else (missing | m)
        self._missing_reported = True
        if missing is not None:
            self._deactivate_missing(missing)

    def update(self):
        """leeway.py:430-494 as one launch."""
        eng, el, torch = self.engine, self.elements, self.engine.torch
        t = self.time
        # drift:current_uncertainty[_uniform] / drift:wind_uncertainty: the draws of this step's environment, made at the top of the
        # loop for the elements active then (_predraw_step_uncertainty), are added inside the launch
        noise_kw = {}
        pre = getattr(self, '_noise0', None)
        if pre:
            cu, cuu, wu = self._uncertainty()
            kinds = (1 if cu > 0 else 0) | (2 if cuu > 0 else 0)
            if kinds:
                arr = np.zeros((2, 2, len(el)))
                if 'cur_n' in pre:
                    arr[0, 0], arr[0, 1] = pre['cur_n']
                if 'cur_u' in pre:
                    arr[1, 0], arr[1, 1] = pre['cur_u']
                noise_kw.update(noise_cur=eng.to_device(arr), noise_kinds=kinds)
            if 'wind' in pre:
                noise_kw['noise_wind'] = eng.to_device(np.stack(pre['wind']))
        gw = self._pair_group('x_wind', 'y_wind', t)
        gc = self._pair_group('x_sea_water_velocity', 'y_sea_water_velocity', t)
        n = len(el)
        caps_kw = {}
        if self.get_config('processes:capsizing'):
            caps_kw['capsizing'] = (self.get_config('capsizing:wind_threshold'), self.get_config('capsizing:wind_threshold_sigma'))
            if self.get_config('gpu:rng') == 'numpy':
                # the reference draws np.random.rand(len(eligible)) BEFORE the jibing draws (:443-451): eligible = not yet
                # capsized in forward runs, capsized in backward runs; laid out per element for the kernel
                capsized = el.dev('capsized', torch.uint8).cpu().numpy()
                can = np.where(capsized == (0 if self.time_step.total_seconds() >= 0 else 1))[0]
                draws = np.full(n, 2.0)
                if len(can) > 0:
                    draws[can] = np.random.rand(len(can))
                caps_kw['rand_capsize'] = eng.to_device(draws)
        rand = eng.to_device(np.random.random(n)) if self.get_config('gpu:rng') == 'numpy' else None
        cols = {'dw_slope': 'downwind_slope', 'dw_offset': 'downwind_offset', 'dw_eps': 'downwind_eps',
                'cw_slope': 'crosswind_slope', 'cw_offset': 'crosswind_offset', 'cw_eps': 'crosswind_eps'}
        d = {k: el.dev(v, torch.float32) for k, v in cols.items()}
        d['orientation'] = el.dev('orientation', torch.uint8)
        d['capsized'] = el.dev('capsized', torch.uint8)
        d['jibe_probability'] = el.dev('jibe_probability')
        if d['jibe_probability'].dtype not in (torch.float32, torch.float64):
            d['jibe_probability'] = d['jibe_probability'].to(torch.float64)
        cats = self.status_categories
        if 'missing_data' in cats:
            missing_code = cats.index('missing_data')
        elif getattr(self, '_coast', None) is None and not getattr(self, '_missing_reported', False):
            cats.append('missing_data')
            missing_code = cats.index('missing_data')
        else:
            # a coastline action numbers its own categories ('stranded') when they first occur: 'missing_data' must not take a
            # number before an element is really missing -- provisional number, named at the top of the next step if it was used
            # (likewise when report_missing_variables has already taken out the elements without forcing: nothing is left to flag)
            missing_code = len(cats)
            self._pending_missing_code = missing_code
        eng.leeway_step(gw, gc, t, self.time_step,
                        el.dev('lon', torch.float64), el.dev('lat', torch.float64), d,
                        moving=el.dev('moving', torch.int32), status=el.dev('status', torch.int32),
                        ids=el.dev('ID', torch.int32), rand=rand, seed=self._seed, step_index=self.steps_calculation,
                        capsize_fraction=self.get_config('capsizing:leeway_fraction'),
                        missing_code=missing_code, pos_f32=el.positions_f32, **caps_kw, **noise_kw)
        el.positions_f32 = False
        self._maybe_deactivated = True           # the kernel may have flagged elements with missing forcing
        self.stokes_drift()

    def _pair_group(self, xname, yname, t):
        """Device field group serving a vector pair at time t: the first reader that covers t, or -- when the pair comes
        from `environment:constant:*` / `environment:fallback:*` only (environment.py:499-923 applies them per
        variable) -- a uniform 2 x 2 global group holding those values."""
        r = self.env.reader_for(xname, t)
        if len(self.env.readers_for(xname, t)) > 1:
            # the fused Leeway launch samples one reader per vector pair; the reference would fill the elements the first
            # reader does not cover from the next one (environment.py:613-780) -- refuse rather than be silently wrong
            raise NotImplementedError('Leeway on the GPU path: several readers provide %s at %s; merge them into one '
                                      'reader (priority lists of readers are only followed by OceanDrift)' % (xname, t))
        if r is not None:
            if not hasattr(r, 'group_of'):
                raise NotImplementedError('Leeway on the GPU path needs gridded readers (got %r)' % r)
            return r.group_of(xname)[0]
        vals = []
        for nme in (xname, yname):
            v = self.env.constant(nme)
            if v is None:
                v = self.env.fallback(nme)
            if v is None:
                raise ValueError('No reader, constant or fallback value for %s at %s' % (nme, t))
            vals.append(float(v))
        key = (xname, tuple(vals))
        cache = self.__dict__.setdefault('_uniform_groups', {})
        if key not in cache:
            lon = np.array([-180.0, 180.0], dtype=np.float32)
            lat = np.array([-90.0, 90.0], dtype=np.float32)
            slab = [np.full((2, 2), v, dtype=np.float32) for v in vals]
            cache[key] = self.engine.add_group(lon, lat, None, 2, [t], lambda ti, c: slab[c], tuple(vals))
        return cache[key]

    def _draws_follow_element_order(self):
        # jibing (and capsizing) draw np.random.random(n) in element order (leeway.py:443-451, 483-487)
        return self.get_config('gpu:rng') == 'numpy' or super()._draws_follow_element_order()

    def update_and_diffuse(self):
        if type(self).update is Leeway.update:
            self.update()                        # stock recipe: no host-side environment needed
        else:
            super().update_and_diffuse()
