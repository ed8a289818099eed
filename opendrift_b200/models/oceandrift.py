"""OceanDrift on the GPU hot path: the reference's model class (opendrift/models/oceandrift.py) with the same
element type, required variables, configuration keys and update() recipe for the advection path:

    update():  advect_ocean_current -> advect_wind -> stokes_drift -> vertical_advection      (:185-211)
    then the run loop's horizontal_diffusion                                                   (basemodel :2280)

When update() is not overridden by a subclass the whole recipe runs as ONE kernel launch per time step
(od_step_oceandrift); a subclass that overrides update() gets the same helpers as separate launches.
Vertical turbulent mixing (:397-571) is the next row of SURVEY.md 8(f) and not on this path yet.
"""
import numpy as np

from ..config import CONFIG_LEVEL_ESSENTIAL, CONFIG_LEVEL_BASIC, CONFIG_LEVEL_ADVANCED
from ..elements import LagrangianArray
from .basemodel import OpenDriftSimulation


class Lagrangian3DArray(LagrangianArray):
    """oceandrift.py:28-51"""
    variables = LagrangianArray.add_variables([
        ('wind_drift_factor', {'dtype': np.float32, 'units': '1', 'default': 0.02,
                               'description': 'Elements at surface are moved with this fraction of the wind vector'}),
        ('current_drift_factor', {'dtype': np.float32, 'units': '1', 'default': 1,
                                  'description': 'Elements are moved with this fraction of the current vector'}),
        ('terminal_velocity', {'dtype': np.float32, 'units': 'm/s', 'default': 0.,
                               'description': 'Terminal rise/sinking velocity (buoyancy)'})])


class OceanDrift(OpenDriftSimulation):
    ElementType = Lagrangian3DArray

    # oceandrift.py:70-92
    required_variables = {
        'x_sea_water_velocity': {'fallback': 0},
        'y_sea_water_velocity': {'fallback': 0},
        'x_wind': {'fallback': 0},
        'y_wind': {'fallback': 0},
        'upward_sea_water_velocity': {'fallback': 0, 'skip_if': ['drift:vertical_advection', 'is', False]},
        'ocean_vertical_diffusivity': {'fallback': 0, 'skip_if': ['drift:vertical_mixing', 'is', False], 'profiles': True},
        'horizontal_diffusivity': {'fallback': 0},
        'sea_surface_wave_stokes_drift_x_velocity': {'fallback': 0, 'skip_if': ['drift:stokes_drift', 'is', False]},
        'sea_surface_wave_stokes_drift_y_velocity': {'fallback': 0, 'skip_if': ['drift:stokes_drift', 'is', False]},
        'sea_floor_depth_below_sea_level': {'fallback': 10000},
        'land_binary_mask': {'fallback': None},
    }

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._add_config({
            'drift:vertical_advection': {'type': 'bool', 'default': True, 'level': CONFIG_LEVEL_ESSENTIAL,
                                         'description': 'Advect elements with vertical component of ocean current.'},
            'drift:vertical_advection_at_surface': {'type': 'bool', 'default': False, 'level': CONFIG_LEVEL_ADVANCED,
                                                    'description': 'Also advect elements at the surface vertically.'},
            'drift:vertical_mixing': {'type': 'bool', 'default': False, 'level': CONFIG_LEVEL_BASIC,
                                      'description': 'Vertical turbulent mixing (not on the GPU path yet).'},
            'drift:stokes_drift': {'type': 'bool', 'default': True, 'level': CONFIG_LEVEL_ADVANCED,
                                   'description': 'Advection with Stokes drift.'},
            'drift:wind_drift_depth': {'type': 'float', 'default': 0.1, 'min': 0, 'max': 10, 'units': 'meters',
                                       'level': CONFIG_LEVEL_ADVANCED,
                                       'description': 'Wind drift decreases linearly to zero at this depth.'},
            'drift:truncate_ocean_model_below_m': {'type': 'float', 'default': None, 'min': 0, 'max': 10000,
                                                   'units': 'm', 'level': CONFIG_LEVEL_ADVANCED,
                                                   'description': 'Ocean model data are only read down to this depth.'},
        })
        self._set_config_default('drift:max_speed', 2)

    # -- the reference recipe, helper by helper (used when a subclass overrides update()) -------------------
    def update(self):
        self.advect_ocean_current()
        self.advect_wind()
        self.stokes_drift()
        if self.get_config('drift:vertical_mixing'):
            raise NotImplementedError('vertical mixing is not on the GPU path yet (SURVEY.md 8(f)1)')
        self.vertical_advection()

    def vertical_advection(self):
        """oceandrift.py:315-350: z = min(0, z + moving*w*dt) below (or at) the surface."""
        if self.get_config('drift:vertical_advection') is False:
            return
        env = self.environment
        if 'upward_sea_water_velocity' not in env:
            return
        eng, el, torch = self.engine, self.elements, self.engine.torch
        w = env.dev('upward_sea_water_velocity', eng)
        z = el.dev('z')
        mv = el.dev('moving').to(torch.float64)
        ok = (z <= 0) if self.get_config('drift:vertical_advection_at_surface') else (z < 0)
        zn = torch.clamp(z.to(torch.float64) + mv * w.to(torch.float64) * self.time_step.total_seconds(), max=0.0)
        el.set_dev('z', torch.where(ok, zn.to(z.dtype), z))

    # -- fused path -----------------------------------------------------------------------------------------------
    def _fused_ok(self):
        return (type(self).update is OceanDrift.update and type(self).advect_ocean_current is OceanDrift.advect_ocean_current
                and not self.get_config('drift:vertical_mixing') and not self.get_config('drift:relative_wind'))

    def run(self, *args, **kwargs):
        self._use_fused = None
        return super().run(*args, **kwargs)

    def _step_fused(self):
        eng, el, torch = self.engine, self.elements, self.engine.torch
        g = self._current_group(self.time)
        if g is None:
            return False
        t = self.time
        wind_r = self.env.reader_for('x_wind', t)
        wind = wind_r.group_of('x_wind')[0] if wind_r is not None and hasattr(wind_r, 'group_of') else None
        if wind is None and (self._constant_or_none('x_wind') or 0) != 0:
            return False                                  # constant non-zero wind: helper path
        wgrp = None
        if self.get_config('drift:vertical_advection'):
            wr = self.env.reader_for('upward_sea_water_velocity', t)
            wgrp = wr.group_of('upward_sea_water_velocity')[0] if wr is not None and hasattr(wr, 'group_of') else None
            if wgrp is None and (self._constant_or_none('upward_sea_water_velocity') or 0) != 0:
                return False
        D = self._constant_or_none('horizontal_diffusivity')
        if D is None:
            return False                                  # gridded diffusivity: helper path
        rand = None
        n = len(el)
        if D != 0:
            rand = (eng.to_device(np.random.normal(scale=1, size=n)), eng.to_device(np.random.normal(scale=1, size=n)))
        moving = el.dev('moving')
        if moving.dtype != torch.int32:
            moving = moving.to(torch.int32)
        fac = el.dev('current_drift_factor')
        z = self._z_for_sampling()
        el.set_dev('z', z)
        eng.step_oceandrift(g, self.get_config('drift:advection_scheme'), t, self.time_step,
                            el.dev('lon', torch.float64), el.dev('lat', torch.float64), z, factor=fac, moving=moving,
                            truncate_below=self.get_config('drift:truncate_ocean_model_below_m'),
                            wind=wind, wdf=el.dev('wind_drift_factor'),
                            wind_drift_depth=self.get_config('drift:wind_drift_depth'), w_group=wgrp,
                            w_at_surface=self.get_config('drift:vertical_advection_at_surface'), rand=rand,
                            diffusivity=float(D), pos_f32=el.positions_f32)
        el.positions_f32 = False
        return True

    def update_and_diffuse(self):
        """One time step: fused kernel when the stock recipe applies, else the helpers one by one."""
        if self._fused_ok() and self._step_fused():
            return
        _ = self.environment          # start-of-step environment, before anything moves
        self.update()
        OpenDriftSimulation.horizontal_diffusion(self)
