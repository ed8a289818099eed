"""Particle state.

`LagrangianArray` is the host-side container with the reference's semantics
(opendrift/elements/elements.py:22-254): one NumPy array (or a scalar when all elements share the value)
per declared variable, the same dtype table, `extend` and `move_elements` (boolean compaction that keeps
relative order, :197-228).  It is used for scheduled and deactivated elements.

`DeviceElements` holds the *active* elements as SoA buffers in HBM (float64 lon/lat, float32/float64/int32
for the rest -- whatever dtype the reference's NumPy arrays would have, because that dtype selects the
arithmetic of update_positions).  Attribute access returns NumPy arrays like the reference (lazy
device->host copy) so that model subclasses written against `self.elements.<var>` keep working; the
kernels take the device buffers directly and never leave the GPU on the fast path.
"""
import copy
from collections import OrderedDict

import numpy as np


class LagrangianArray:
    variables = OrderedDict([
        ('ID', {'dtype': np.int32, 'seed': False, 'default': -1}),
        ('status', {'dtype': np.int32, 'seed': False, 'default': 0}),
        ('moving', {'dtype': np.int32, 'seed': False, 'default': 1}),
        ('age_seconds', {'dtype': np.float32, 'units': 's', 'seed': False, 'default': 0}),
        ('origin_marker', {'dtype': np.int32, 'unit': '', 'default': 0,
                           'description': 'An integer kept constant during the simulation.'}),
        ('lon', {'dtype': np.float32, 'units': 'degrees_east', 'standard_name': 'longitude', 'seed': False}),
        ('lat', {'dtype': np.float32, 'units': 'degrees_north', 'standard_name': 'latitude', 'seed': False}),
        ('z', {'dtype': np.float32, 'units': 'm', 'standard_name': 'z', 'positive': 'up', 'default': 0})])

    def __init__(self, **kwargs):
        self.variables = copy.deepcopy(self.variables)
        defaults = {v: spec['dtype'](spec['default']) for v, spec in self.variables.items() if 'default' in spec}
        if not kwargs:
            kwargs = {v: [] for v in self.variables}
        missing = set(self.variables) - set(kwargs) - set(defaults)
        if missing:
            raise TypeError('Missing arguments: %s' % sorted(missing))
        extra = (set(kwargs) | set(defaults)) - set(self.variables)
        if extra:
            raise TypeError('Redundant arguments: %s' % sorted(extra))
        lengths = set()
        for k, v in kwargs.items():
            if hasattr(v, 'ndim') and v.ndim > 1:
                kwargs[k] = v = v.ravel()
            try:
                lengths.add(len(v))
            except TypeError:
                pass
        if len(lengths - {1}) > 1:
            raise TypeError('Input arrays must have same length. Lengths given: %s' % sorted(lengths))
        for k, v in defaults.items():
            setattr(self, k, v)
        for k, v in kwargs.items():                     # cast to the declared dtype (elements.py:156-158)
            setattr(self, k, self.variables[k]['dtype'](v))
        self.dtype = np.dtype([(v, spec['dtype']) for v, spec in self.variables.items()])
        if not isinstance(self.status, np.ndarray):
            self.status = self.status * np.ones(self.lon.shape)

    @classmethod
    def add_variables(cls, new_variables):
        variables = cls.variables.copy()
        variables.update(new_variables)
        return variables

    def __len__(self):
        return max(len(np.atleast_1d(getattr(self, v))) for v in self.variables)

    def extend(self, other):
        n_self, n_other = len(self), len(other)
        for v in self.variables:
            a, b = getattr(self, v), getattr(other, v)
            if not isinstance(a, np.ndarray) and not isinstance(b, np.ndarray) and a == b:
                continue
            if not hasattr(a, '__len__'):
                a = a * np.ones(n_self)
            if not hasattr(b, '__len__'):
                b = b * np.ones(n_other)
            setattr(self, v, np.concatenate((a, b)))

    def move_elements(self, other, indices):
        """Move the elements selected by the boolean array `indices` to `other`, keeping order."""
        n_self, n_other = len(self), len(other)
        for v in self.variables:
            a, b = getattr(self, v), getattr(other, v)
            if not isinstance(a, np.ndarray) and not isinstance(b, np.ndarray) and b == a:
                if np.sum(indices) == n_self:
                    setattr(self, v, [])
                continue
            a, b = np.atleast_1d(a), np.atleast_1d(b)
            if len(a) < n_self:
                a = a * np.ones(n_self)            # scalar -> float64 array, as the reference does
            if len(b) < n_other:
                b = b * np.ones(n_other)
            setattr(other, v, np.concatenate((b, a[indices])) if len(a) > 0 else a[indices])
            setattr(self, v, a[~indices])

    def __repr__(self):
        return ''.join('%s: %s\n' % (v, getattr(self, v)) for v in self.variables)


_NP2T = {}


def _torch_dtype(torch, npdt):
    return {np.dtype('float64'): torch.float64, np.dtype('float32'): torch.float32,
            np.dtype('int32'): torch.int32, np.dtype('int64'): torch.int64,
            np.dtype('bool'): torch.bool, np.dtype('int8'): torch.int8,
            np.dtype('uint8'): torch.uint8}[np.dtype(npdt)]


class DeviceElements:
    """Active elements: SoA device buffers with NumPy-returning attribute access."""

    _own = ('variables', '_engine', '_dev', '_host', '_n', 'dtype', 'positions_f32', 'status_touched')

    def __init__(self, element_type, engine):
        object.__setattr__(self, 'variables', copy.deepcopy(element_type.variables))
        object.__setattr__(self, '_engine', engine)
        object.__setattr__(self, '_dev', {})           # name -> device tensor (authoritative unless in _host)
        object.__setattr__(self, '_host', {})          # name -> NumPy array handed out / assigned (authoritative)
        object.__setattr__(self, '_n', 0)
        object.__setattr__(self, 'dtype', np.dtype([(v, s['dtype']) for v, s in self.variables.items()]))
        # lon/lat still carry float32 values (no update_positions yet): selects NumPy's float32 index arithmetic
        object.__setattr__(self, 'positions_f32', True)
        # elements.status was handed out as / assigned from a host array: a subclass may have deactivated elements through it
        object.__setattr__(self, 'status_touched', False)

    def __len__(self):
        return self._n

    # -- NumPy view (reference API) ---------------------------------------------------------
    def __getattr__(self, name):
        variables = object.__getattribute__(self, 'variables')
        if name not in variables:
            raise AttributeError(name)
        host = object.__getattribute__(self, '_host')
        if name == 'status':
            object.__setattr__(self, 'status_touched', True)
        if name not in host:
            dev = object.__getattribute__(self, '_dev')
            host[name] = dev[name].cpu().numpy() if name in dev else np.zeros(0, dtype=variables[name]['dtype'])
            dev.pop(name, None)                       # the caller may modify the array in place
        return host[name]

    def __setattr__(self, name, value):
        if name in DeviceElements._own or name not in self.variables:
            object.__setattr__(self, name, value)
            return
        value = np.asarray(value)
        if value.ndim == 0:
            value = value * np.ones(self._n)
        if name == 'status':
            object.__setattr__(self, 'status_touched', True)
        self._host[name] = value
        self._dev.pop(name, None)

    # -- device view (kernels) ----------------------------------------------------------------
    def dev(self, name, dtype=None):
        """Device tensor of a variable (uploads a host-modified array first)."""
        torch = self._engine.torch
        if name in self._host:
            a = np.ascontiguousarray(self._host.pop(name))
            self._dev[name] = self._engine.to_device(a)
        t = self._dev[name]
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
            self._dev[name] = t
        return t

    def set_dev(self, name, tensor):
        self._dev[name] = tensor
        self._host.pop(name, None)

    def names(self):
        return list(self.variables)

    # -- growth / compaction --------------------------------------------------------------------
    def append_host(self, other, indices):
        """Append `other[indices]` (a host LagrangianArray selection) -- LagrangianArray.move_elements
        seen from the receiving side, including its scalar -> float64 promotion."""
        torch = self._engine.torch
        n_other = len(other)
        n_new = int(np.sum(indices))
        for v in self.variables:
            b = np.atleast_1d(getattr(other, v))
            if len(b) < n_other:
                b = b * np.ones(n_other)
            new = np.ascontiguousarray(b[indices])
            if self._n == 0:
                merged = self._engine.to_device(new)
            else:
                old = self.dev(v)
                res = np.result_type(new.dtype, np.dtype(str(old.dtype).replace('torch.', '')))
                td = _torch_dtype(torch, res)
                merged = torch.cat((old.to(td), self._engine.to_device(new).to(td)))
            self._dev[v] = merged
            self._host.pop(v, None)
        object.__setattr__(self, '_n', self._n + n_new)

    def compact(self):
        """Move the elements with status != 0 out (stable partition on the device, od_partition_active +
        od_permute per column); returns the removed ones as host arrays, or None if nothing was removed."""
        torch = self._engine.torch
        status = self.dev('status')
        if status.dtype != torch.int32:
            status = status.to(torch.int32)
        perm, n_keep = self._engine.partition_active(status)
        if n_keep == self._n:
            return None
        removed = {}
        for v in self.variables:
            t = self._engine.permute(perm, self.dev(v))
            removed[v] = t[n_keep:].cpu().numpy()
            self._dev[v] = t[:n_keep].clone()
        object.__setattr__(self, '_n', n_keep)
        return removed

    def permute(self, perm):
        for v in self.variables:
            self._dev[v] = self._engine.permute(perm, self.dev(v))

    def to_host_array(self, name):
        if name in self._host:
            return np.asarray(self._host[name])
        return self._dev[name].cpu().numpy()
