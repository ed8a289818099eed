#!/usr/bin/env python
"""Double gyre, advection schemes -- the reference's examples/example_double_gyre_advection_schemes.py (BASELINE
configs[0]) on the drop-in classes: same reader constructor, config keys, seed_elements / run calls; only the two import
lines differ and the plot at the end is replaced by a table (plotting is outside this package's scope).

Needs a B200 and the built library (python -c "import __graft_entry__ as g; g.build()")."""
from datetime import timedelta

from opendrift_b200.readers import reader_double_gyre            # reference: from opendrift.readers import reader_double_gyre
from opendrift_b200.models.oceandrift import OceanDrift          # reference: from opendrift.models.oceandrift import OceanDrift

double_gyre = reader_double_gyre.Reader(epsilon=.25, omega=0.628, A=0.25)
duration = timedelta(seconds=6)
x = [.6]
y = [.3]
lon, lat = double_gyre.xy2lonlat(x, y)

runs = []
leg = []
for scheme in ['euler', 'runge-kutta', 'runge-kutta4']:
    for time_step in [0.01, 0.1]:
        leg.append(scheme + ', T=%.2fs' % time_step)
        o = OceanDrift(loglevel=50)
        o.set_config('environment:fallback:land_binary_mask', 0)
        o.set_config('drift:advection_scheme', scheme)
        o.add_reader(double_gyre)
        o.seed_elements(lon, lat, time=double_gyre.initial_time)
        o.run(duration=duration, time_step=time_step)
        runs.append(o)

print('%-26s %12s %12s' % ('scheme, time step', 'x [m]', 'y [m]'))
for name, o in zip(leg, runs):
    fx, fy = double_gyre.lonlat2xy(o.elements.lon, o.elements.lat)
    print('%-26s %12.6f %12.6f' % (name, fx[0], fy[0]))
