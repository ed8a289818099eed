"""bench.py's reference arm (`--impl reference`: the reference's CPU path, here the bit-identical port) runs without a GPU and
prints ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

import common


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(common.ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
                          '--cpu-particles', '5000'], capture_output=True, text=True, timeout=600, cwd=common.ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
              'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['higher_is_better'] is True and d['value'] > 0 and d['steps'] == 1
    assert d['unit'] == 'particle-steps/s' and 'workload' in d['config']
    assert d['cpu_baseline']['kind'] in ('port', 'reference') and d['cpu_baseline']['cores'] >= 1
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0 and d['e2e']['value'] == d['value']
