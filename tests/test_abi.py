"""The C-ABI library loads and exports every symbol include/odcuda.h declares (no GPU needed: nothing
is computed), and the product refuses to run without it."""
import ctypes as C
import os
import re

import pytest

import common
from opendrift_b200 import _lib, build


def _declared():
    text = open(os.path.join(common.ROOT, 'include', 'odcuda.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(od_[a-z_0-9]+)\s*\(', text)))


def test_library_builds_and_exports_every_declared_symbol():
    build.build()
    lib = C.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(_lib.SYMBOLS)
    lib.od_abi_version.restype = C.c_int
    assert lib.od_abi_version() == 1


def test_struct_layouts_match_header():
    # sizes the header implies (64-bit): guards against ctypes / C drift
    assert C.sizeof(_lib.TimeSample) == 24
    assert C.sizeof(_lib.HostIO) == 5 * 8 + 8 + 8
    assert C.sizeof(_lib.GroupDesc) == 32 + 64 + 8 + (8 + 10 * 8) + 8
    assert C.sizeof(_lib.AdvectArgs) == 8 + 3 * 24 + 16 + 4 * 8 + 8 + 3 * 8 + 8 + 2 * 8 + 8 + 16 + 16 + 6 * 24
    assert C.sizeof(_lib.StepArgs) == C.sizeof(_lib.AdvectArgs) + 8 + 24 + 16 + 8 + 24 + 8 + 24 + 8 + 8
    assert C.sizeof(_lib.MixArgs) == 8 + 24 + 8 + 9 * 8 + 2 * 8 + 8 + 8 * 4 + 2 * 8 + 3 * 8 + 2 * 8 + 8 + 8 + 8


STRUCTS = {'od_time_sample': 'TimeSample', 'od_group_desc': 'GroupDesc', 'od_host_io': 'HostIO', 'od_advect_args': 'AdvectArgs',
           'od_step_args': 'StepArgs', 'od_mix_args': 'MixArgs', 'od_leeway_args': 'LeewayArgs', 'od_stokes_args': 'StokesArgs',
           'od_proj_desc': 'ProjDesc', 'od_analytic_desc': 'AnalyticDesc', 'od_analytic_advect_args': 'AnalyticAdvectArgs',
           'od_history_args': 'HistoryArgs', 'od_buoyancy_args': 'BuoyancyArgs', 'od_bookkeep_args': 'BookkeepArgs', 'od_pack_args': 'PackArgs',
           'od_coast_args': 'CoastArgs'}


def test_every_struct_field_has_the_offset_the_c_compiler_gives_it(tmp_path):
    """gcc's sizeof / offsetof of every field of every struct in include/odcuda.h against the ctypes mirror (same field
    names): a renamed, re-ordered, re-typed or missing field fails here, without a GPU."""
    import subprocess
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "odcuda.h"', 'int main(void) {']
    expect = []
    for cname, pyname in STRUCTS.items():
        st = getattr(_lib, pyname)
        lines.append('  printf("%%zu\\n", sizeof(%s));' % cname)
        expect.append(('sizeof ' + cname, C.sizeof(st)))
        for fname, _ in st._fields_:
            lines.append('  printf("%%zu\\n", offsetof(%s, %s));' % (cname, fname))
            expect.append(('%s.%s' % (cname, fname), getattr(st, fname).offset))
    lines += ['  return 0;', '}']
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'layout'
    subprocess.check_call(['gcc', '-I', os.path.join(common.ROOT, 'include'), '-o', str(exe), str(src)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert len(got) == len(expect)
    bad = [(name, e, g) for (name, e), g in zip(expect, got) if e != g]
    assert not bad, bad
    # every struct the header defines is mirrored
    text = open(os.path.join(common.ROOT, 'include', 'odcuda.h')).read()
    assert set(re.findall(r'}\s*(od_[a-z_0-9]+)\s*;', text)) == set(STRUCTS)


def test_no_cpu_fallback():
    import torch
    from opendrift_b200.engine import Engine
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(RuntimeError):
        Engine(0)
