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
    assert C.sizeof(_lib.GroupDesc) == 32 + 64 + 8
    assert C.sizeof(_lib.AdvectArgs) == 8 + 3 * 24 + 16 + 4 * 8 + 8 + 3 * 8 + 8 + 2 * 8 + 8 + 16
    assert C.sizeof(_lib.StepArgs) == C.sizeof(_lib.AdvectArgs) + 8 + 24 + 16 + 8 + 24 + 8 + 24 + 8 + 8
    assert C.sizeof(_lib.MixArgs) == 8 + 24 + 8 + 9 * 8 + 2 * 8 + 8 + 8 * 4 + 2 * 8 + 3 * 8


def test_no_cpu_fallback():
    import torch
    from opendrift_b200.engine import Engine
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(RuntimeError):
        Engine(0)
