"""CPU-only: the C-ABI library loads and exports every symbol the header
declares; the ctypes table covers exactly the header; the product path fails
loudly without a GPU (no CPU fallback)."""
import os
import re

import pytest
import torch

from mc3_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, 'include', 'mc3b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(mc3b_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_functions():
    syms = header_symbols()
    assert len(syms) >= 20
    assert 'mc3b_model_chisq' in syms and 'mc3b_metropolis' in syms


def test_library_exports_every_declared_symbol():
    import ctypes
    assert os.path.exists(_lib.LIBPATH), 'build with python -m mc3_b200.build'
    lib = ctypes.CDLL(_lib.LIBPATH)
    for name in header_symbols():
        assert hasattr(lib, name), f'{name} declared in mc3b200.h but not exported'


def test_ctypes_table_matches_header():
    assert sorted(_lib.EXPORTS) == header_symbols()


def test_version_and_plan_without_device():
    import ctypes
    lib = _lib.load()
    assert lib.mc3b_version() == 100
    ns = ctypes.c_int(0)
    _lib.call('mc3b_model_chisq_plan', 4096, 100000, _lib.F64, ctypes.byref(ns))
    assert 1 <= ns.value <= 4096
    ns2 = ctypes.c_int(0)
    _lib.call('mc3b_model_chisq_plan', 4096, 100000, _lib.F64, ctypes.byref(ns2))
    assert ns.value == ns2.value            # deterministic shape policy


def test_bad_arguments_raise():
    import ctypes
    ns = ctypes.c_int(0)
    with pytest.raises(_lib.Mc3bError, match='bad'):
        _lib.call('mc3b_model_chisq_plan', 0, 10, _lib.F64, ctypes.byref(ns))
    with pytest.raises(_lib.Mc3bError, match='2\\^k'):
        _lib.call('mc3b_daub4', 8, 12, 1, 8, 16, None)


def test_sampler_struct_layout_matches_header():
    """Field order/offsets of the ctypes mirror follow the C struct."""
    text = open(os.path.join(ROOT, 'include', 'mc3b200.h')).read()
    body = text[text.index('typedef struct mc3b_sampler {'):text.index('} mc3b_sampler_t;')]
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    names = []
    for decl in body.split('{', 1)[1].split(';'):
        decl = decl.strip()
        if decl:
            for part in decl.split(','):
                names.append(re.findall(r'[A-Za-z_0-9]+', part)[-1])
    assert names == [f[0] for f in _lib.SamplerStruct._fields_]


@pytest.mark.skipif(torch.cuda.is_available(), reason='CPU-box behaviour')
def test_no_cpu_fallback():
    import numpy as np
    import mc3_b200 as mc3
    with pytest.raises(_lib.Mc3bError):
        mc3.stats.chisq(np.ones(4), np.ones(4), np.ones(4))
