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


@pytest.mark.parametrize('nchains,n', [(4096, 100000), (4096, 100001), (1024, 100000), (128, 100000),
                                       (65536, 1000000), (8192, 10**7), (96, 5000), (7, 100),
                                       (4096, 127), (2048, 1 << 20)])
def test_split_boundaries_cover_the_series(nchains, n):
    """The (possibly decreasing-size) data splits of the model kernel partition
    [0, n): contiguous, whole tiles except the tail, sizes never growing."""
    import ctypes
    import numpy as np
    ns = ctypes.c_int(0)
    _lib.call('mc3b_model_chisq_plan', nchains, n, _lib.F64, ctypes.byref(ns))
    cap = ns.value + 1
    buf = (ctypes.c_int64*cap)()
    ns2 = ctypes.c_int(0)
    _lib.call('mc3b_model_chisq_splits', nchains, n, _lib.F64, buf, cap, ctypes.byref(ns2))
    assert ns2.value == ns.value
    b = np.array(buf[:cap])
    assert b[0] == 0 and b[-1] == n
    assert np.all(np.diff(b) >= 0)
    assert np.all(b[:-1] % 128 == 0)                 # fp64 tiles of 128 points
    if ns.value > 1:
        sizes = np.diff(b)[:-1]//128                 # all but the tail-carrying last split
        if sizes.size > 1 and n >= 128*ns.value:
            assert np.all(np.diff(sizes[:-1]) <= 1)  # equal (+-1) or decreasing
    with pytest.raises(_lib.Mc3bError, match='entries'):
        _lib.call('mc3b_model_chisq_splits', nchains, n, _lib.F64, buf, ns.value, ctypes.byref(ns2))


def test_large_population_gets_decreasing_splits():
    import ctypes
    import numpy as np
    buf = (ctypes.c_int64*512)()
    ns = ctypes.c_int(0)
    _lib.call('mc3b_model_chisq_splits', 4096, 100000, _lib.F64, buf, 512, ctypes.byref(ns))
    sizes = np.diff(np.array(buf[:ns.value + 1]))//128
    assert sizes[0] >= 3*sizes[-2] and sizes[:-1].min() >= 4     # factoring, at least 4 tiles
    assert sizes[:-1].sum() + sizes[-1] == 100000//128


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


def test_chisq_opts_layout_matches_header():
    text = open(os.path.join(ROOT, 'include', 'mc3b200.h')).read()
    body = text[text.index('typedef struct mc3b_chisq_opts {'):text.index('} mc3b_chisq_opts_t;')]
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    names = []
    for decl in body.split('{', 1)[1].split(';'):
        decl = decl.strip()
        if decl:
            for part in decl.split(','):
                names.append(re.findall(r'[A-Za-z_0-9]+', part)[-1])
    assert names == [f[0] for f in _lib.ChisqOpts._fields_]
    import ctypes
    assert ctypes.sizeof(_lib.ChisqOpts) == 104


def test_plan_is_a_function_of_plan_chains_only():
    """The split boundaries of a launch planned for the whole population do not
    depend on how many chains the launch holds (include/mc3b200.h, plan_chains)."""
    import ctypes
    import numpy as np
    out = []
    for _ in range(2):
        buf = (ctypes.c_int64*512)()
        ns = ctypes.c_int(0)
        _lib.call('mc3b_model_chisq_splits', 4096, 100000, _lib.F64, buf, 512, ctypes.byref(ns))
        out.append(np.array(buf[:ns.value + 1]))
    assert np.array_equal(out[0], out[1])


@pytest.mark.skipif(torch.cuda.is_available(), reason='CPU-box behaviour')
def test_no_cpu_fallback():
    import numpy as np
    import mc3_b200 as mc3
    with pytest.raises(_lib.Mc3bError):
        mc3.stats.chisq(np.ones(4), np.ones(4), np.ones(4))
