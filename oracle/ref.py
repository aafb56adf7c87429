"""ORACLE (test infrastructure only): loaders for the real reference.

* ref_ext()   -- the reference's own C extension modules compiled by
                 oracle/Makefile into oracle/_ref/ (these travel to the GPU box).
* ref_py()    -- the reference's Python hot path (mc3.mcmc_driver, mc3.chain,
                 mc3.stats, mc3.utils) imported from /root/reference through a
                 stub package that skips mc3/__init__.py (it needs matplotlib,
                 which is not installed).  Authoring container ONLY: nothing that
                 runs on the GPU box may call this.
"""
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = '/root/reference'
REF_DIR = os.path.join(_HERE, '_ref')


def have_ref_ext():
    return os.path.isdir(REF_DIR) and any(
        f.startswith('_chisq') for f in os.listdir(REF_DIR))


def ref_ext():
    """Return (_chisq, _dwt, _time_averaging, _binarray) from oracle/_ref."""
    if not have_ref_ext():
        raise RuntimeError('oracle/_ref not built (run make -C oracle ref)')
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    return tuple(importlib.import_module(m) for m in
                 ('_chisq', '_dwt', '_time_averaging', '_binarray'))


def have_ref_py():
    return os.path.isdir(os.path.join(REF_ROOT, 'mc3'))


def ref_py():
    """Import the reference's Python hot path under the name `mc3` via a stub."""
    if not have_ref_py():
        raise RuntimeError('/root/reference is not present on this machine')
    ref_ext()
    if 'mc3' not in sys.modules or not getattr(sys.modules['mc3'], '_orc_stub', False):
        stub = types.ModuleType('mc3')
        stub.__path__ = [os.path.join(REF_ROOT, 'mc3')]
        stub._orc_stub = True
        sys.modules['mc3'] = stub
        ver = importlib.import_module('mc3.version')
        stub.__version__ = ver.__version__
    mu = importlib.import_module('mc3.utils')
    ms = importlib.import_module('mc3.stats')
    drv = importlib.import_module('mc3.mcmc_driver')
    ch = importlib.import_module('mc3.chain')
    return types.SimpleNamespace(utils=mu, stats=ms, mcmc_driver=drv, chain=ch)
