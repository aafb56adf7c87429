"""ORACLE (test infrastructure only): loaders for the real reference.

* ref_ext()   -- the reference's own C extension modules compiled by
                 oracle/Makefile into oracle/_ref/ (these travel to the GPU box).
* ref_py()    -- the reference's Python hot path (mc3.mcmc_driver, mc3.chain,
                 mc3.stats, mc3.utils) imported from /root/reference through a
                 stub package that skips mc3/__init__.py (it needs matplotlib,
                 which is not installed) -- or, on the GPU box, from the byte-
                 compiled copies (.bc) `make -C oracle refpy` builds into oracle/_ref/mc3
                 (used by bench.py's reference arm only).
"""
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get('ORACLE_REF_ROOT', '/root/reference')
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


def _ref_py_dir():
    """The reference's package directory: its sources in the authoring container,
    else the byte-compiled modules `make -C oracle refpy` left in oracle/_ref/mc3
    (build outputs, which travel to the GPU box)."""
    src = os.path.join(REF_ROOT, 'mc3')
    if os.path.isdir(src):
        return src
    bc = os.path.join(REF_DIR, 'mc3')
    if os.path.exists(os.path.join(bc, 'mcmc_driver.bc')):
        return bc
    return None


class _ByteCodeFinder:
    """Import mc3.* from the byte-compiled files `make -C oracle refpy` wrote
    (<module>.bc, packages as <package>/__init__.bc)."""

    def __init__(self, root):
        self.root = root

    def find_spec(self, fullname, path=None, target=None):
        import importlib.machinery as mach
        import importlib.util as iu
        if not fullname.startswith('mc3.'):
            return None
        rel = fullname.split('.')[1:]
        mod = os.path.join(self.root, *rel) + '.bc'
        pkg = os.path.join(self.root, *rel, '__init__.bc')
        if os.path.exists(pkg):
            return iu.spec_from_file_location(
                fullname, pkg, loader=mach.SourcelessFileLoader(fullname, pkg),
                submodule_search_locations=[os.path.dirname(pkg)])
        if os.path.exists(mod):
            return iu.spec_from_file_location(
                fullname, mod, loader=mach.SourcelessFileLoader(fullname, mod))
        return None


def have_ref_py():
    return _ref_py_dir() is not None


def ref_py():
    """Import the reference's Python hot path under the name `mc3` via a stub."""
    if not have_ref_py():
        raise RuntimeError('neither /root/reference nor oracle/_ref/mc3 is present on this machine')
    ref_ext()
    if 'mc3' not in sys.modules or not getattr(sys.modules['mc3'], '_orc_stub', False):
        stub = types.ModuleType('mc3')
        root = _ref_py_dir()
        stub.__path__ = [root]
        if root.startswith(REF_DIR) and not any(isinstance(f, _ByteCodeFinder) for f in sys.meta_path):
            sys.meta_path.insert(0, _ByteCodeFinder(root))
        stub._orc_stub = True
        sys.modules['mc3'] = stub
        ver = importlib.import_module('mc3.version')
        stub.__version__ = ver.__version__
    mu = importlib.import_module('mc3.utils')
    ms = importlib.import_module('mc3.stats')
    drv = importlib.import_module('mc3.mcmc_driver')
    ch = importlib.import_module('mc3.chain')
    return types.SimpleNamespace(utils=mu, stats=ms, mcmc_driver=drv, chain=ch)
