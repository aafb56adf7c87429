"""ORACLE (test infrastructure only): numpy wrappers over liboracle.so.

Function names, argument order and return conventions follow the reference's
Python wrappers so tests read like the reference's own:
  chisq / residuals / dwt_chisq / dwt_daub4 / bin_array   mc3/stats/stats.py:36-284, 577-611
  time_avg                                                mc3/stats/time_averaging.py:17-60
  log_prior                                               mc3/stats/stats.py:287-392
  gelman_rubin                                            mc3/stats/gelman.py:12-92
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_dp = ctypes.POINTER(ctypes.c_double)


def build():
    """Compile liboracle.so (and oracle/_ref when the reference tree is here)."""
    targets = ['oracle']
    if os.path.isdir('/root/reference/src_c'):
        targets.append('ref')
    subprocess.run(['make', '-s', '-C', _HERE] + targets, check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, 'liboracle.so')
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        L.orc_priors.restype = ctypes.c_double
        L.orc_chisq.restype = ctypes.c_double
        L.orc_dwt_chisq.restype = ctypes.c_double
        L.orc_residuals.restype = None
        L.orc_daub4.restype = ctypes.c_long
        L.orc_binrms.restype = ctypes.c_long
        L.orc_binarray.restype = ctypes.c_long
        L.orc_invgamma.restype = None
        _LIB = L
    return _LIB


def _d(a):
    return np.ascontiguousarray(a, dtype=np.double)


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _prior_args(params, priors, priorlow, priorup):
    """stats.py:208-216 -- mask Gaussian priors (low>0 & up>0)."""
    if params is None or priors is None or priorlow is None or priorup is None:
        return None, None, None, 0
    params, priors = _d(params), _d(priors)
    priorlow, priorup = _d(priorlow), _d(priorup)
    ip = (priorlow > 0) & (priorup > 0)
    off = _d((params - priors)[ip])
    return off, _d(priorlow[ip]), _d(priorup[ip]), int(off.size)


def chisq(model, data, uncert, params=None, priors=None, priorlow=None,
          priorup=None):
    model, data, uncert = _d(model), _d(data), _d(uncert)
    off, lo, up, npr = _prior_args(params, priors, priorlow, priorup)
    return lib().orc_chisq(
        _p(model), _p(data), _p(uncert), ctypes.c_long(model.size),
        _p(off), _p(lo), _p(up), ctypes.c_long(npr))


def residuals(model, data, uncert, params=None, priors=None, priorlow=None,
              priorup=None):
    model, data, uncert = _d(model), _d(data), _d(uncert)
    off, lo, up, npr = _prior_args(params, priors, priorlow, priorup)
    out = np.empty(model.size + npr)
    lib().orc_residuals(
        _p(model), _p(data), _p(uncert), ctypes.c_long(model.size),
        _p(off), _p(lo), _p(up), ctypes.c_long(npr), _p(out))
    return out


def dwt_chisq(model, data, params, priors=None, priorlow=None, priorup=None):
    if len(params) < 3:
        raise ValueError('Wavelet chisq should have at least three parameters')
    model, data, params = _d(model), _d(data), _d(params)
    off, lo, up, npr = _prior_args(params, priors, priorlow, priorup)
    return lib().orc_dwt_chisq(
        _p(params), ctypes.c_long(params.size), _p(model), _p(data),
        ctypes.c_long(data.size), _p(off), _p(lo), _p(up), ctypes.c_long(npr))


def dwt_daub4(array, inverse=False):
    array = _d(array)
    n = array.size
    out = np.empty(1 << int(np.ceil(np.log2(n))))
    lib().orc_daub4(_p(array), ctypes.c_long(n),
                    ctypes.c_int(-1 if inverse else 1), _p(out))
    return out


def time_avg(data, maxbins=None, binstep=1):
    data = _d(data)
    if maxbins is None:
        maxbins = data.size // 2
    maxbins, binstep = int(maxbins), int(binstep)
    nout = (maxbins - 1)//binstep + 1
    outs = [np.empty(nout) for _ in range(5)]
    lib().orc_binrms(_p(data), ctypes.c_long(data.size),
                     ctypes.c_long(maxbins), ctypes.c_long(binstep),
                     *[_p(o) for o in outs])
    return outs


def invgamma(M, s, ds):
    lo, hi = ctypes.c_double(), ctypes.c_double()
    lib().orc_invgamma(ctypes.c_int(M), ctypes.c_double(s),
                       ctypes.c_double(ds), ctypes.byref(lo), ctypes.byref(hi))
    return lo.value, hi.value


def bin_array(data, binsize, uncert=None):
    data = _d(data)
    binsize = int(binsize)
    nb = data.size // binsize
    bd = np.empty(nb)
    if uncert is None:
        lib().orc_binarray(_p(data), ctypes.c_long(data.size),
                           ctypes.c_long(binsize), None, _p(bd), None)
        return bd
    uncert = _d(uncert)
    bs = np.empty(nb)
    lib().orc_binarray(_p(data), ctypes.c_long(data.size),
                       ctypes.c_long(binsize), _p(uncert), _p(bd), _p(bs))
    return [bd, bs]


def log_prior(posterior, prior, priorlow, priorup, pstep):
    """stats.py:367-392 restated: -0.5*sum of squared, width-scaled offsets of
    the free parameters with Gaussian priors; 2*log(p) squared-term for
    priorlow<0 (kept as the reference computes it)."""
    post = np.atleast_2d(np.asarray(posterior, dtype=float))
    ifree = np.where(pstep > 0)[0]
    terms = np.zeros_like(post)
    for i, k in enumerate(ifree):
        if priorlow[k] > 0 and priorup[k] > 0:
            d = post[:, i] - prior[k]
            terms[:, i] = np.where(d < 0, d/priorlow[k],
                                   np.where(d > 0, d/priorup[k], d))
        elif priorlow[k] < 0:
            terms[:, i] = 2.0*np.log(post[:, i])
    logp = -0.5*np.sum(terms**2, axis=1)
    return logp[0] if logp.size == 1 else logp


def gelman_rubin(Z, zchain, burnin):
    """gelman.py:36-92 restated: PSRF per parameter from the first
    min-over-chains(count-burnin) post-burn samples of every chain."""
    nchains = int(np.amax(zchain)) + 1
    npars = Z.shape[1]
    counts = np.array([np.sum(zchain == c) for c in range(nchains)]) - burnin
    niter = int(np.amin(counts))
    if niter < 1:
        return np.zeros(npars)
    x = np.empty((nchains, niter, npars))
    for c in range(nchains):
        x[c] = Z[np.where(zchain == c)[0][burnin:burnin + niter]]
    W = np.mean(np.var(x, axis=1), axis=0)
    mu = np.mean(x, axis=1)
    B = niter/(nchains - 1.0)*np.sum((mu - np.mean(mu, axis=0))**2, axis=0)
    V = W*(niter - 1.0)/niter + B*(nchains + 1.0)/(niter*nchains)
    return np.sqrt(V/W)
