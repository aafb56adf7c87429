"""Synthetic inputs of the BASELINE.json configurations, as plain
numpy (no device work).  Used by bench.py and the full-size tests."""
import numpy as np


def sinusoid_np(p, x):
    return p[0]*np.sin(2.0*np.pi*x/p[1] + p[2]) + p[3] + p[4]*x


def config2(n=100_000, seed=20260102):
    """DEMC, 5-parameter sinusoid+line, N=1e5, Gaussian priors (one two-sided)."""
    rs = np.random.RandomState(seed)
    x = np.linspace(0, 10, n)
    ptrue = np.array([1.0, 2.5, 0.3, 5.0, -0.2])
    data = sinusoid_np(ptrue, x) + rs.normal(0, 0.5, n)
    return dict(
        name='config2: DEMC, 5-param sinusoid+line, N=1e5, fp64 chisq + Gaussian priors',
        model='sinusoid', x=x, data=data, uncert=np.full(n, 0.5),
        params=ptrue*1.01, pstep=np.array([1e-2, 1e-3, 1e-2, 1e-2, 1e-3]),
        pmin=np.array([0.0, 1.0, -np.pi, 0.0, -1.0]),
        pmax=np.array([5.0, 5.0, np.pi, 10.0, 1.0]),
        prior=np.array([0.0, 2.5, 0.0, 5.0, 0.0]),
        priorlow=np.array([0.0, 0.1, 0.0, 0.2, 0.0]),
        priorup=np.array([0.0, 0.1, 0.0, 0.4, 0.0]),
        sampler='demc', fepsilon=0.01, thinning=1, wlike=False,
        flops_per_point=10)        # model 7 + residual/square 3


def config3(n=1 << 20, seed=20260103):
    """snooker, transit-like box, N=2^20, wavelet likelihood."""
    rs = np.random.RandomState(seed)
    x = np.linspace(-0.5, 0.5, n)
    ptrue = np.array([0.01, 0.0, 0.1, 1.0])
    y = ptrue[3] - ptrue[0]*(np.abs(x - ptrue[1]) < 0.5*ptrue[2])
    data = y + rs.normal(0, 1e-3, n)
    return dict(
        name='config3: snooker, box light curve N=2^20, wlike',
        model='box', x=x, data=data, uncert=np.full(n, 1e-3),
        params=np.array([0.0101, 0.001, 0.1003, 1.0, 1.0, 5e-4, 1e-3]),
        pstep=np.array([2e-4, 1e-3, 1e-3, 1e-4, 0.0, 1e-4, 5e-5]),
        pmin=np.array([0.0, -0.2, 0.01, 0.9, 0.0, 1e-5, 1e-4]),
        pmax=np.array([0.05, 0.2, 0.3, 1.1, 2.0, 1e-2, 1e-2]),
        prior=np.zeros(7), priorlow=np.zeros(7), priorup=np.zeros(7),
        sampler='snooker', fepsilon=0.0, thinning=1, wlike=True,
        flops_per_point=21)


def config4(n=100_000_000, seed=20260104):
    """white + AR(1) red noise series for time_avg / bin_array."""
    rs = np.random.RandomState(seed)
    white = rs.normal(0.0, 1.0, n)
    e = rs.normal(0.0, 0.2, n)
    # AR(1) via a blocked recursion (vectorised): r_i = phi r_{i-1} + e_i
    from scipy.signal import lfilter
    red = lfilter([1.0], [1.0, -0.95], e)
    return white + red
