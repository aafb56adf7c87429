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


def inverse_daub4(coef):
    """Inverse Daubechies-4 pyramid of a length-2^M coefficient vector laid out as
    the reference's transform leaves it (wavelet.h:109-128: two smooth coefficients,
    then 2^m details of scale m at [2^m, 2^(m+1))).  Vectorised numpy, used only to
    synthesise red noise for config 3; the product transform is
    mc3b_daub4."""
    c0, c1 = (1 + np.sqrt(3))/(4*np.sqrt(2)), (3 + np.sqrt(3))/(4*np.sqrt(2))
    c2, c3 = (3 - np.sqrt(3))/(4*np.sqrt(2)), (1 - np.sqrt(3))/(4*np.sqrt(2))
    a = np.array(coef, dtype=float)
    n = a.size
    nn = 4
    while nn <= n:
        nh = nn//2
        s, d = a[:nh].copy(), a[nh:nn].copy()
        sp, dp = np.roll(s, 1), np.roll(d, 1)            # element i-1, periodic
        a[0:nn:2] = c2*sp + c1*dp + c0*s + c3*d
        a[1:nn:2] = c3*sp - c0*dp + c1*s - c2*d
        nn *= 2
    return a


def red_noise(n, sigma_r, gamma=1.0, rs=None):
    """1/f^gamma noise of Carter & Winn (2009) as the wavelet likelihood models it:
    independent wavelet coefficients with variance sigma_r^2 2^(-gamma m) at scale m
    and sigma_r^2 2^(-gamma) g(gamma) for the two scaling coefficients (_dwt.c:96-110),
    taken back to the time domain."""
    rs = rs or np.random.RandomState(0)
    M = int(np.log2(n))
    assert 1 << M == n
    g = 0.72134752 if gamma == 1.0 else 1.0/(2.0**(1.0 - gamma) - 1.0)
    c = np.empty(n)
    c[0:2] = rs.normal(0, sigma_r*np.sqrt(2.0**(-gamma)*g), 2)
    for m in range(1, M):
        c[1 << m:1 << (m + 1)] = rs.normal(0, sigma_r*2.0**(-0.5*gamma*m), 1 << m)
    return inverse_daub4(c)


def config3(n=1 << 20, seed=20260103):
    """snooker, transit-like box, N=2^20, wavelet likelihood: white noise
    sigma_w = 1e-3 plus red noise sigma_r = 5e-3 synthesised by the inverse D4
    transform."""
    rs = np.random.RandomState(seed)
    x = np.linspace(-0.5, 0.5, n)
    ptrue = np.array([0.01, 0.0, 0.1, 1.0])
    y = ptrue[3] - ptrue[0]*(np.abs(x - ptrue[1]) < 0.5*ptrue[2])
    data = y + rs.normal(0, 1e-3, n) + red_noise(n, 5e-3, 1.0, rs)
    return dict(
        name='config3: snooker, box light curve N=2^20, wlike, white + red (inverse D4) noise',
        model='box', x=x, data=data, uncert=np.full(n, 1e-3),
        params=np.array([0.0101, 0.001, 0.1003, 1.0, 1.0, 5e-3, 1e-3]),
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
