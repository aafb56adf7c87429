"""ORACLE (test infrastructure only): seeded inputs of the parity cases.

Everything is generated with numpy's frozen legacy RandomState streams, so the
same arrays are rebuilt on any machine; tests/golden/*.npz hold only the
reference's OUTPUTS for them (plus an input checksum).
"""
import numpy as np

from . import models as om


# ---------------------------------------------------------------------------
# Kernel cases (chisq / dwt / time_avg / bin_array / gelman_rubin)
# ---------------------------------------------------------------------------
def chisq_case(n=1000, seed=1):
    rs = np.random.RandomState(seed)
    data = rs.normal(5.0, 2.0, n)
    uncert = rs.uniform(0.5, 1.5, n)
    model = data + rs.normal(0.0, 1.0, n)*uncert
    params = np.array([1.0, 2.6, 0.3, 5.3, -0.2, 7.0])
    priors = np.array([0.0, 2.5, 0.0, 5.0, 0.0, 7.5])
    priorlow = np.array([0.0, 0.1, 0.0, 0.2, 0.0, 0.3])
    priorup = np.array([0.0, 0.1, 0.0, 0.4, 0.0, 0.6])
    return dict(model=model, data=data, uncert=uncert, params=params,
                priors=priors, priorlow=priorlow, priorup=priorup)


def dwt_case(n=1024, seed=2):
    rs = np.random.RandomState(seed)
    data = 1.0 + rs.normal(0.0, 2e-3, n)
    model = np.ones(n) - 0.01*(np.abs(np.linspace(-0.5, 0.5, n)) < 0.05)
    params = np.array([0.01, 0.0, 0.1, 1.0, 1.0, 5e-3, 1e-3])
    priors = np.array([0.0, 0.0, 0.0, 1.0, 0.0, 4e-3, 0.0])
    priorlow = np.array([0.0, 0.0, 0.0, 0.1, 0.0, 1e-3, 0.0])
    priorup = np.array([0.0, 0.0, 0.0, 0.2, 0.0, 2e-3, 0.0])
    return dict(model=model, data=data, params=params, priors=priors,
                priorlow=priorlow, priorup=priorup)


def teststats_series():
    """reference tests/test_stats.py:10-15 (legacy global stream, seed 12)."""
    rs = np.random.RandomState(12)
    N = 1000
    white = rs.normal(0, 5, N)
    red = np.sin(np.arange(N)/(0.1*N))*rs.normal(1.0, 1.0, N)
    return white, white + red


def series_case(n, seed, phi=0.95):
    """white N(0,1) + AR(1) red noise (phi, sigma 0.2), BASELINE config 4 form."""
    rs = np.random.RandomState(seed)
    white = rs.normal(0.0, 1.0, n)
    e = rs.normal(0.0, 0.2, n)
    red = np.empty(n)
    acc = 0.0
    for i in range(n):
        acc = phi*acc + e[i]
        red[i] = acc
    return white + red


def binarray_case(n=100003, seed=6):
    rs = np.random.RandomState(seed)
    return rs.normal(3.0, 1.0, n), np.abs(rs.normal(0.0, 1.0, n)) + 0.5


def gelman_case(seed=7, nchains=5, nfree=3):
    """Irregular zchain: initial rows (-1), interleaved chains of unequal length."""
    rs = np.random.RandomState(seed)
    counts = np.array([40, 37, 45, 41, 39])[:nchains]
    zchain = np.concatenate([-np.ones(10, int)] +
                            [np.full(c, k) for k, c in enumerate(counts)])
    zchain[10:] = rs.permutation(zchain[10:])
    Z = rs.normal(0.0, 1.0, (zchain.size, nfree)) \
        + 0.3*np.maximum(zchain, 0)[:, None]
    return Z, zchain, 5


# ---------------------------------------------------------------------------
# MCMC cases (inputs of tests/golden/mcmc_<case>_<sampler>.npz)
# ---------------------------------------------------------------------------
def mcmc_case(name):
    """Returns a dict of mcmc() inputs.  `model` names the built-in model."""
    z = np.zeros
    if name == 'quad':
        # examples/get_started.py:20-36
        np_state = np.random.RandomState(3)
        x = np.linspace(0, 10, 100)
        y = om.quad([3.0, -2.4, 0.5], x)
        uncert = np.sqrt(np.abs(y))
        data = y + np_state.normal(0, uncert)
        return dict(model='polynomial', x=x, data=data, uncert=uncert,
                    params=np.array([10.0, -2.0, 0.1]),
                    pstep=np.array([0.03, 0.03, 0.05]),
                    pmin=np.full(3, -np.inf), pmax=np.full(3, np.inf),
                    prior=z(3), priorlow=z(3), priorup=z(3),
                    nchains=7, nsamples=2100, thinning=1, burnin=50,
                    fepsilon=0.0, wlike=False)
    if name == 'sine':
        # small BASELINE-config-2 look-alike: bounds that get hit, a one- and a
        # two-sided Gaussian prior, live support draw (fepsilon > 0), thinning.
        rs = np.random.RandomState(20260102)
        n = 512
        x = np.linspace(0, 10, n)
        ptrue = np.array([1.0, 2.5, 0.3, 5.0, -0.2])
        data = om.sinusoid(ptrue, x) + rs.normal(0, 0.5, n)
        return dict(model='sinusoid', x=x, data=data, uncert=np.full(n, 0.5),
                    params=ptrue*1.01,
                    pstep=np.array([1e-2, 1e-3, 1e-2, 1e-2, 1e-3])*3,
                    pmin=np.array([0.0, 1.0, -np.pi, 0.0, -0.205]),
                    pmax=np.array([1.04, 5.0, np.pi, 10.0, 1.0]),
                    prior=np.array([0.0, 2.5, 0.0, 5.0, 0.0]),
                    priorlow=np.array([0.0, 0.1, 0.0, 0.2, 0.0]),
                    priorup=np.array([0.0, 0.1, 0.0, 0.4, 0.0]),
                    nchains=8, nsamples=2400, thinning=2, burnin=40,
                    fepsilon=0.01, wlike=False)
    if name == 'share':
        # quartic with one shared (p3 := p1) and one fixed (p4) parameter.
        rs = np.random.RandomState(11)
        n = 200
        x = np.linspace(-1, 1, n)
        ptrue = np.array([1.0, 0.5, -0.7, 0.5, 0.25])
        data = om.polynomial(ptrue, x) + rs.normal(0, 0.1, n)
        return dict(model='polynomial', x=x, data=data, uncert=np.full(n, 0.1),
                    params=np.array([1.1, 0.4, -0.6, 0.4, 0.25]),
                    pstep=np.array([0.02, 0.02, 0.03, -2.0, 0.0]),
                    pmin=np.full(5, -5.0), pmax=np.full(5, 5.0),
                    prior=z(5), priorlow=z(5), priorup=z(5),
                    nchains=6, nsamples=1200, thinning=1, burnin=20,
                    fepsilon=0.0, wlike=False)
    if name == 'wave':
        # box light curve with the wavelet likelihood (gamma fixed at 1).
        rs = np.random.RandomState(20260103)
        n = 256
        x = np.linspace(-0.5, 0.5, n)
        ptrue = np.array([0.01, 0.0, 0.1, 1.0])
        data = om.box(ptrue, x) + rs.normal(0, 1e-3, n)
        return dict(model='box', x=x, data=data, uncert=np.full(n, 1e-3),
                    params=np.array([0.0101, 0.001, 0.1003, 1.0, 1.0, 5e-4, 1e-3]),
                    pstep=np.array([2e-4, 1e-3, 1e-3, 1e-4, 0.0, 1e-4, 5e-5]),
                    pmin=np.array([0.0, -0.2, 0.01, 0.9, 0.0, 1e-5, 1e-4]),
                    pmax=np.array([0.05, 0.2, 0.3, 1.1, 2.0, 1e-2, 1e-2]),
                    prior=z(7), priorlow=z(7), priorup=z(7),
                    nchains=6, nsamples=900, thinning=1, burnin=20,
                    fepsilon=0.0, wlike=True)
    raise KeyError(name)


MCMC_CASES = ('quad', 'sine', 'share', 'wave')
SAMPLERS = ('snooker', 'demc', 'mrw')
PARENT_SEED, CHILD_SEED = 77, 4242


def checksum(*arrays):
    import hashlib
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a, dtype=float).tobytes())
    return h.hexdigest()[:16]


# ---------------------------------------------------------------------------
# Known answers printed in the reference's tests/test_stats.py (lines 17-86)
# ---------------------------------------------------------------------------
KAT_DAUB4_INV = np.array([
    -0.0301851821, -0.0522822690, -0.0662912607, -0.0824674511, -0.0905555462,
    -0.1008108399, -0.1132333322, -0.1250751254, 0.1325825215, 0.3180280110,
    0.4312613433, 0.5638438647, 0.1412513157, -0.1325825215, -0.2576576469,
    -0.4225925490, -0.1671021007, -0.0242642855, 0.0059208966, 0.0662912607,
    0.0140089918, -0.0080880952] + [0.0]*10)
KAT_DAUB4_FWD = np.array([
    0.1625300592, 0.0874699408, -0.0463140877, 0.2795672632, -0.0905555462,
    0.0, 0.0140089918, 0.1412513157, 0.3537658774, -0.0625, 0.0, 0.0, 0.0,
    0.0, 0.0, -0.1082531755, 0.0, 0.8365163037, -0.1294095226] + [0.0]*13)


KAT_RED_RMS = [5.20512494, 2.36785563, 1.72466452, 1.49355819, 1.52934937,
           1.35774105, 1.11881588, 1.13753563, 1.16566184, 1.03510878,
           1.11692786, 0.95551055, 1.04041202, 0.86876758, 0.93962365,
           0.95093077, 0.86283389, 0.89332354, 0.95500342, 0.82927083]
KAT_RED_RMSHI = [0.11639013, 0.12995296, 0.1285489, 0.13412548, 0.15774034,
             0.15574358, 0.1611256, 0.18169027, 0.20020244, 0.19264249,
             0.22147211, 0.20384028, 0.23076986, 0.2007309, 0.22759927,
             0.24306181, 0.23335404, 0.25645724, 0.29446565, 0.26262799]




def hpd_case(n=20000, seed=77):
    """Posterior sample [n, 4] for the HPD golden (tests/golden/hpd.npz): a Gaussian,
    a skewed (log-normal), a bimodal and a bounded (half-normal) marginal; 5
    parameters with one fixed and one shared entry in pstep."""
    rs = np.random.RandomState(seed)
    g = rs.normal(3.0, 0.2, n)
    ln = np.exp(rs.normal(0.0, 0.5, n))
    bi = np.where(rs.uniform(size=n) < 0.35, rs.normal(-1.0, 0.3, n), rs.normal(1.5, 0.5, n))
    hn = np.abs(rs.normal(0.0, 1.0, n))
    post = np.column_stack([g, ln, bi, hn])
    pstep = np.array([0.1, 0.1, 0.0, 0.1, -1.0, 0.1])
    bestp = np.array([3.0, 1.0, 7.0, 1.4, 3.0, 0.1])
    return post, bestp, pstep
