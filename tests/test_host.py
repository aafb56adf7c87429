"""CPU-only host logic of the drop-in API: argument validation and the
reference's error strings (reference tests/test_mcmc.py:328-409), burn() golden
masks (reference tests/test_utils.py:115-156), names, logger."""
import re

import numpy as np
import pytest

import mc3_b200 as mc3
from mc3_b200 import utils as mu


def quad(p, x):
    return p[0] + p[1]*x + p[2]*x**2.0


np.random.seed(12)
x = np.linspace(0, 10, 100)
y = quad([4.5, -2.4, 0.5], x)
uncert = np.sqrt(np.abs(y))
data = y + np.random.normal(0, uncert)
params = np.array([10.0, -2.0, 0.1])
pstep = np.array([0.03, 0.03, 0.05])
QUIET = dict(log=mu.Log(verb=-1))


@pytest.mark.parametrize('drop,msg', [
    ('data', "'data' is a required argument"),
    ('uncert', "'uncert' is a required argument"),
    ('func', "'func' must be either a callable or an iterable"),
    ('params', "'params' is a required argument"),
    ('sampler', "'sampler' is a required argument"),
    ('nsamples', "'nsamples' is a required argument for MCMC runs"),
])
def test_required_arguments(drop, msg):
    kw = dict(data=data, uncert=uncert, func=quad, params=np.copy(params),
              sampler='snooker', indparams=[x], pstep=pstep, nsamples=1e4,
              burnin=100, **QUIET)
    kw.pop(drop)
    with pytest.raises(ValueError, match=msg):
        mc3.sample(**kw)


def test_burnin_larger_than_chain():
    msg = re.escape('The number of burned-in samples (2000) is greater than the '
                    'number of iterations per chain (1429)')
    with pytest.raises(ValueError, match=msg):
        mc3.sample(data, uncert, func=quad, params=np.copy(params),
                   sampler='snooker', indparams=[x], pstep=pstep, nsamples=1e4,
                   burnin=2000, **QUIET)


def test_leastsq_error():
    msg = re.escape("Invalid 'leastsq' input (invalid). Must select from ['lm', 'trf']")
    with pytest.raises(ValueError, match=msg):
        mc3.sample(data, uncert, func=quad, params=np.copy(params),
                   sampler='snooker', indparams=[x], pstep=pstep,
                   leastsq='invalid', nsamples=1e4, burnin=100, **QUIET)


def test_out_of_bounds_initial_guess():
    with pytest.raises(ValueError, match='Some initial-guess values are out of bounds'):
        mc3.sample(data, uncert, func=quad, params=np.copy(params),
                   sampler='snooker', indparams=[x], pstep=pstep, nsamples=1e4,
                   pmin=[11.0, -5, -5], pmax=[20.0, 5, 5], **QUIET)


def test_func_output_size_mismatch():
    with pytest.raises(ValueError, match='does not match the size of the func'):
        mc3.sample(data, uncert, func=quad, params=np.copy(params),
                   sampler='snooker', indparams=[x[:50]], pstep=pstep,
                   nsamples=1e4, **QUIET)


def test_uncert_not_mutated_by_validation():
    u0 = np.copy(uncert)
    with pytest.raises(ValueError):
        mc3.sample(data, uncert, func=quad, params=np.copy(params),
                   sampler='snooker', indparams=[x], pstep=pstep, nsamples=1e4,
                   burnin=2000, **QUIET)
    assert np.array_equal(uncert, u0)


def test_burn_golden_masks():
    Z = np.expand_dims([0., 1, 10, 20, 30, 11, 31, 21, 12, 22, 32], axis=1)
    zchain = np.array([-1, -1, 0, 1, 2, 0, 2, 1, 0, 1, 2])
    zd = {'posterior': Z, 'zchain': zchain, 'burnin': 1}
    post, zc, zm = mu.burn(zd)
    np.testing.assert_equal(post[:, 0], [11., 12., 21., 22., 31., 32.])
    np.testing.assert_equal(zc, [0, 0, 1, 1, 2, 2])
    np.testing.assert_equal(zm, [5, 8, 7, 9, 6, 10])
    post, zc, zm = mu.burn(zd, sort=False)
    np.testing.assert_equal(post[:, 0], [11., 31., 21., 12., 22., 32.])
    post, zc, zm = mu.burn(zd, burnin=0)
    np.testing.assert_equal(post[:, 0], [10., 11., 12., 20., 21., 22., 30., 31., 32.])
    post, zc, zm = mu.burn(Z=Z, zchain=zchain, burnin=1)
    np.testing.assert_equal(post[:, 0], [11., 12., 21., 22., 31., 32.])
    with pytest.raises(ValueError, match='Need to input either Zdict'):
        mu.burn(Z=Z, zchain=zchain)


def test_default_parnames_and_log_error():
    np.testing.assert_equal(mu.default_parnames(3), ['Param 1', 'Param 2', 'Param 3'])
    with pytest.raises(ValueError, match='boom'):
        mu.Log(verb=-1).error('boom')


def test_sample_statistics_expand_fixed_and_shared():
    from mc3_b200 import stats as ms
    rs = np.random.RandomState(5)
    post = rs.normal([1.0, 2.0], [0.1, 0.2], (4000, 2))
    bestp = np.array([1.0, 2.0, 2.0, 7.0])
    pstep = np.array([0.1, 0.1, -2.0, 0.0])
    med, mean, std, lo, hi = ms.calc_sample_statistics(post, bestp, pstep)
    assert med[2] == med[1] and std[2] == std[1]          # shared copies its source
    assert std[3] == 0.0 and med[3] == 7.0                # fixed keeps bestp
    np.testing.assert_allclose(mean[:2], [1.0, 2.0], atol=0.02)
    np.testing.assert_allclose(hi[:2] - lo[:2], [0.2, 0.4], rtol=0.1)


def test_log_prior_doc_values():
    from mc3_b200 import stats as ms
    post = np.array([[3.0, 2.0], [3.1, 1.0], [3.6, 1.5]])
    lp = ms.log_prior(post, np.array([3.5, 0.0]), np.array([0.1, 0.0]),
                      np.array([0.1, 0.0]), np.array([1.0, 1.0]))
    np.testing.assert_allclose(lp, [-12.5, -8.0, -0.5])


def test_config3_red_noise_generator_matches_oracle_transform():
    """workloads.inverse_daub4 (used to synthesise config 3's red noise)
    is the reference's inverse D4 pyramid; the noise has the 1/f scale variances."""
    import numpy as np
    from mc3_b200 import workloads
    from oracle import kernels as ok
    rs = np.random.RandomState(5)
    for n in (4, 8, 64, 1024):
        c = rs.normal(0, 1, n)
        np.testing.assert_allclose(workloads.inverse_daub4(c), ok.dwt_daub4(c, True), rtol=0, atol=1e-15)
    r = workloads.red_noise(1 << 14, 5e-3, 1.0, np.random.RandomState(1))
    coef = ok.dwt_daub4(r)
    for m in (6, 9, 12):                              # scale variances sigma_r^2 2^-m
        got = np.var(coef[1 << m:1 << (m + 1)])
        assert abs(got/(5e-3**2*2.0**-m) - 1) < 0.35
    w = workloads.config3(n=1 << 12)
    assert w['data'].size == 1 << 12 and w['params'][5] == 5e-3


def test_sample_statistics_host_path_matches_reference_golden():
    """calc_sample_statistics (host numpy/scipy path) against the reference's own
    output on the seeded posterior of problems.hpd_case (tests/golden/hpd.npz,
    written by oracle/make_golden.py from mc3/stats/stats.py:876-964)."""
    import os
    import numpy as np
    from mc3_b200 import stats as ms
    from oracle import problems as pb
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'hpd.npz'))
    post, bestp, pstep = pb.hpd_case()
    assert str(g['in_checksum']) == pb.checksum(post, bestp, pstep)
    for q in (0.683, 0.9545):
        st = ms.calc_sample_statistics(post, bestp, pstep, quantile=q, calc_hpd=True)
        for name, v in zip(('median', 'mean', 'std', 'med_lo', 'med_hi', 'mode', 'hpd_lo', 'hpd_hi'), st):
            np.testing.assert_allclose(v, g[f'{name}_{q}'], rtol=1e-9, atol=1e-12, err_msg=name)


def test_summary_stats_text_matches_reference_golden(tmp_path):
    """<root>_statistics.txt: same text as the reference's summary_stats
    (stats.py:967-1112) on the seeded sample of problems.hpd_case
    (tests/golden/summary_stats.txt, written by oracle/make_golden.py)."""
    import os
    import numpy as np
    from mc3_b200 import stats as ms
    from oracle import problems as pb
    here = os.path.dirname(os.path.abspath(__file__))
    want = open(os.path.join(here, 'golden', 'summary_stats.txt')).read()
    post, bestp, pstep = pb.hpd_case()
    out = {'bestp': bestp, 'pstep': pstep, 'pnames': [f'Param {i+1}' for i in range(len(bestp))],
           'texnames': [rf'$\\alpha_{i}$' for i in range(len(bestp))], 'best_chisq': 1234.56789,
           'best_log_post': -620.0, 'BIC': 1290.123456, 'red_chisq': 1.0345678,
           'stddev_residuals': 0.4987654321}
    got = ms.summary_stats(post, out, filename=str(tmp_path/'s.txt'), device=False)
    assert got == open(tmp_path/'s.txt').read()
    assert got == want


def test_tile_layout_of_piecewise_uniform_abscissae():
    """mc3_b200/gridseg.py: runs of a constant cadence are cut into whole 128-point tiles;
    what fills no tile is left over; anything else is refused."""
    from mc3_b200.gridseg import tile_layout, TILE
    x = np.linspace(0, 10, 100000)
    L = tile_layout(x)
    assert L['starts'].size == 100000//TILE and L['nleft'] == 100000 % TILE
    keep = np.ones(x.size, bool)
    keep[20000:23000] = False
    keep[60000:60010] = False
    keep[77777] = False
    xg = x[keep]
    xg[xg > 7.0] += 0.4*(x[1] - x[0])
    L = tile_layout(xg)
    assert L is not None and 0 < L['nleft'] <= 4*TILE
    assert np.array_equal(np.sort(L['perm']), np.arange(xg.size))
    assert abs(L['dx'] - (x[1] - x[0])) < 1e-15
    nt = L['starts'].size
    tiles = xg[L['perm'][:nt*TILE]].reshape(nt, TILE)
    assert np.array_equal(tiles[:, 0], xg[L['starts']])
    assert np.max(np.abs(tiles - (tiles[:, :1] + np.arange(TILE)*L['dx']))) <= 8*np.finfo(float).eps*10.4
    assert tile_layout(x + 1e-9*np.sin(np.arange(x.size))) is None              # jitter
    assert tile_layout(np.sort(np.random.RandomState(0).uniform(0, 10, 5000))) is None
    assert tile_layout(x[::-1]) is None and tile_layout(x[:100]) is None
    # many short runs: too much would be left over
    short = np.concatenate([i*5.0 + np.arange(150)*0.01 for i in range(40)])
    assert tile_layout(short) is None


def test_tile_layout_invariants_on_random_gapped_series():
    """Whatever the gaps, an accepted layout is a permutation, its tiles follow their own
    origins to the kernel's tolerance, and nothing but tile points precedes the leftovers."""
    from mc3_b200.gridseg import tile_layout, TILE
    rs = np.random.RandomState(12)
    accepted = 0
    for trial in range(40):
        n = int(rs.randint(600, 30000))
        dx = 10.0**rs.uniform(-4, 1)
        x = rs.uniform(-50, 50) + dx*np.arange(n)
        keep = np.ones(n, bool)
        for _ in range(rs.randint(0, 6)):
            lo = rs.randint(0, n)
            keep[lo:lo + rs.randint(1, max(2, n//10))] = False
        x = x[keep]
        if rs.rand() < 0.5 and x.size > 300:                     # the cadence resumes off the grid
            x[x.size//2:] += rs.uniform(0, 1)*dx
        L = tile_layout(x)
        if L is None:
            continue
        accepted += 1
        nt = L['starts'].size
        assert np.array_equal(np.sort(L['perm']), np.arange(x.size))
        assert L['perm'].size - nt*TILE == L['nleft'] <= max(64, int(0.02*x.size))
        tiles = x[L['perm'][:nt*TILE]].reshape(nt, TILE)
        tol = 8*np.finfo(float).eps*np.max(np.abs(x))
        assert np.max(np.abs(tiles - (tiles[:, :1] + np.arange(TILE)*L['dx']))) <= tol
        assert abs(L['dx'] - dx) <= 1e-12*dx
    assert accepted >= 10


def test_pair_and_moment_algebra_against_the_oracle():
    """The two rearrangements of the uniform-grid sinusoid chi-squared the CUDA kernels use
    (profiles/fold_error.py: mirrored pairs; profiles/moment_error.py: sufficient statistics
    of those pairs), emulated in numpy with the kernels' operation order, against the
    oracle's per-point chi-squared (src_c/_chisq.c:111-140 restated) -- no GPU involved."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles'))
    import fold_error as fe
    import moment_error as me
    from oracle import kernels as ok, models as om
    rs = np.random.RandomState(8)
    n = 128*40
    x = np.linspace(0.5, 9.5, n)
    truth = np.array([1.0, 2.5, 0.3, 5.0, -0.2])
    sigma = 0.5
    d = om.sinusoid(truth, x) + rs.normal(0, sigma, n)
    P = truth*(1 + 0.02*rs.standard_normal((24, 5)))
    P[:4, 1] = 10**rs.uniform(-2.5, -1, 4)                       # a few samples per period
    want = np.array([ok.chisq(om.sinusoid(p, x), d, np.full(n, sigma)) for p in P])
    x0, dx = x[0], (x[-1] - x[0])/(n - 1)
    pair = fe.folded_chisq(P, x0, dx, fe.fold(d), n)/sigma**2
    np.testing.assert_allclose(pair, want, rtol=1e-12)
    slr, c0r = np.polyfit(x, d, 1)
    f, mom, D2 = me.prepare(d, x0, dx, c0r, slr)
    mo = me.moment_chisq(P, x0, dx, f, mom, n, c0r, slr)/sigma**2
    amp = me.amp_bound(P, n, x0, dx, D2, c0r, slr, want*sigma**2)
    assert amp.max() < 4000                                       # the kernel's guard would pass them all
    np.testing.assert_allclose(mo, want, rtol=1e-11)
