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
