"""GPU tests of the sampler path.

(1) Replay: fed the reference's recorded random stream (tests/golden/mcmc_*.npz,
    generated from the real reference by oracle/make_golden.py) the CUDA path
    must retrace the reference's accept/reject trajectory exactly -- history
    rows, chain ids, acceptance count, out-of-bounds counters, best fit.
(2) Production (lock-step, Philox): posterior moments agree with the reference
    run within Monte-Carlo error; determinism; API contract of sample().
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import kernels as ok
from oracle import models as om
from oracle import problems as pb

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.fixture(scope='module')
def mc3():
    import mc3_b200
    return mc3_b200


def close_abs(a, b, tol):
    """assert |a-b| <= tol elementwise (array tolerances)."""
    a, b, tol = np.asarray(a), np.asarray(b), np.asarray(tol)
    assert np.all(np.abs(a - b) <= tol), (a, b, tol)


def _population(mc3, p, sampler, **kw):
    from mc3_b200.engine import Population
    nzchain = int(np.ceil(p['nsamples']/p['nchains']/p['thinning']))
    return Population(
        p['data'], p['uncert'], mc3.models.BUILTIN[p['model']], p['params'],
        [p['x']], {}, p['pstep'], p['pmin'], p['pmax'], p['prior'],
        p['priorlow'], p['priorup'], nchains=p['nchains'], sampler=sampler,
        wlike=p['wlike'], fgamma=1.0, fepsilon=p['fepsilon'], hsize=10,
        thinning=p['thinning'], nzchain=nzchain, **kw)


@pytest.mark.parametrize('case', pb.MCMC_CASES)
@pytest.mark.parametrize('sampler', pb.SAMPLERS)
def test_replay_retraces_reference(mc3, case, sampler):
    fx = np.load(os.path.join(GOLD, f'mcmc_{case}_{sampler}.npz'))
    p = pb.mcmc_case(case)
    assert pb.checksum(p['x'], p['data'], p['uncert']) == str(fx['in_checksum'])
    pop = _population(mc3, p, sampler)
    pop.set_initial(fx['Z0'], fx['log_post0'])
    # the device re-evaluates the initial rows: same log-posterior as the reference
    P0 = np.tile(p['params'], (pop.M0, 1))
    P0[:, pop.ifree] = fx['Z0']
    for s in pop.ishare:
        P0[:, s] = P0[:, -int(p['pstep'][s]) - 1]
    lp0 = -0.5*pop.chisq(torch.as_tensor(P0, device=pop.dev)).cpu().numpy()
    np.testing.assert_allclose(lp0, fx['log_post0'], rtol=1e-10)

    draws = {k[5:]: fx[k] for k in fx.files if k.startswith('draw_')}
    pop.replay(draws)
    torch.cuda.synchronize()

    M0 = pop.M0
    nrows = fx['ref_posterior'].shape[0]
    Z = pop.Z[M0:M0 + nrows].cpu().numpy()
    zchain = pop.zchain[M0:M0 + nrows].cpu().numpy()
    log_post = pop.log_post[M0:M0 + nrows].cpu().numpy()
    # accept/reject trajectory: a row repeats its predecessor exactly when the
    # proposal was rejected, so equal chain ids + equal rows = equal trajectory.
    assert np.array_equal(zchain, fx['ref_zchain'])
    c = pop.counters()
    assert c['numaccept'] == int(fx['numaccept'])
    assert np.array_equal(c['outbounds'], fx['outbounds'])
    np.testing.assert_allclose(Z, fx['ref_posterior'], rtol=1e-10, atol=1e-13)
    moved_ref = np.any(np.diff(fx['ref_posterior'].reshape(-1, p['nchains'], pop.nfree), axis=0) != 0, axis=2)
    moved_gpu = np.any(np.diff(Z.reshape(-1, p['nchains'], pop.nfree), axis=0) != 0, axis=2)
    assert np.array_equal(moved_ref, moved_gpu)
    np.testing.assert_allclose(log_post, fx['ref_log_post'], rtol=1e-10)
    np.testing.assert_allclose(c['bestp'], fx['ref_bestp'], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(c['best_log_post'], fx['ref_best_log_post'], rtol=1e-10)
    np.testing.assert_allclose(pop.X.cpu().numpy(), fx['final_freepars'], rtol=1e-10,
                               atol=1e-13)


@pytest.mark.parametrize('sampler', ['snooker', 'demc'])
def test_production_matches_reference_docs_get_started(mc3, sampler):
    """BASELINE config 1 (examples/get_started.py data, seed 3).  The reference's
    documentation pins its posterior for this problem (docs/get_started.rst:104-128):
    best chisq 112.5923 at [3.0768, -2.5000, 0.5089]; medians [3.0761, -2.4981,
    0.50868] with 1-sigma bounds (-0.3797,+0.3895), (-0.2288,+0.2133),
    (-0.02647,+0.02742).  The lock-step run must land on the same posterior."""
    p = pb.mcmc_case('quad')
    out = mc3.sample(
        p['data'], p['uncert'], func=mc3.models.polynomial, params=p['params'],
        indparams=[p['x']], pstep=p['pstep'], sampler=sampler, nchains=256,
        nsamples=256*900, burnin=400, thinning=1, grtest=True, seed=11,
        log=mc3.Log(verb=-1), fepsilon=0.0 if sampler != 'demc' else 0.001)
    post, _, _ = mc3.utils.burn(out)
    assert post.shape == (256*500, 3)
    sig = np.array([0.385, 0.221, 0.027])
    assert 112.58 < out['best_chisq'] < 112.5923 + 0.01     # true minimum is 112.5898
    np.testing.assert_allclose(
        out['best_chisq'],
        ok.chisq(om.polynomial(out['bestp'], p['x']), p['data'], p['uncert']), rtol=1e-10)
    close_abs(out['bestp'], [3.0768, -2.5000, 0.5089], 0.15*sig)
    close_abs(out['medianp'], [3.0761, -2.4981, 0.50868], 0.1*sig)
    lo = out['median_low_bounds'] - out['medianp']
    hi = out['median_high_bounds'] - out['medianp']
    np.testing.assert_allclose(lo, [-0.37968, -0.22876, -0.026467], rtol=0.08)
    np.testing.assert_allclose(hi, [0.38946, 0.21325, 0.027415], rtol=0.08)
    np.testing.assert_allclose(out['stdp'], sig, rtol=0.08)
    assert 10.0 < out['acceptance_rate'] < 45.0      # reference: 28.36 %


def test_production_posterior_vs_analytic_linear_model(mc3):
    """For a linear model with Gaussian errors the posterior is Gaussian with
    known mean/covariance (weighted least squares): a sharp, size-independent
    check of the whole loop (proposals, chi-squared, accept, thinning)."""
    rs = np.random.RandomState(42)
    n = 4000
    x = np.linspace(-1, 1, n)
    ptrue = np.array([1.0, -0.5, 0.25])
    unc = rs.uniform(0.5, 1.5, n)
    data = om.polynomial(ptrue, x) + rs.normal(0, 1, n)*unc
    A = np.vstack([x**0, x, x**2]).T/unc[:, None]
    cov = np.linalg.inv(A.T @ A)
    mean = cov @ (A.T @ (data/unc))
    sig = np.sqrt(np.diag(cov))
    for sampler in ('snooker', 'demc', 'mrw'):
        ngen = 500 if sampler != 'mrw' else 2500
        out = mc3.sample(data, unc, func=mc3.models.polynomial, params=mean + sig,
                         indparams=[x], pstep=sig*(1.0 if sampler != 'mrw' else 0.6),
                         sampler=sampler, nchains=1024, nsamples=1024*ngen,
                         burnin=ngen*2//5, seed=5, fepsilon=1e-3,
                         thinning=1 if sampler != 'mrw' else 5,
                         log=mc3.Log(verb=-1))
        post, _, _ = mc3.utils.burn(out)
        close_abs(post.mean(axis=0), mean, 0.05*sig)
        np.testing.assert_allclose(post.std(axis=0), sig, rtol=0.05)
        corr = np.corrcoef(post.T)
        want = cov/np.outer(sig, sig)
        np.testing.assert_allclose(corr, want, atol=0.05)
        close_abs(out['bestp'], mean, 0.3*sig)
        assert 5.0 < out['acceptance_rate'] < 70.0


def test_same_seed_same_bytes_and_graph_equals_eager(mc3):
    p = pb.mcmc_case('sine')
    kw = dict(data=p['data'], uncert=p['uncert'], func=mc3.models.sinusoid,
              params=p['params'], indparams=[p['x']], pstep=p['pstep'],
              pmin=p['pmin'], pmax=p['pmax'], prior=p['prior'],
              priorlow=p['priorlow'], priorup=p['priorup'], sampler='demc',
              nchains=128, nsamples=128*50, burnin=10, thinning=2, fepsilon=0.01,
              seed=3, log=mc3.Log(verb=-1))    # > 64 chains: per-generation kernels
    a = mc3.sample(**kw, use_graph=True)
    b = mc3.sample(**kw, use_graph=True)
    c = mc3.sample(**kw, use_graph=False)
    for k in ('posterior', 'log_post', 'zchain', 'bestp'):
        assert np.array_equal(a[k], b[k]), k
        assert np.array_equal(a[k], c[k]), k
    assert a['acceptance_rate'] == c['acceptance_rate']
    d = mc3.sample(**{**kw, 'seed': 4})
    assert not np.array_equal(a['posterior'], d['posterior'])


def test_sample_api_contract(mc3, tmp_path):
    """Output keys / shapes / semantics the reference's tests rely on
    (reference tests/test_mcmc.py:51-120, 184-196, 270-280)."""
    p = pb.mcmc_case('share')
    unc0 = np.copy(p['uncert'])
    os.chdir(tmp_path)
    out = mc3.sample(p['data'], p['uncert'], func=mc3.models.polynomial,
                     params=np.copy(p['params']), indparams=[p['x']],
                     pstep=p['pstep'], pmin=p['pmin'], pmax=p['pmax'],
                     sampler='snooker', nchains=14, nsamples=14*300, burnin=50,
                     seed=1, log=mc3.Log(verb=-1),
                     prior=np.array([1.0, 0, 0, 0, 0]), priorlow=np.array([0.05, 0, 0, 0, 0]),
                     priorup=np.array([0.05, 0, 0, 0, 0]))
    keys = {'pnames', 'texnames', 'pstep', 'ifree', 'burnin', 'posterior', 'zchain',
            'chisq', 'log_post', 'acceptance_rate', 'bestp', 'best_chisq', 'red_chisq',
            'BIC', 'best_log_post', 'best_model', 'stddev_residuals', 'zmask',
            'medianp', 'meanp', 'stdp', 'median_low_bounds', 'median_high_bounds',
            'mode', 'hpd_low_bounds', 'hpd_high_bounds', 'CRlo', 'CRhi', 'chisq_factor'}
    assert keys <= set(out)
    assert out['posterior'].shape == (14*300, 3)
    assert out['zchain'].min() == 0 and out['zchain'].max() == 13
    assert out['bestp'][3] == out['bestp'][1]                 # shared parameter
    assert out['stdp'][4] == 0 and out['CRlo'][4] == 0 and out['CRhi'][4] == 0   # fixed
    assert out['bestp'][4] == p['params'][4]
    assert np.all(-2*out['log_post'] > out['chisq'] - 1e-9)   # Gaussian prior adds
    assert np.any(-2*out['log_post'] > out['chisq'] + 1e-6)
    assert out['best_model'].shape == p['data'].shape
    assert np.array_equal(p['uncert'], unc0)                  # caller's uncert untouched
    assert os.path.exists('mc3_statistics.txt')
    np.testing.assert_allclose(
        out['best_chisq'],
        ok.chisq(om.polynomial(out['bestp'], p['x']), p['data'], p['uncert']), rtol=1e-9)
    # device-side chi-squared column and burned-sample statistics == host formulas
    prior = np.array([1.0, 0, 0, 0, 0])
    plo = np.array([0.05, 0, 0, 0, 0])
    lpr = ok.log_prior(out['posterior'], prior, plo, plo, p['pstep'])
    np.testing.assert_allclose(out['chisq'], -2.0*(out['log_post'] - lpr), rtol=1e-12)
    post, zc, zmask = mc3.utils.burn(out)
    assert np.array_equal(zmask, out['zmask'])
    want = mc3.stats.calc_sample_statistics(post, out['bestp'], p['pstep'])
    # (sample() later overwrites these keys with the 20000-row-subsample values;
    # here all rows are used, so both agree)
    for k, wv in zip(('medianp', 'meanp', 'stdp', 'median_low_bounds',
                      'median_high_bounds'), want):
        np.testing.assert_allclose(out[k], wv, rtol=1e-12, atol=1e-15)


def test_user_callables_numpy_and_torch(mc3):
    """func as the reference's numpy callable (host model, GPU chi-squared) and
    as a batched torch callable: same sampler, same seed, same Philox draws ->
    same trajectory as the built-in model to rounding."""
    p = pb.mcmc_case('quad')
    kw = dict(data=p['data'], uncert=p['uncert'], params=p['params'],
              indparams=[p['x']], pstep=p['pstep'], sampler='demc', nchains=12,
              nsamples=12*40, seed=9, log=mc3.Log(verb=-1))
    a = mc3.sample(func=mc3.models.polynomial, **kw)
    b = mc3.sample(func=om.quad, **kw)

    def tquad(P, x):
        return P[:, 0:1] + P[:, 1:2]*x[None, :] + P[:, 2:3]*x[None, :]**2
    c = mc3.sample(func=mc3.TorchModel(tquad), **kw)
    for o in (b, c):
        assert np.array_equal(a['zchain'], o['zchain'])
        np.testing.assert_allclose(a['posterior'], o['posterior'], rtol=1e-9)
        np.testing.assert_allclose(a['log_post'], o['log_post'], rtol=1e-9)


def test_wavelet_likelihood_run_and_gr_break(mc3):
    p = pb.mcmc_case('wave')
    out = mc3.sample(p['data'], p['uncert'], func=mc3.models.box, params=p['params'],
                     indparams=[p['x']], pstep=p['pstep'], pmin=p['pmin'],
                     pmax=p['pmax'], sampler='snooker', nchains=64, nsamples=64*300,
                     burnin=50, wlike=True, seed=2, log=mc3.Log(verb=-1))
    assert out['posterior'].shape == (64*300, 6)
    close_abs(out['bestp'][:4], [0.01, 0.0, 0.1, 1.0], [2e-3, 0.02, 0.02, 1e-3])
    # best chi-squared equals the oracle's wavelet likelihood at bestp
    np.testing.assert_allclose(
        -2*out['best_log_post'],
        ok.dwt_chisq(om.box(out['bestp'][:4], p['x']), p['data'], out['bestp']), rtol=1e-9)
    # Gelman-Rubin early stop (reference tests/test_mcmc.py:199-267)
    q = pb.mcmc_case('quad')
    o2 = mc3.sample(q['data'], q['uncert'], func=mc3.models.polynomial, params=q['params'],
                    indparams=[q['x']], pstep=q['pstep'], sampler='snooker', nchains=64,
                    nsamples=64*2000, burnin=100, grtest=True, grbreak=1.05, grnmin=0.2,
                    seed=4, log=mc3.Log(verb=-1))
    assert o2['posterior'].shape[0] < 64*2000*0.9


def test_gelman_rubin_device_matches_oracle(mc3):
    p = pb.mcmc_case('quad')
    out = mc3.sample(p['data'], p['uncert'], func=mc3.models.polynomial, params=p['params'],
                     indparams=[p['x']], pstep=p['pstep'], sampler='demc', nchains=40,
                     nsamples=40*120, burnin=20, seed=6, log=mc3.Log(verb=-1),
                     return_population=False)
    want = ok.gelman_rubin(out['posterior'], out['zchain'], 20)
    got = mc3.stats.gelman_rubin(out['posterior'], out['zchain'], 20)
    np.testing.assert_allclose(got, want, rtol=1e-10)
    # the hub's own path: lock-step row formula on the device-resident history
    from mc3_b200.mcmc_driver import mcmc
    o2 = mcmc(p['data'], p['uncert'], mc3.models.polynomial, p['params'], [p['x']], {},
              p['pmin'], p['pmax'], p['pstep'], p['prior'], p['priorlow'], p['priorup'],
              40, None, 40*120, 'demc', False, None, True, 0.0, 0.5, 40, 2, 1.0, 0.0,
              10, 'normal', None, False, mc3.Log(verb=-1), None, None, seed=6,
              return_population=True)
    pop = o2['_population']
    zburn = o2['burnin']
    assert zburn == 20 and pop.thinned_done() == 60
    np.testing.assert_allclose(pop.gelman_rubin(zburn),
                               ok.gelman_rubin(o2['posterior'], o2['zchain'], zburn), rtol=1e-10)


@pytest.mark.parametrize('sampler', pb.SAMPLERS)
def test_persistent_small_kernel_matches_per_generation_kernels(mc3, sampler):
    """Small populations run inside one resident CTA (mc3b_run_small); the same
    seed through the per-generation kernels (use_graph=False) gives the same
    trajectory -- only the chi-squared summation order differs."""
    p = pb.mcmc_case('sine')
    kw = dict(data=p['data'], uncert=p['uncert'], func=mc3.models.sinusoid,
              params=p['params'], indparams=[p['x']], pstep=p['pstep'], pmin=p['pmin'],
              pmax=p['pmax'], prior=p['prior'], priorlow=p['priorlow'],
              priorup=p['priorup'], sampler=sampler, nchains=8, nsamples=8*400,
              burnin=20, thinning=2, fepsilon=0.01, seed=13, log=mc3.Log(verb=-1))
    a = mc3.sample(**kw)                       # persistent kernel
    b = mc3.sample(**kw, use_graph=False)      # propose / model_chisq / metropolis launches
    assert np.array_equal(a['zchain'], b['zchain'])
    same = np.all(a['posterior'] == b['posterior'], axis=1).mean()
    assert same > 0.97, same
    n = 8*20
    np.testing.assert_allclose(a['posterior'][:n], b['posterior'][:n], rtol=1e-12)
    np.testing.assert_allclose(a['log_post'][:n], b['log_post'][:n], rtol=1e-11)
    assert abs(a['acceptance_rate'] - b['acceptance_rate']) < 2.0


def test_reflect_mode_folds_proposals_inside_bounds(mc3):
    """Opt-in, non-reference behaviour (BASELINE north_star wording): out-of-bounds
    proposals are folded back instead of rejected.  Nothing is ever counted out of
    bounds, every sample respects the bounds, and the posterior still matches."""
    p = pb.mcmc_case('sine')
    kw = dict(data=p['data'], uncert=p['uncert'], func=mc3.models.sinusoid,
              params=p['params'], indparams=[p['x']], pstep=p['pstep'], pmin=p['pmin'],
              pmax=p['pmax'], sampler='demc', nchains=128, nsamples=128*300, burnin=100,
              fepsilon=0.01, seed=8, log=mc3.Log(verb=-1))
    from mc3_b200.mcmc_driver import mcmc
    outs = {}
    for reflect in (False, True):
        o = mcmc(p['data'], p['uncert'], mc3.models.sinusoid, p['params'], [p['x']], {},
                 p['pmin'], p['pmax'], p['pstep'], None, None, None, 128, None, 128*300,
                 'demc', False, None, False, 0.0, 0.5, 100, 1, 1.0, 0.01, 10, 'normal',
                 None, False, mc3.Log(verb=-1), None, None, seed=8, reflect=reflect,
                 return_population=True)
        outs[reflect] = (o, o['_population'].counters()['outbounds'])
    assert outs[False][1].sum() > 0            # the bounds of this case do get hit
    assert outs[True][1].sum() == 0
    post = outs[True][0]['posterior']
    assert np.all(post >= p['pmin'][None, :]) and np.all(post <= p['pmax'][None, :])
    a = outs[False][0]['posterior'][128*100:]
    b = post[128*100:]
    assert np.all(np.abs(a.mean(axis=0) - b.mean(axis=0)) < 0.5*a.std(axis=0))
