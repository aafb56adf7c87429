"""GPU: the API-contract tests of the reference's tests/test_mcmc.py (same
dataset, same calls, same assertions) run against the drop-in mc3_b200.sample().
Out-of-scope cases (plots, func given as strings, CLI) are not mirrored."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def quad(p, x):
    return p[0] + p[1]*x + p[2]*x**2.0


# reference tests/test_mcmc.py:29-48 (legacy global stream, seed 12)
_rs = np.random.RandomState(12)
x = np.linspace(0, 10, 100)
p0 = [4.5, -2.4, 0.5]
y = quad(p0, x)
uncert = np.sqrt(np.abs(y))
data = y + _rs.normal(0, uncert)
params = np.array([10.0, -2.0, 0.1])
pstep = np.array([0.03, 0.03, 0.05])
pnames = ["constant", "linear", "quadratic"]
texnames = ["$\\alpha$", "$\\log(\\beta)$", "quadratic"]


@pytest.fixture(scope='module')
def mc3():
    import mc3_b200
    return mc3_b200


@pytest.mark.parametrize('sampler', ['snooker', 'demc', 'mrw'])
def test_mcmc_minimal(mc3, sampler, tmp_path):             # test_mcmc.py:51-64
    os.chdir(tmp_path)
    out = mc3.sample(data, uncert, func=quad, params=np.copy(params), sampler=sampler,
                     indparams=[x], pstep=pstep, nsamples=1e4, burnin=100)
    assert out is not None
    assert out['posterior'].shape == (10003, 3)            # 7 chains x 1429 thinned steps
    assert out['zchain'].max() == 6


def test_mcmc_indparams_dict_and_names(mc3, capsys, tmp_path):   # :78-86, 120-153
    os.chdir(tmp_path)
    out = mc3.sample(data, uncert, func=quad, params=np.copy(params), sampler='snooker',
                     indparams_dict={'x': x}, pstep=pstep, nsamples=1e4, burnin=100,
                     pnames=pnames, texnames=texnames)
    cap = capsys.readouterr().out
    for name in pnames:
        assert name in cap
    assert list(out['pnames']) == pnames and list(out['texnames']) == texnames
    out = mc3.sample(data, uncert, func=quad, params=np.copy(params), sampler='snooker',
                     indparams=[x], pstep=pstep, nsamples=1e4, burnin=100)
    cap = capsys.readouterr().out
    assert 'Param 1' in cap and 'Param 3' in cap


def test_mcmc_shared_fixed_bounds(mc3, tmp_path):          # :88-118
    os.chdir(tmp_path)
    out = mc3.sample(data, uncert, func=quad, params=np.array([4.5, 4.5, 0.5]),
                     sampler='snooker', indparams=[x], pstep=[0.03, -1, 0.05],
                     nsamples=1e4, burnin=100)
    assert out['bestp'][1] == out['bestp'][0]
    pars = np.copy(params)
    pars[0] = p0[0]
    out = mc3.sample(data, uncert, func=quad, params=np.copy(pars), sampler='snooker',
                     indparams=[x], pstep=[0, 0.03, 0.05], nsamples=1e4, burnin=100)
    assert len(out['bestp']) == len(params)
    assert out['bestp'][0] == pars[0]
    assert out['CRlo'][0] == 0 and out['CRhi'][0] == 0 and out['stdp'][0] == 0
    out = mc3.sample(data, uncert, func=quad, params=np.copy(params), sampler='snooker',
                     indparams=[x], pstep=pstep, nsamples=1e4, burnin=100,
                     pmin=[-10.0, -20.0, -2.0], pmax=[40.0, 20.0, 5.0])
    assert np.all(out['posterior'] >= [-10.0, -20.0, -2.0])
    assert np.all(out['posterior'] <= [40.0, 20.0, 5.0])


@pytest.mark.parametrize('leastsq', ['lm', 'trf'])
def test_mcmc_optimize(mc3, capsys, leastsq, tmp_path):    # :156-181
    os.chdir(tmp_path)
    out = mc3.sample(data, uncert, func=quad, params=np.copy(params), sampler='snooker',
                     indparams=[x], pstep=pstep, nsamples=1e4, burnin=100, leastsq=leastsq)
    cap = capsys.readouterr().out
    assert "Least-squares best-fitting parameters:" in cap
    np.testing.assert_allclose(out['bestp'], [4.28263253, -2.40781859, 0.49534411],
                               rtol=1e-7)


def test_mcmc_optimize_builtin_model_and_chisqscale(mc3, capsys, tmp_path):   # :184-196
    os.chdir(tmp_path)
    unc = np.copy(uncert)
    out = mc3.sample(data, uncert, func=mc3.models.polynomial, params=np.copy(params),
                     sampler='snooker', indparams=[x], pstep=pstep, nsamples=1e4,
                     burnin=100, leastsq='lm', chisqscale=True)
    cap = capsys.readouterr().out
    assert "Least-squares best-fitting parameters (rescaled chisq):" in cap
    assert "Reduced chi-squared:                  1.0000" in cap
    np.testing.assert_equal(uncert, unc)
    np.testing.assert_allclose(out['bestp'], [4.28263253, -2.40781859, 0.49534411],
                               rtol=1e-6)
    assert out['chisq_factor'] != 1.0


def test_mcmc_gr_text_and_priors(mc3, capsys, tmp_path):   # :199-211, 270-280
    os.chdir(tmp_path)
    mc3.sample(data, uncert, func=quad, params=np.copy(params), sampler='snooker',
               indparams=[x], pstep=pstep, nsamples=1e4, burnin=100, grtest=True)
    assert "Gelman-Rubin statistics for free parameters" in capsys.readouterr().out
    out = mc3.sample(data, uncert, func=quad, params=np.copy(params), sampler='snooker',
                     indparams=[x], pstep=pstep, nsamples=1e4, burnin=100,
                     prior=np.array([4.5, 0.0, 0.0]), priorlow=np.array([0.1, 0.0, 0.0]),
                     priorup=np.array([0.1, 0.0, 0.0]))
    assert -2*out['best_log_post'] > out['best_chisq']
    assert np.all(-2*out['log_post'] > out['chisq'])


def test_mcmc_log_savefile_resume(mc3, capsys, tmp_path):  # :283-310 (+ resume)
    os.chdir(tmp_path)
    mc3.sample(data, uncert, func=quad, params=np.copy(params), sampler='snooker',
               indparams=[x], pstep=pstep, nsamples=1e4, burnin=100, log='MCMC.log')
    assert 'MCMC.log' in capsys.readouterr().out and 'MCMC.log' in os.listdir('.')
    out = mc3.sample(data, uncert, func=quad, params=np.copy(params), sampler='demc',
                     indparams=[x], pstep=pstep, nsamples=7000, burnin=100,
                     savefile='MCMC.npz', seed=3)
    assert 'MCMC.npz' in os.listdir('.') and 'MCMC_statistics.txt' in os.listdir('.')
    saved = np.load('MCMC.npz')
    for k in ('posterior', 'zchain', 'log_post', 'acceptance_rate', 'bestp',
              'best_log_post', 'chisq_factor'):
        assert k in saved.files                   # the keys the reference's resume reads
    out2 = mc3.sample(data, uncert, func=quad, params=np.copy(params), sampler='demc',
                      indparams=[x], pstep=pstep, nsamples=7000, burnin=100,
                      savefile='MCMC.npz', resume=True, seed=4)
    n1 = out['posterior'].shape[0]
    assert out2['posterior'].shape[0] == 2*n1
    assert np.array_equal(out2['posterior'][:n1], out['posterior'])
    assert out2['best_log_post'] >= out['best_log_post']
    # every chain continued from its last sample: first new row of a chain is
    # either that sample (rejected step) or a move away from it
    last = out['posterior'][-7:]
    first = out2['posterior'][n1:n1 + 7]
    assert np.any(np.all(first == last, axis=1)) or np.all(np.abs(first - last) < 5)


def test_cannot_populate_initial_sample(mc3, tmp_path):    # :412-430
    os.chdir(tmp_path)

    def limited_quad(p, x):
        yy = p[0] + p[1]*x + p[2]*x**2.0
        if p[0] > 4.0:
            yy[:] = np.inf
        return yy
    with pytest.raises(ValueError, match='Cannot populate an initial sample set of parameters'):
        mc3.sample(data, uncert, func=limited_quad, params=np.copy(params), indparams=[x],
                   pstep=pstep, sampler='snooker', nsamples=1e4, burnin=100)


def test_kickoff_uniform_and_small_populations(mc3, tmp_path):
    os.chdir(tmp_path)
    out = mc3.sample(data, uncert, func=mc3.models.polynomial, params=np.copy(params),
                     indparams=[x], pstep=pstep, sampler='demc', nchains=3, nsamples=3000,
                     burnin=100, kickoff='uniform', pmin=[0.0, -5.0, 0.0], pmax=[12.0, 0.0, 1.0],
                     thinning=3, log=mc3.Log(verb=-1))
    assert out['posterior'].shape == (int(np.ceil(3000/3/3))*3, 3)
    assert np.all(out['posterior'] >= [0.0, -5.0, 0.0]) and np.all(out['posterior'] <= [12.0, 0.0, 1.0])
    with pytest.raises(Exception, match='at least 3 chains'):
        mc3.sample(data, uncert, func=mc3.models.polynomial, params=np.copy(params),
                   indparams=[x], pstep=pstep, sampler='demc', nchains=2, nsamples=300,
                   log=mc3.Log(verb=-1))
