"""Pins the oracle (oracle/) against the reference: (1) the known-answer values
of the reference's own tests/test_stats.py, (2) tests/golden/kernels.npz (outputs
of the reference's C extensions, made by oracle/make_golden.py), (3) the
reference mcmc() trajectories in tests/golden/mcmc_*.npz.  CPU only."""
import os

import numpy as np
import pytest

from oracle import kernels as ok
from oracle import mcmc as omc
from oracle import models as om
from oracle import problems as pb

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
K = np.load(os.path.join(GOLD, 'kernels.npz'))
RTOL = 1e-12


# ---- (1) reference tests/test_stats.py known answers ---------------------
def test_kat_chisq():                                  # test_stats.py:135-152
    data = np.array([1.1, 1.2, 0.9, 1.0])
    model, unc = np.ones(4), np.full(4, 0.1)
    np.testing.assert_allclose(ok.chisq(model, data, unc), 6.0, rtol=1e-14)
    c = ok.chisq(model, data, unc, np.array([2.5, 5.5]), np.array([2.0, 5.0]),
                 np.array([0.0, 1.0]), np.array([0.0, 1.0]))
    np.testing.assert_allclose(c, 6.25, rtol=1e-14)


def test_kat_residuals():                              # test_stats.py:115-132
    data = np.array([1.1, 1.2, 0.9, 1.0])
    r = ok.residuals(np.ones(4), data, np.full(4, 0.1), np.array([2.5, 5.5]),
                     np.array([2.0, 5.0]), np.array([0.0, 1.0]),
                     np.array([0.0, 1.0]))
    np.testing.assert_allclose(r, [-1.0, -2.0, 1.0, 0.0, 0.5], atol=1e-14)


def test_kat_dwt_chisq():                              # test_stats.py:155-180
    data = np.array([2.0, 0.0, 3.0, -2.0, -1.0, 2.0, 2.0, 0.0])
    params = np.array([1.0, 0.1, 0.1])
    np.testing.assert_allclose(ok.dwt_chisq(np.ones(8), data, params),
                               1693.22308882)
    c = ok.dwt_chisq(np.ones(8), data, params, np.array([1.0, 0.2, 0.3]),
                     np.array([0.0, 0.0, 0.1]), np.array([0.0, 0.0, 0.1]))
    np.testing.assert_allclose(c, 1697.2230888243134, rtol=1e-14)
    with pytest.raises(ValueError, match='at least three parameters'):
        ok.dwt_chisq(np.ones(8), data, params[:2])


DAUB4_INV, DAUB4_FWD = pb.KAT_DAUB4_INV, pb.KAT_DAUB4_FWD


def test_kat_daub4():                                  # test_stats.py:70-86, 280-302
    e4 = np.zeros(32)
    e4[4] = 1.0
    inv = ok.dwt_daub4(e4, True)
    np.testing.assert_allclose(inv, DAUB4_INV, atol=1e-10)
    np.testing.assert_allclose(ok.dwt_daub4(e4), DAUB4_FWD, atol=1e-10)
    np.testing.assert_allclose(ok.dwt_daub4(inv), e4, atol=1e-8)


def test_kat_bin_array():                              # test_stats.py:96-112
    data = np.array([0, 1, 2, 3, 3, 3, 3, 3, 4])
    unc = np.array([3, 1, 1, 1, 2, 3, 2, 2, 4])
    np.testing.assert_allclose(ok.bin_array(data, 3), [1.0, 3.0, 10/3])
    bd, bs = ok.bin_array(data, 3, unc)
    np.testing.assert_allclose(bd, [1.42105263, 3.0, 3.11111111])
    np.testing.assert_allclose(bs, [0.68824720, 0.85714286, 1.33333333])


RED_RMS, RED_RMSHI = pb.KAT_RED_RMS, pb.KAT_RED_RMSHI


def test_kat_time_avg():                               # test_stats.py:11-68, 306-343
    white, red = pb.teststats_series()
    rms, lo, hi, err, bsz = ok.time_avg(red, len(red)/10, 5)
    np.testing.assert_almost_equal(rms, RED_RMS)
    np.testing.assert_almost_equal(hi, RED_RMSHI)
    np.testing.assert_almost_equal(bsz, 1 + 5*np.arange(20))
    assert len(ok.time_avg(red)[0]) == 500
    assert len(ok.time_avg(red, 500, 2)[0]) == 250


def test_kat_log_prior():                              # stats.py:330-348 doc values
    post = np.array([[3.0, 2.0], [3.1, 1.0], [3.6, 1.5]])
    lp = ok.log_prior(post, np.array([3.5, 0.0]), np.array([0.1, 0.0]),
                      np.array([0.1, 0.0]), np.array([1.0, 1.0]))
    np.testing.assert_allclose(lp, [-12.5, -8.0, -0.5])


# ---- (2) golden outputs of the reference's C extensions -------------------
def test_golden_chisq():
    c = pb.chisq_case()
    assert pb.checksum(c['model'], c['data'], c['uncert']) == str(K['chisq_in'])
    np.testing.assert_allclose(
        ok.chisq(c['model'], c['data'], c['uncert']), K['chisq_noprior'],
        rtol=RTOL)
    args = (c['params'], c['priors'], c['priorlow'], c['priorup'])
    np.testing.assert_allclose(
        ok.chisq(c['model'], c['data'], c['uncert'], *args), K['chisq_prior'],
        rtol=RTOL)
    np.testing.assert_allclose(
        ok.residuals(c['model'], c['data'], c['uncert'], *args),
        K['residuals_prior'], rtol=RTOL)


@pytest.mark.parametrize('n', [8, 1024, 16384])
def test_golden_dwt_chisq(n):
    d = pb.dwt_case(n)
    assert pb.checksum(d['model'], d['data']) == str(K[f'dwt_in_{n}'])
    np.testing.assert_allclose(
        ok.dwt_chisq(d['model'], d['data'], d['params']),
        K[f'dwt_noprior_{n}'], rtol=RTOL)
    np.testing.assert_allclose(
        ok.dwt_chisq(d['model'], d['data'], d['params'], d['priors'],
                     d['priorlow'], d['priorup']),
        K[f'dwt_prior_{n}'], rtol=RTOL)


def test_golden_daub4():
    rs = np.random.RandomState(3)
    v1000, v4096 = rs.normal(0, 1, 1000), rs.normal(0, 1, 4096)
    for v, tag in ((v1000, '1000'), (v4096, '4096')):
        np.testing.assert_allclose(ok.dwt_daub4(v), K['daub4_fwd_' + tag],
                                   rtol=1e-11, atol=1e-13)
        np.testing.assert_allclose(ok.dwt_daub4(v, True), K['daub4_inv_' + tag],
                                   rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize('key,args', [
    ('tavg_red', ('red', 100, 5)), ('tavg_white', ('white', 100, 5)),
    ('tavg_red_default', ('red', None, 1)), ('tavg_2000', (2000, 1000, 1)),
    ('tavg_50k', (50000, 300, 7)), ('tavg_50k_big', (50000, 25000, 997))])
def test_golden_time_avg(key, args):
    src, maxbins, binstep = args
    if src in ('red', 'white'):
        white, red = pb.teststats_series()
        data = red if src == 'red' else white
    else:
        data = pb.series_case(src, {2000: 5, 50000: 8}[src])
    out = np.array(ok.time_avg(data, maxbins, binstep))
    np.testing.assert_allclose(out, K[key], rtol=1e-10)


@pytest.mark.parametrize('bs', [100, 7, 4099])
def test_golden_bin_array(bs):
    d, u = pb.binarray_case()
    np.testing.assert_allclose(ok.bin_array(d, bs), K[f'bin_unw_{bs}'],
                               rtol=RTOL)
    np.testing.assert_allclose(np.array(ok.bin_array(d, bs, u)),
                               K[f'bin_w_{bs}'], rtol=RTOL)


def test_golden_gelman_and_log_prior():
    Z, zc, burn = pb.gelman_case()
    np.testing.assert_allclose(ok.gelman_rubin(Z, zc, burn), K['gelman'],
                               rtol=RTOL)
    c = pb.chisq_case()
    post = np.random.RandomState(9).normal(c['params'], 0.3, (50, 6))
    np.testing.assert_allclose(
        ok.log_prior(post, c['priors'], c['priorlow'], c['priorup'],
                     np.ones(6)), K['log_prior'], rtol=RTOL)


# ---- (3) reference mcmc() trajectories ------------------------------------
@pytest.mark.parametrize('case', pb.MCMC_CASES)
@pytest.mark.parametrize('sampler', pb.SAMPLERS)
def test_golden_mcmc(case, sampler):
    """The oracle loop, run on the oracle's own C kernels, retraces the
    reference's single-process run: identical accept/reject flags and posterior
    rows, log-posterior to 1e-12 (the chi-squared sums differ in order)."""
    fx = np.load(os.path.join(GOLD, f'mcmc_{case}_{sampler}.npz'))
    p = pb.mcmc_case(case)
    assert pb.checksum(p['x'], p['data'], p['uncert']) == str(fx['in_checksum'])
    with np.errstate(all='ignore'):
        b = omc.mcmc(
            p['data'], p['uncert'], om.MODELS[p['model']], p['params'],
            [p['x']], {}, p['pmin'], p['pmax'], p['pstep'], p['prior'],
            p['priorlow'], p['priorup'], nchains=p['nchains'],
            nsamples=p['nsamples'], sampler=sampler, wlike=p['wlike'],
            burnin=p['burnin'], thinning=p['thinning'], fepsilon=p['fepsilon'],
            parent_seed=pb.PARENT_SEED, child_seed=pb.CHILD_SEED)
    assert np.array_equal(b['draws']['acc'], fx['draw_acc'])
    assert np.array_equal(b['zchain'], fx['ref_zchain'])
    np.testing.assert_allclose(b['posterior'], fx['ref_posterior'], rtol=1e-13)
    np.testing.assert_allclose(b['log_post'], fx['ref_log_post'], rtol=1e-12)
    np.testing.assert_allclose(b['bestp'], fx['ref_bestp'], rtol=1e-13)
    assert b['numaccept'] == int(fx['numaccept'])
    assert np.array_equal(b['outbounds'], fx['outbounds'])


def test_fixture_decisions_are_not_marginal():
    """Replay parity is meaningful only if no Metropolis decision of the
    fixtures sits within rounding of its threshold: the relative gap between
    exp(0.5 (chisq - chisq*)) and the uniform draw stays far above the 1e-10
    band in which the CUDA chi-squared may differ from the reference's."""
    worst = np.inf
    for case in pb.MCMC_CASES:
        for sampler in pb.SAMPLERS:
            fx = np.load(os.path.join(GOLD, f'mcmc_{case}_{sampler}.npz'))
            cur = -2.0*fx['log_post0'][:fx['draw_u'].shape[1]].copy()
            G = int(fx['draw_ngen'])
            for g in range(G):
                for j in range(cur.size):
                    if not fx['draw_done'][g, j] or not fx['draw_inb'][g, j]:
                        continue
                    prop, u = fx['draw_chisq'][g, j], fx['draw_u'][g, j]
                    plain = sampler != 'snooker' or not fx['draw_usj'][g, j] < 0.1
                    if plain and np.isfinite(prop):
                        ratio = np.exp(0.5*(cur[j] - prop))
                        assert (ratio > u) == bool(fx['draw_acc'][g, j])
                        worst = min(worst, abs(ratio - u)/max(ratio, u))
                    if fx['draw_acc'][g, j]:
                        cur[j] = prop
    assert worst > 1e-7, worst
