"""ORACLE (test infrastructure only): write tests/golden/*.npz from the REAL
reference.  Runs in the authoring container only (needs /root/reference and
oracle/_ref); the fixtures are committed so the tests never touch the
reference at run time.

    python -m oracle.make_golden            # from the repo root

kernels.npz               outputs of the reference's mc3.stats / C extensions on
                          the seeded inputs of oracle/problems.py
mcmc_<case>_<sampler>.npz reference mcmc() run single-process with pinned seeds
                          (posterior, zchain, log_post, bestp, ...) plus the
                          replay draw log.  The log is recorded by the oracle
                          loop, and this script ASSERTS that the oracle's
                          posterior / log_post / zchain / bestp are byte-equal
                          to the reference's before writing anything, so the
                          log is the reference's own random stream.
"""
import os
import random
import sys

import numpy as np

from . import kernels as ok
from . import mcmc as omc
from . import models as om
from . import problems as pb
from . import ref

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                    'tests', 'golden')


def kernels_golden(R):
    ms = R.stats
    cs, dwt, ta, ba = ref.ref_ext()
    g = {}
    # chisq / residuals (stats.py:94-216 -> _chisq.c)
    c = pb.chisq_case()
    g['chisq_in'] = pb.checksum(c['model'], c['data'], c['uncert'])
    g['chisq_noprior'] = ms.chisq(c['model'], c['data'], c['uncert'])
    g['chisq_prior'] = ms.chisq(c['model'], c['data'], c['uncert'], c['params'],
                                c['priors'], c['priorlow'], c['priorup'])
    g['residuals_prior'] = ms.residuals(
        c['model'], c['data'], c['uncert'], c['params'], c['priors'],
        c['priorlow'], c['priorup'])
    # dwt chisq (stats.py:219-284 -> _dwt.c) at N = 2^k
    for n in (8, 1024, 16384):
        d = pb.dwt_case(n)
        g[f'dwt_in_{n}'] = pb.checksum(d['model'], d['data'])
        g[f'dwt_noprior_{n}'] = ms.dwt_chisq(d['model'], d['data'], d['params'])
        g[f'dwt_prior_{n}'] = ms.dwt_chisq(
            d['model'], d['data'], d['params'], d['priors'], d['priorlow'],
            d['priorup'])
    # daub4 (stats.py:577-611)
    rs = np.random.RandomState(3)
    v1000, v4096 = rs.normal(0, 1, 1000), rs.normal(0, 1, 4096)
    g['daub4_fwd_1000'] = ms.dwt_daub4(v1000)
    g['daub4_inv_1000'] = ms.dwt_daub4(v1000, True)
    g['daub4_fwd_4096'] = ms.dwt_daub4(v4096)
    g['daub4_inv_4096'] = ms.dwt_daub4(v4096, True)
    # time_avg (time_averaging.py -> _time_averaging.c)
    white, redw = pb.teststats_series()
    g['tavg_red'] = np.array(ms_time_avg(R, redw, 100, 5))
    g['tavg_white'] = np.array(ms_time_avg(R, white, 100, 5))
    g['tavg_red_default'] = np.array(ms_time_avg(R, redw, None, 1))
    s2000 = pb.series_case(2000, 5)
    g['tavg_2000'] = np.array(ms_time_avg(R, s2000, 1000, 1))
    s50k = pb.series_case(50000, 8)
    g['tavg_50k'] = np.array(ms_time_avg(R, s50k, 300, 7))
    g['tavg_50k_big'] = np.array(ms_time_avg(R, s50k, 25000, 997))
    # bin_array (stats.py:36-91 -> _binarray.c)
    bd, bu = pb.binarray_case()
    for bs in (100, 7, 4099):
        g[f'bin_unw_{bs}'] = ms.bin_array(bd, bs)
        w = ms.bin_array(bd, bs, bu)
        g[f'bin_w_{bs}'] = np.array(w)
    # gelman_rubin (gelman.py)
    import importlib
    gel = importlib.import_module('mc3.stats.gelman')
    Z, zc, burn = pb.gelman_case()
    g['gelman'] = gel.gelman_rubin(Z, zc, burn)
    # log_prior (stats.py:287-392)
    c = pb.chisq_case()
    post = np.random.RandomState(9).normal(c['params'], 0.3, (50, 6))
    g['log_prior'] = ms.log_prior(post, c['priors'], c['priorlow'],
                                  c['priorup'], np.ones(6))
    np.savez(os.path.join(GOLD, 'kernels.npz'), **g)
    print('kernels.npz:', len(g), 'entries')


def ms_time_avg(R, data, maxbins, binstep):
    import importlib
    tav = importlib.import_module('mc3.stats.time_averaging')
    return tav.time_avg(data, maxbins, binstep)


def run_reference(R, case, sampler):
    """Reference mcmc() with ncpu=1 and both seeds pinned."""
    p = pb.mcmc_case(case)
    func = om.MODELS[p['model']]
    random.randint = lambda a, b: pb.CHILD_SEED     # feeds chain.py:180
    np.random.seed(pb.PARENT_SEED)
    log = R.utils.Log(verb=0)
    return R.mcmc_driver.mcmc(
        p['data'], np.copy(p['uncert']), func, np.copy(p['params']),
        [p['x']], {}, p['pmin'], p['pmax'], p['pstep'],
        p['prior'], p['priorlow'], p['priorup'], p['nchains'], 1,
        p['nsamples'], sampler, p['wlike'], None, False, 0.0, 0.5,
        p['burnin'], p['thinning'], 1.0, p['fepsilon'], 10, 'normal',
        None, False, log, None, None)


def run_oracle(R, case, sampler, use_ref_kernels=True):
    p = pb.mcmc_case(case)
    kw = {}
    if use_ref_kernels:
        kw = dict(chisq_fn=R.stats.chisq, dwt_fn=R.stats.dwt_chisq)
    return omc.mcmc(
        p['data'], p['uncert'], om.MODELS[p['model']], p['params'], [p['x']],
        {}, p['pmin'], p['pmax'], p['pstep'], p['prior'], p['priorlow'],
        p['priorup'], nchains=p['nchains'], nsamples=p['nsamples'],
        sampler=sampler, wlike=p['wlike'], burnin=p['burnin'],
        thinning=p['thinning'], fepsilon=p['fepsilon'],
        parent_seed=pb.PARENT_SEED, child_seed=pb.CHILD_SEED, **kw)


def mcmc_golden(R):
    for case in pb.MCMC_CASES:
        for sampler in pb.SAMPLERS:
            a = run_reference(R, case, sampler)
            b = run_oracle(R, case, sampler)
            for k in ('posterior', 'zchain', 'log_post', 'bestp'):
                assert np.array_equal(a[k], b[k]), (case, sampler, k)
            assert a['best_log_post'] == b['best_log_post']
            assert a['acceptance_rate'] == b['acceptance_rate']
            p = pb.mcmc_case(case)
            fx = {
                'in_checksum': pb.checksum(p['x'], p['data'], p['uncert']),
                'ref_posterior': a['posterior'], 'ref_zchain': a['zchain'],
                'ref_log_post': a['log_post'], 'ref_chisq': a['chisq'],
                'ref_bestp': a['bestp'], 'ref_best_log_post': a['best_log_post'],
                'ref_best_chisq': a['best_chisq'],
                'ref_acceptance_rate': a['acceptance_rate'],
                'ref_best_model': a['best_model'],
                'numaccept': b['numaccept'], 'outbounds': b['outbounds'],
                'Z0': b['Z'][:b['M0']], 'log_post0': b['log_post_full'][:b['M0']],
                'final_freepars': b['freepars'], 'final_chisq': b['chisq_cur'],
                'generations': b['generations'],
            }
            for k, v in b['draws'].items():
                fx['draw_' + k] = v
            np.savez_compressed(
                os.path.join(GOLD, f'mcmc_{case}_{sampler}.npz'), **fx)
            print(f'mcmc_{case}_{sampler}.npz  rows={a["posterior"].shape[0]} '
                  f'acc={a["acceptance_rate"]:.2f}%  '
                  f'oob={b["outbounds"].tolist()}')


def hpd_golden(R):
    """Reference calc_sample_statistics(calc_hpd=True) (stats.py:876-964, scipy KDE)
    on the seeded posterior of problems.hpd_case, at two quantiles."""
    post, bestp, pstep = pb.hpd_case()
    g = {'in_checksum': pb.checksum(post, bestp, pstep)}
    for q in (0.683, 0.9545):
        st = R.stats.calc_sample_statistics(post, bestp, pstep, quantile=q, calc_hpd=True)
        for name, v in zip(('median', 'mean', 'std', 'med_lo', 'med_hi', 'mode', 'hpd_lo', 'hpd_hi'), st):
            g[f'{name}_{q}'] = v
    np.savez(os.path.join(GOLD, 'hpd.npz'), **g)
    print('hpd.npz:', len(g), 'entries')
    # the reference's <root>_statistics.txt for the same sample (stats.py:967-1112);
    # `post` only needs the attributes summary_stats reads
    import types
    ifree = np.where(pstep > 0)[0]
    pdfs = [R.stats.cred_region(post[:, i])[0:2] for i in range(post.shape[1])]
    fake = types.SimpleNamespace(posterior=post, bestp=bestp[ifree], npars=post.shape[1],
                                 pnames=[f'p{i}' for i in range(post.shape[1])],
                                 pdf=[p[0] for p in pdfs], xpdf=[p[1] for p in pdfs])
    out = {'bestp': bestp, 'pstep': pstep, 'pnames': [f'Param {i+1}' for i in range(len(bestp))],
           'texnames': [rf'$\\alpha_{i}$' for i in range(len(bestp))], 'best_chisq': 1234.56789,
           'best_log_post': -620.0, 'BIC': 1290.123456, 'red_chisq': 1.0345678,
           'stddev_residuals': 0.4987654321}
    R.stats.summary_stats(fake, out, filename=os.path.join(GOLD, 'summary_stats.txt'))
    print('summary_stats.txt written')


def savefile_golden(R):
    """A savefile as the REFERENCE writes it (mcmc_driver.py:321-324 / np.savez of the
    output dict): the wire format resume=True must read."""
    import tempfile
    p = pb.mcmc_case('sine')
    func = om.MODELS[p['model']]
    random.randint = lambda a, b: pb.CHILD_SEED
    np.random.seed(pb.PARENT_SEED)
    log = R.utils.Log(verb=0)
    with tempfile.TemporaryDirectory() as d:
        sv = os.path.join(d, 'ref_run.npz')
        out = R.mcmc_driver.mcmc(
            p['data'], np.copy(p['uncert']), func, np.copy(p['params']), [p['x']], {},
            p['pmin'], p['pmax'], p['pstep'], p['prior'], p['priorlow'], p['priorup'],
            p['nchains'], 1, p['nsamples'], 'demc', False, None, True, 0.0, 0.5,
            p['burnin'], p['thinning'], 1.0, p['fepsilon'], 10, 'normal', sv, False, log,
            None, None)
        out['chisq_factor'] = 1.0                  # sampler_driver.py:563-565 adds it before the final save
        np.savez(sv, **out)
        with open(sv, 'rb') as f, open(os.path.join(GOLD, 'ref_savefile_sine_demc.npz'), 'wb') as g:
            g.write(f.read())
    print('ref_savefile_sine_demc.npz: keys', sorted(out.keys()))


def main():
    if not ref.have_ref_py():
        sys.exit('make_golden needs /root/reference (authoring container only)')
    ok.build()
    R = ref.ref_py()
    os.makedirs(GOLD, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == 'hpd':
        hpd_golden(R)
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'savefile':
        savefile_golden(R)
        return
    kernels_golden(R)
    mcmc_golden(R)
    hpd_golden(R)
    savefile_golden(R)


if __name__ == '__main__':
    main()
