"""GPU parity of the round-2 paths, through the C ABI, against the oracle:
uniform-uncertainty kernels (mc3b_chisq_opts_t.uniform_sigma), the dedicated
uniform-grid kernel of config 2, launch shapes planned for the whole population
(plan_chains: identical bits for any subset), and the Metropolis step fused into
the model kernel's tail (identical bytes to the separate kernels).
fp64 tolerance 1e-10 relative, fp32 1e-5 (BASELINE.json north_star)."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import kernels as ok
from oracle import models as om
from oracle import problems as pb

R64, R32 = 1e-10, 1e-5
IDS = {'polynomial': 0, 'sinusoid': 1, 'gaussian': 2, 'box': 3, 'grid': 4}


@pytest.fixture(scope='module')
def mc3():
    import mc3_b200
    return mc3_b200


def chisq_ex(model_id, P, x, data, uncert, dtype='f64', usig=False, plan_chains=0,
             partial_rows=False):
    """mc3b_model_chisq_ex + sum over splits.  uncert: array, or ONE value with usig."""
    from mc3_b200 import _lib
    dev = torch.device('cuda')
    nb, npars = P.shape
    tdt = torch.float64 if dtype == 'f64' else torch.float32
    code = _lib.F64 if dtype == 'f64' else _lib.F32
    dP = torch.from_numpy(np.ascontiguousarray(P)).to(dev)
    dx = torch.from_numpy(np.ascontiguousarray(x)).to(dev).to(tdt)
    dd = torch.from_numpy(np.ascontiguousarray(data)).to(dev).to(tdt)
    w = np.atleast_1d(1.0/np.asarray(uncert, float))
    dw = torch.from_numpy(w[:1] if usig else w).to(dev).to(tdt)
    ns = ctypes.c_int(0)
    _lib.call('mc3b_model_chisq_plan', max(plan_chains, nb), x.size, code, ctypes.byref(ns))
    part = torch.empty((ns.value, nb), dtype=torch.float64, device=dev)
    o = _lib.ChisqOpts()
    o.plan_chains, o.uniform_sigma = plan_chains, 1 if usig else 0
    _lib.call('mc3b_model_chisq_ex', model_id, code, dP.data_ptr(), npars, nb, npars,
              dx.data_ptr(), dd.data_ptr(), dw.data_ptr(), x.size, part.data_ptr(), nb,
              ns.value, ctypes.byref(o), _lib.stream_ptr())
    torch.cuda.synchronize()
    if partial_rows:
        return part.cpu().numpy()
    acc = torch.zeros(nb, dtype=torch.float64, device=dev)
    for s in range(ns.value):                      # split order, as k_metropolis adds them
        acc += part[s]
    return acc.cpu().numpy()


def _problem(name, n, seed, sigma=0.1):
    rs = np.random.RandomState(seed)
    if name == 'sinusoid':
        x, p0 = np.linspace(0, 10, n), np.array([1.0, 2.5, 0.3, 5.0, -0.2])
        sc = np.array([0.05, 0.02, 0.1, 0.1, 0.01])
    elif name == 'gaussian':
        x, p0 = np.linspace(-5, 5, n), np.array([2.0, 0.3, 1.2, 0.5])
        sc = np.array([0.1, 0.1, 0.05, 0.05])
    elif name == 'box':
        x, p0 = np.linspace(-0.5, 0.5, n), np.array([0.01, 0.0, 0.1, 1.0])
        sc = np.array([1e-3, 1e-2, 1e-2, 1e-3])
    else:
        x, p0, sc = np.linspace(-1, 1, n), rs.normal(0, 1, 3), np.full(3, 0.05)
        name = 'polynomial'
    data = om.MODELS[name](p0, x) + rs.normal(0, sigma, n)
    return name, x, data, p0, sc


@pytest.mark.parametrize('name', ['sinusoid', 'gaussian', 'box', 'poly3'])
@pytest.mark.parametrize('nchains,n', [(1, 37), (40, 4097), (300, 2560), (1000, 5000)])
def test_uniform_sigma_fp64(mc3, name, nchains, n):
    """One uncertainty for all points: chi-squared = sigma^-2 sum (m-d)^2 must equal
    the reference formula sum ((m-d)/sigma_i)^2 (_chisq.c:131-133) at 1e-10."""
    mname, x, data, p0, sc = _problem(name, n, 50 + nchains)
    P = p0 + np.random.RandomState(nchains).normal(0, 1, (nchains, p0.size))*sc
    unc = np.full(n, 0.1)
    got = chisq_ex(IDS[mname], P, x, data, unc, usig=True)
    want = np.array([ok.chisq(om.MODELS[mname](p, x), data, unc) for p in P])
    np.testing.assert_allclose(got, want, rtol=R64)
    # and the general-weights kernel on the same inputs
    got2 = chisq_ex(IDS[mname], P, x, data, unc, usig=False)
    np.testing.assert_allclose(got2, want, rtol=R64)


@pytest.mark.parametrize('name', ['sinusoid', 'gaussian', 'poly3'])
def test_uniform_sigma_fp32(mc3, name):
    mname, x, data, p0, sc = _problem(name, 20000, 9)
    P = p0 + np.random.RandomState(4).normal(0, 1, (256, p0.size))*sc
    unc = np.full(x.size, 0.1)
    got = chisq_ex(IDS[mname], P, x, data, unc, dtype='f32', usig=True)
    want = np.array([ok.chisq(om.MODELS[mname](p, x), data, unc) for p in P])
    np.testing.assert_allclose(got, want, rtol=R32)


def test_grid_usig_accuracy_zero_data(mc3):
    """Dedicated uniform-grid kernel (no abscissa / weight stream, line carried by
    additions, sine by the Reinsch recurrence restarted every two tiles): with
    data = 0 chi-squared is sum(model^2); benign and stiff phases, long series."""
    rs = np.random.RandomState(17)
    for n, span, nch in ((100000, 10.0, 256), (1 << 20, 5000.0, 128), (4099, 1.0, 200)):
        x = np.linspace(-0.3*span, 0.7*span, n)
        P = np.column_stack([rs.uniform(0.5, 2, nch), rs.uniform(0.01, 3.0, nch),
                             rs.uniform(-3, 3, nch), rs.uniform(-1, 1, nch),
                             rs.uniform(-0.1, 0.1, nch)])
        got = chisq_ex(4, P, x, np.zeros(n), 1.0, usig=True)
        plain = chisq_ex(1, P, x, np.zeros(n), np.ones(n))
        want = np.array([np.sum(om.sinusoid(p, x)**2) for p in P[:24]])
        np.testing.assert_allclose(got[:24], want, rtol=5e-13)
        np.testing.assert_allclose(got, plain, rtol=1e-11)


def test_grid_usig_with_data(mc3):
    """Same kernel against data: high S/N, offsets much larger than the noise, a few
    samples per period (guarded chains take the direct evaluation), ragged sizes."""
    rs = np.random.RandomState(23)
    for n, span, snr in ((100000, 10.0, 10.0), (65536 + 77, 300.0, 1e4), (3001, 2.0, 1.0),
                         (128*7, 1.0, 3.0)):
        nch = 160
        x = np.linspace(0.1*span, 1.1*span, n)
        truth = np.array([1.3, 0.37*span/10, 0.4, 25.0, 0.02])
        sig = truth[0]/snr
        data = om.sinusoid(truth, x) + rs.normal(0, 1, n)*sig
        P = truth + rs.normal(0, 1, (nch, 5))*np.array([0.05, 1e-4, 0.05, 0.05, 1e-3])/snr
        if snr <= 10:
            P[::7, 1] = rs.uniform(4, 12, P[::7].shape[0])*(x[1] - x[0])
        got = chisq_ex(4, P, x, data, sig, usig=True)
        want = np.array([np.sum(((om.sinusoid(p, x) - data)/sig)**2) for p in P])
        np.testing.assert_allclose(got, want, rtol=R64)


def test_grid_usig_config2_sample_vs_oracle(mc3):
    """BASELINE config 2 at full size (4096 chains x 1e5 points, sigma = 0.5): a
    sample of chains against the oracle C chi-squared; permuting the chains
    permutes the results bit for bit; run-to-run determinism."""
    from mc3_b200 import workloads
    w = workloads.config2()
    rs = np.random.RandomState(5)
    P = w['params'] + rs.normal(0, 1, (4096, 5))*w['pstep']*3
    got = chisq_ex(4, P, w['x'], w['data'], 0.5, usig=True)
    pick = rs.choice(4096, 12, replace=False)
    want = np.array([ok.chisq(om.sinusoid(p, w['x']), w['data'], w['uncert']) for p in P[pick]])
    np.testing.assert_allclose(got[pick], want, rtol=R64)
    perm = rs.permutation(4096)
    assert np.array_equal(chisq_ex(4, P[perm], w['x'], w['data'], 0.5, usig=True), got[perm])
    assert np.array_equal(chisq_ex(4, P, w['x'], w['data'], 0.5, usig=True), got)


@pytest.mark.parametrize('model_id,usig,nch,n', [(4, True, 4096, 100000), (4, True, 512, 20011),
                                                 (1, True, 1024, 50000 + 37)])
def test_partial_rows_are_the_oracle_over_their_split_usig(mc3, model_id, usig, nch, n):
    from mc3_b200 import _lib
    rs = np.random.RandomState(model_id + n)
    x = np.linspace(0.0, 10.0, n)
    P = np.column_stack([rs.uniform(0.5, 2, nch), rs.uniform(0.3, 3.0, nch), rs.uniform(-3, 3, nch),
                         rs.uniform(-1, 1, nch), rs.uniform(-0.1, 0.1, nch)])
    sig = 0.7
    data = om.sinusoid(P[0], x) + rs.normal(0, sig, n)
    ns = ctypes.c_int(0)
    _lib.call('mc3b_model_chisq_plan', nch, n, _lib.F64, ctypes.byref(ns))
    bounds = (ctypes.c_int64*(ns.value + 1))()
    _lib.call('mc3b_model_chisq_splits', nch, n, _lib.F64, bounds, ns.value + 1, ctypes.byref(ns))
    b = np.array(bounds[:ns.value + 1])
    got = chisq_ex(model_id, P, x, data, sig, usig=usig, partial_rows=True)
    for c in (0, 1, nch//2, nch - 1):
        r2 = ((om.sinusoid(P[c], x) - data)/sig)**2
        want = np.array([r2[b[s]:b[s + 1]].sum() for s in range(ns.value)])
        np.testing.assert_allclose(got[:, c], want, rtol=1e-9, atol=1e-9*want.max())
        np.testing.assert_allclose(got[:, c].sum(), r2.sum(), rtol=R64)


@pytest.mark.parametrize('model_id,usig', [(4, True), (4, False), (1, False), (0, True)])
def test_plan_chains_makes_bits_independent_of_the_launch(mc3, model_id, usig):
    """Planned for the whole population, any subset of chains gets the same bits as
    in the full launch (what a multi-GPU chain partition relies on)."""
    rs = np.random.RandomState(31 + model_id)
    n, nch = 40000 + 11, 2048
    x = np.linspace(0.0, 10.0, n)
    if model_id == 0:
        P = rs.normal(0, 1, (nch, 3))
        data = om.quad(P[0], x) + rs.normal(0, 1, n)
    else:
        P = np.column_stack([rs.uniform(0.5, 2, nch), rs.uniform(0.3, 3.0, nch), rs.uniform(-3, 3, nch),
                             rs.uniform(-1, 1, nch), rs.uniform(-0.1, 0.1, nch)])
        data = om.sinusoid(P[0], x) + rs.normal(0, 1, n)
    unc = np.full(n, 0.8) if usig else rs.uniform(0.5, 1.5, n)
    full = chisq_ex(model_id, P, x, data, unc, usig=usig, plan_chains=nch)
    for lo, hi in ((0, 256), (256, 512), (1024, 2048), (2047, 2048), (100, 133)):
        sub = chisq_ex(model_id, P[lo:hi], x, data, unc, usig=usig, plan_chains=nch)
        assert np.array_equal(sub, full[lo:hi]), (lo, hi)


@pytest.mark.parametrize('sampler', ['demc', 'snooker', 'mrw'])
@pytest.mark.parametrize('case', ['sine', 'quad'])
def test_fused_metropolis_equals_separate_kernels(mc3, sampler, case, monkeypatch):
    """The Metropolis step taken by the last CTA of each chain group inside the
    model kernel (2 launches per generation) leaves exactly the bytes the separate
    k_metropolis + k_advance launches leave: history, log-posterior, counters."""
    from mc3_b200.engine import Population
    p = pb.mcmc_case(case)
    # same model kernel on both sides: the sufficient-statistics form exists only inside the
    # fused launch and rounds differently (tests/test_gpu_moment.py compares it at 1e-11)
    monkeypatch.setenv('MC3B_NO_MOMENT', '1')

    def run(fuse, graph):
        if fuse:
            monkeypatch.delenv('MC3B_NO_FUSE', raising=False)
        else:
            monkeypatch.setenv('MC3B_NO_FUSE', '1')
        pop = Population(p['data'], p['uncert'], mc3.models.BUILTIN[p['model']], p['params'],
                         [p['x']], {}, p['pstep'], p['pmin'], p['pmax'], p['prior'],
                         p['priorlow'], p['priorup'], nchains=384, sampler=sampler,
                         fepsilon=0.01, thinning=2, nzchain=30, seed=11)
        assert pop.fused == fuse
        pop.init_population('normal')
        pop.run(60, use_graph=graph)
        c = pop.counters()
        return (pop.Z.cpu().numpy(), pop.log_post.cpu().numpy(), pop.zchain.cpu().numpy(),
                c['numaccept'], c['outbounds'], c['bestp'], int(pop.gen_dev.item()),
                int(pop.fuse_done.abs().sum().item()))
    ref = run(False, False)
    for graph in (False, True):
        got = run(True, graph)
        for a, b in zip(ref[:6], got[:6]):
            assert np.array_equal(a, b)
        assert got[6] == 60 and got[7] == 0        # generation counter advanced, counters left zero


def test_population_uses_uniform_sigma_and_grid(mc3):
    from mc3_b200.engine import Population
    from mc3_b200 import workloads
    w = workloads.config2(n=20000)
    kw = dict(pstep=w['pstep'], pmin=w['pmin'], pmax=w['pmax'], nchains=128, sampler='demc')
    pop = Population(w['data'], w['uncert'], mc3.models.sinusoid, w['params'], [w['x']], {}, **kw)
    assert pop.grid and pop.usig and pop.fused
    unc = np.array(w['uncert'])
    unc[5] *= 1.0000001
    pop2 = Population(w['data'], unc, mc3.models.sinusoid, w['params'], [w['x']], {}, **kw)
    assert pop2.grid and not pop2.usig
    P = torch.from_numpy(w['params'] + np.zeros((128, 1))).cuda()
    a, b = pop.chisq(P).cpu().numpy(), pop2.chisq(P).cpu().numpy()
    np.testing.assert_allclose(a, b, rtol=1e-9)
    want = ok.chisq(om.sinusoid(w['params'], w['x']), w['data'], w['uncert'])
    np.testing.assert_allclose(a, want, rtol=R64)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_device_keyword_selects_the_gpu(mc3):
    """sample(..., device='cuda:1') while cuda:0 is current runs on GPU 1."""
    p = pb.mcmc_case('sine')
    kw = dict(data=p['data'], uncert=p['uncert'], func=mc3.models.sinusoid,
              params=p['params'], indparams=[p['x']], pstep=p['pstep'], pmin=p['pmin'],
              pmax=p['pmax'], sampler='demc', nchains=128, nsamples=128*20, burnin=2,
              seed=3, log=mc3.Log(verb=-1))
    torch.cuda.set_device(0)
    a = mc3.sample(**kw)
    b = mc3.sample(**kw, device='cuda:1')
    assert torch.cuda.current_device() == 0
    assert np.array_equal(a['posterior'], b['posterior'])


def test_hpd_statistics_on_device_match_reference_golden(mc3):
    """mc3b_hpd (device KDE on 100 points, 3000-point resample, descending running
    sum) against the reference's calc_sample_statistics(calc_hpd=True) on the seeded
    posterior of problems.hpd_case: Gaussian, skewed, bimodal and bounded marginals,
    two quantiles; rtol 1e-6 (scipy's KDE sums in another order)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'hpd.npz'))
    post, bestp, pstep = pb.hpd_case()
    assert str(g['in_checksum']) == pb.checksum(post, bestp, pstep)
    for q in (0.683, 0.9545):
        st = mc3.stats.calc_sample_statistics(post, bestp, pstep, quantile=q, calc_hpd=True, device=True)
        for name, v in zip(('median', 'mean', 'std', 'med_lo', 'med_hi', 'mode', 'hpd_lo', 'hpd_hi'), st):
            np.testing.assert_allclose(v, g[f'{name}_{q}'], rtol=1e-6, atol=1e-9, err_msg=f'{name} q={q}')
    # more samples than the KDE keeps (every n/120000-th): thinning path
    rs = np.random.RandomState(3)
    big = rs.normal(1.0, 2.0, (250000, 2))
    big[:, 1] = np.exp(0.3*big[:, 1])
    mode, lo, hi = mc3.stats.hpd_statistics(big, 0.683)
    for i in range(2):
        pdf, xpdf, hmin = mc3.stats.cred_region(big[:, i], 0.683)
        sel = xpdf[pdf > hmin]
        np.testing.assert_allclose([mode[i], lo[i], hi[i]], [xpdf[np.argmax(pdf)], sel.min(), sel.max()],
                                   rtol=1e-6, atol=1e-9)


def test_resume_from_reference_written_savefile_and_gelman_rubin(mc3, tmp_path):
    """resume=True reads a savefile written by the REFERENCE (tests/golden/
    ref_savefile_sine_demc.npz: np.savez of its output dict, mcmc_driver.py:321-324):
    the old samples come first, every chain restarts from its last row
    (chain.py:166-169), and Gelman-Rubin counts each chain's earlier samples
    (burn-in from the start of the chain, gelman.py:43-55)."""
    import os
    import shutil
    from mc3_b200.mcmc_driver import mcmc
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref_savefile_sine_demc.npz')
    sv = str(tmp_path/'run.npz')
    shutil.copy(gold, sv)
    old = np.load(gold)
    p = pb.mcmc_case('sine')
    out = mcmc(p['data'], p['uncert'], mc3.models.sinusoid, p['params'], [p['x']], {},
               p['pmin'], p['pmax'], p['pstep'], p['prior'], p['priorlow'], p['priorup'],
               8, None, 1600, 'demc', False, None, True, 0.0, 0.5, 40, 2, 1.0, 0.01,
               10, 'normal', sv, True, mc3.Log(verb=-1), None, None, seed=5,
               return_population=True)
    pop = out.pop('_population')
    n_old = old['posterior'].shape[0]
    assert out['posterior'].shape[0] == n_old + 800
    assert np.array_equal(out['posterior'][:n_old], old['posterior'])
    assert np.array_equal(out['zchain'][:n_old], old['zchain'])
    assert np.array_equal(out['log_post'][:n_old], old['log_post'])
    assert np.array_equal(out['zchain'][n_old:], np.tile(np.arange(8), 100))
    # first new row of chain c follows its last old state
    for c in range(8):
        last = old['posterior'][np.where(old['zchain'] == c)[0][-1]]
        first_new = out['posterior'][n_old + c]
        assert np.all(np.abs(first_new - last) < 30*p['pstep'])
    # Gelman-Rubin over old + new samples, burn-in taken from the start of each chain
    got = pop.gelman_rubin(20)
    want = ok.gelman_rubin(out['posterior'], out['zchain'], 20)
    np.testing.assert_allclose(got, want, rtol=1e-10)
    # the file was rewritten with the reference's key set and loads again
    with np.load(sv) as f:
        assert f['posterior'].shape[0] == n_old + 800
        for k in ('posterior', 'zchain', 'chisq', 'log_post', 'acceptance_rate', 'bestp',
                  'best_chisq', 'red_chisq', 'BIC', 'best_log_post', 'best_model',
                  'stddev_residuals', 'burnin', 'pstep', 'ifree'):
            assert k in f.files, k


def test_dwt_chisq_population_at_config3_size(mc3):
    """Wavelet likelihood at BASELINE config 3's data size (N = 2^20, white + red noise
    synthesised by the inverse D4 transform) for a population of 1024 chains: every
    chain against the oracle's restatement of _dwt.c:56-119 at 1e-10, and identical
    chains give identical bits."""
    from mc3_b200 import _lib, workloads
    w = workloads.config3()
    n, nb = w['x'].size, 1024
    rs = np.random.RandomState(8)
    P = np.tile(w['params'], (nb, 1))
    P[:, :4] += rs.normal(0, 1, (nb, 4))*np.array([1e-3, 1e-2, 1e-2, 1e-4])
    P[:, 5:] *= rs.uniform(0.5, 2.0, (nb, 2))
    P[777] = P[3]
    dev = torch.device('cuda')
    dP, dx, dd = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (P, w['x'], w['data']))
    lib = _lib.load()
    ws = torch.empty(max(lib.mc3b_dwt_workspace(nb, n), 8)//8, dtype=torch.float64, device=dev)
    out = torch.empty(nb, dtype=torch.float64, device=dev)
    _lib.call('mc3b_dwt_chisq', mc3.models.box.model_id, dP.data_ptr(), 7, nb, 7, 4,
              dx.data_ptr(), None, 0, dd.data_ptr(), n, ws.data_ptr(), out.data_ptr(),
              _lib.stream_ptr())
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert got[777] == got[3]
    want = np.array([ok.dwt_chisq(om.box(p[:4], w['x']), w['data'], p) for p in P])
    np.testing.assert_allclose(got, want, rtol=R64)
