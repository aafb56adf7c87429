"""GPU parity tests of the sufficient-statistics form of the uniform-grid sinusoid kernel
(csrc/chisq_grid.cu k_sinefold<MOM>, include/mc3b200.h mc3b_moment_t) against the
per-point evaluation the reference performs (src_c/_chisq.c:111-140 on the model values
of the numpy sinusoid).  All calls go through the C ABI."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'profiles'))
from oracle import kernels as ok            # noqa: E402
from oracle import models as om             # noqa: E402

pytestmark = pytest.mark.gpu
R64 = 1e-10


@pytest.fixture(scope='module')
def mc3():
    import mc3_b200
    return mc3_b200


def test_moment_prepare_matches_numpy(mc3):
    from mc3_b200 import _lib
    import moment_error as me
    dev = torch.device('cuda')
    rs = np.random.RandomState(2)
    for n in (128, 1000, 100000):
        x0, dx = 0.3, 1e-4
        d = 5.0 - 0.2*(x0 + dx*np.arange(n)) + rs.normal(0, 0.5, n)
        c0r, slr = 4.9, -0.19
        f, mom, _ = me.prepare(d, x0, dx, c0r, slr)
        dd = torch.from_numpy(d).to(dev)
        df = torch.full((n,), 7.0, dtype=torch.float64, device=dev)
        dt = torch.zeros((n//128, 4), dtype=torch.float64, device=dev)
        _lib.call('mc3b_moment_prepare', dd.data_ptr(), n//128, x0, dx, None, c0r, slr, 0, df.data_ptr(),
                  dt.data_ptr(), _lib.stream_ptr())
        # the same tiles addressed through their origins (piecewise-uniform layout)
        tx = torch.from_numpy(x0 + 128*dx*np.arange(n//128)).to(dev)
        df2, dt2 = torch.full_like(df, 7.0), torch.zeros_like(dt)
        _lib.call('mc3b_moment_prepare', dd.data_ptr(), n//128, 0.0, dx, tx.data_ptr(), c0r, slr, 0,
                  df2.data_ptr(), dt2.data_ptr(), _lib.stream_ptr())
        np.testing.assert_allclose(df2.cpu().numpy(), df.cpu().numpy(), rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(dt2.cpu().numpy(), dt.cpu().numpy(), rtol=1e-10, atol=1e-10)
        # layout 1 (mma fragments): the same numbers, entry 64 eo + 32 (p / 4) + 4 b + p % 4 of each tile
        df3, dt3 = torch.full_like(df, 7.0), torch.zeros_like(dt)
        _lib.call('mc3b_moment_prepare', dd.data_ptr(), n//128, x0, dx, None, c0r, slr, 1,
                  df3.data_ptr(), dt3.data_ptr(), _lib.stream_ptr())
        l0 = df.cpu().numpy()[:n//128*128].reshape(-1, 8, 8, 2)            # tile, block, pair, (e, o)
        l1 = df3.cpu().numpy()[:n//128*128].reshape(-1, 2, 2, 8, 4)        # tile, eo, ks, block, q
        assert np.array_equal(l1, l0.reshape(-1, 8, 2, 4, 2).transpose(0, 4, 2, 1, 3))
        assert np.array_equal(dt3.cpu().numpy(), dt.cpu().numpy())
        nt = n//128*128
        np.testing.assert_allclose(df.cpu().numpy()[:nt], f, rtol=1e-13, atol=1e-14)
        assert np.all(df.cpu().numpy()[nt:] == 7.0)
        got = dt.cpu().numpy()
        want = np.column_stack([-2*mom[:, 0], -2*mom[:, 1], mom[:, 2], np.zeros(n//128)])
        np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-11)


def _population(mc3, n, snr, offset, nchains=256, seed=5, **kw):
    from mc3_b200.engine import Population
    rs = np.random.RandomState(seed)
    x = np.linspace(0.0, 10.0, n)
    truth = np.array([1.0, 2.5, 0.3, offset, -0.2])
    sigma = truth[0]/snr
    data = om.sinusoid(truth, x) + rs.normal(0, sigma, n)
    step = np.array([1e-2, 1e-3, 1e-2, 1e-2, 1e-3])/max(1.0, snr/2)
    pop = Population(data, np.full(n, sigma), mc3.models.sinusoid, truth*(1 + 1e-3/snr), [x], {},
                     pstep=step, pmin=np.array([0.0, 1.0, -np.pi, -1e9, -10.0]),
                     pmax=np.array([50.0, 5.0, np.pi, 1e9, 10.0]),
                     prior=np.array([0.0, 2.5, 0.0, 0.0, 0.0]), priorlow=np.array([0.0, 0.1, 0.0, 0.0, 0.0]),
                     priorup=np.array([0.0, 0.2, 0.0, 0.0, 0.0]),
                     nchains=nchains, sampler='demc', fepsilon=0.01, nzchain=40, seed=seed, **kw)
    return pop, x, data, sigma


def _one_generation(pop, gen):
    """host-driven generation `gen`; returns (proposals, data chi-squared from the partial rows,
    in-bounds flags, chi-squared of every chain before and after)"""
    before = pop.chisq_cur.clone()
    pop._generation(gen)
    torch.cuda.synchronize()
    part = pop._last_part
    return (pop.nextp.cpu().numpy(), part.sum(dim=0).cpu().numpy(), pop.inb.cpu().numpy() != 0,
            before.cpu().numpy(), pop.chisq_cur.cpu().numpy())


@pytest.mark.parametrize('n,snr,offset', [(100000, 2.0, 5.0), (20000 + 77, 20.0, 5e3), (4096, 0.5, -3.0)])
def test_moment_kernel_matches_oracle(mc3, n, snr, offset):
    """Low and moderate signal-to-noise: the expansion is accurate (no chain trips the
    guard); its partial rows sum to the oracle's chi-squared and the Metropolis step
    used exactly that value."""
    pop, x, data, sigma = _population(mc3, n, snr, offset)
    assert pop.use_moment
    pop.init_population('normal')
    pop.gen_dev.fill_(0)
    for gen in range(3):
        P, got, inb, before, after = _one_generation(pop, gen)
        want = np.array([np.sum(((om.sinusoid(p, x) - data)/sigma)**2) for p in P])
        np.testing.assert_allclose(got[inb], want[inb], rtol=R64)
        prior = ((P[:, 1] - 2.5)/np.where(P[:, 1] > 2.5, 0.2, 0.1))**2
        moved = after != before
        assert moved.any()
        np.testing.assert_allclose(after[moved], (want + prior)[moved], rtol=R64)
    assert int(pop.guard_hits.item()) == 0


def test_moment_guard_reevaluates_chains_point_by_point(mc3, monkeypatch):
    """High signal-to-noise: chains near the mode amplify the rounding of the expansion
    beyond the contract; the guard must catch every one of them (the accepted values
    equal the oracle's to 1e-10), count them, and the policy must leave the moment form."""
    monkeypatch.setenv('MC3B_MOMENT_AMP', '4000')
    pop, x, data, sigma = _population(mc3, 20000, 3000.0, 7.0, nchains=128)
    assert pop.moment is not None
    pop.init_population('normal')
    pop.use_moment = True                    # (the a-priori estimate would have turned it off: see below)
    pop.gen_dev.fill_(0)
    hits = 0
    for gen in range(3):
        P, got, inb, before, after = _one_generation(pop, gen)
        want = np.array([np.sum(((om.sinusoid(p, x) - data)/sigma)**2) for p in P])
        prior = ((P[:, 1] - 2.5)/np.where(P[:, 1] > 2.5, 0.2, 0.1))**2
        moved = after != before
        np.testing.assert_allclose(after[moved], (want + prior)[moved], rtol=R64)
        n_hits = int(pop.guard_hits.item())
        assert n_hits > hits
        hits = n_hits
    # the policy: a fresh population of the same problem never starts in the moment form ...
    pop2, *_ = _population(mc3, 20000, 3000.0, 7.0, nchains=128)
    pop2.init_population('normal')
    pop2.run(2)
    assert not pop2.use_moment
    # ... and one that is forced into it leaves it after the guard has fired
    pop.gen = 3
    for _ in range(4):
        pop.run(1, use_graph=False)
    torch.cuda.synchronize()
    assert not pop.use_moment


def test_population_moment_and_pair_kernels_walk_the_same_chain(mc3, monkeypatch):
    from mc3_b200 import workloads
    from mc3_b200.engine import Population
    w = workloads.config2(n=30000)
    kw = dict(pstep=w['pstep'], pmin=w['pmin'], pmax=w['pmax'], prior=w['prior'], priorlow=w['priorlow'],
              priorup=w['priorup'], nchains=512, sampler='demc', fepsilon=w['fepsilon'], nzchain=30, seed=11)
    runs = []
    for env in (None, 'MC3B_NO_MOMENT'):
        if env:
            monkeypatch.setenv(env, '1')
        pop = Population(w['data'], w['uncert'], mc3.models.sinusoid, w['params'], [w['x']], {}, **kw)
        assert pop.use_moment == (env is None)
        pop.init_population('normal')
        pop.run(30)
        torch.cuda.synchronize()
        runs.append((pop.zchain.cpu().numpy(), pop.log_post.cpu().numpy(), pop.Z.cpu().numpy(),
                     pop.counters()['numaccept']))
        assert pop.use_moment == (env is None)
    assert np.array_equal(runs[0][0], runs[1][0])
    assert runs[0][3] == runs[1][3]
    np.testing.assert_allclose(runs[0][1], runs[1][1], rtol=1e-11)
    assert np.array_equal(runs[0][2], runs[1][2])


def test_moment_run_is_deterministic_and_graph_equals_eager(mc3):
    from mc3_b200 import workloads
    from mc3_b200.engine import Population
    w = workloads.config2(n=20000)
    kw = dict(pstep=w['pstep'], pmin=w['pmin'], pmax=w['pmax'], prior=w['prior'], priorlow=w['priorlow'],
              priorup=w['priorup'], nchains=256, sampler='demc', fepsilon=w['fepsilon'], nzchain=16, seed=4)
    outs = []
    for graph in (True, True, False):
        pop = Population(w['data'], w['uncert'], mc3.models.sinusoid, w['params'], [w['x']], {}, **kw)
        pop.init_population('normal')
        for _ in range(4):
            pop.run(4, use_graph=graph)
        torch.cuda.synchronize()
        assert pop.use_moment
        outs.append((pop.Z.cpu().numpy(), pop.log_post.cpu().numpy()))
    for o in outs[1:]:
        assert np.array_equal(o[0], outs[0][0]) and np.array_equal(o[1], outs[0][1])


# ---- piecewise-uniform abscissa: a constant cadence with gaps ----------------------
def _gapped(n, seed):
    rs = np.random.RandomState(seed)
    x = np.linspace(0.0, 10.0, n)
    keep = np.ones(n, dtype=bool)
    for lo, w in ((n//5, n//33), (n//2, 10), (3*n//4, 1), (n - 700, 300)):
        keep[lo:lo + w] = False
    x = x[keep]
    x[x > 6.0] += 0.37*(x[1] - x[0])                 # the cadence resumes off the original grid
    return x, rs


def test_pair_kernel_on_a_gapped_series_matches_oracle(mc3):
    """k_sinefold with tile origins (opts.tile_x): whole tiles of the runs first, the
    points that fill no tile shared out over the splits; against the per-point
    evaluation and against the plain sinusoid kernel on the same arrays."""
    from mc3_b200 import _lib, gridseg
    dev = torch.device('cuda')
    for n, nch in ((100000, 256), (20000, 160)):
        x, rs = _gapped(n, n)
        L = gridseg.tile_layout(x)
        assert L is not None and L['nleft'] > 0 and L['starts'].size >= 10
        truth = np.array([1.0, 2.5, 0.3, 5.0, -0.2])
        sigma = 0.5
        data = om.sinusoid(truth, x) + rs.normal(0, sigma, x.size)
        P = truth*(1 + 0.02*rs.standard_normal((nch, 5)))
        P[::9, 1] = rs.uniform(2.2, 12, P[::9].shape[0])*L['dx']
        xp, dp = x[L['perm']], data[L['perm']]
        dP, dx_, dd = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (P, xp, dp))
        dw = torch.tensor([1.0/sigma], dtype=torch.float64, device=dev)
        tx = torch.from_numpy(np.ascontiguousarray(x[L['starts']])).to(dev)
        nt = L['starts'].size
        df = torch.zeros_like(dd)
        _lib.call('mc3b_fold_data', dd.data_ptr(), 128*nt, df.data_ptr(), _lib.stream_ptr())
        ns = ctypes.c_int(0)
        _lib.call('mc3b_model_chisq_plan', nch, x.size, _lib.F64, ctypes.byref(ns))
        outs = []
        for work in (False, True):
            o = _lib.ChisqOpts()
            o.uniform_sigma = 1
            o.folded, o.tile_x, o.dx, o.ntiles = df.data_ptr(), tx.data_ptr(), L['dx'], nt
            if work:
                wk = torch.empty((_lib.FOLD_WORK, nch), dtype=torch.float64, device=dev)
                o.work = wk.data_ptr()
            part = torch.empty((ns.value, nch), dtype=torch.float64, device=dev)
            _lib.call('mc3b_model_chisq_ex', 4, _lib.F64, dP.data_ptr(), 5, nch, 5, dx_.data_ptr(),
                      dd.data_ptr(), dw.data_ptr(), x.size, part.data_ptr(), nch, ns.value,
                      ctypes.byref(o), _lib.stream_ptr())
            outs.append(part.sum(dim=0).cpu().numpy())
        want = np.array([np.sum(((om.sinusoid(p, x) - data)/sigma)**2) for p in P])
        np.testing.assert_allclose(outs[0], want, rtol=R64)
        assert np.array_equal(outs[0], outs[1])


def test_population_on_a_gapped_series(mc3, monkeypatch):
    """Population finds the piecewise-uniform layout, its chi-squared equals the oracle's
    on the original arrays, and the run (moment form inside the loop) walks the same
    chain as the plain sinusoid kernel (MC3B_NO_SEG=1)."""
    from mc3_b200.engine import Population
    x, rs = _gapped(40000, 8)
    truth = np.array([1.0, 2.5, 0.3, 5.0, -0.2])
    sigma = 0.5
    data = om.sinusoid(truth, x) + rs.normal(0, sigma, x.size)
    kw = dict(pstep=np.array([1e-2, 1e-3, 1e-2, 1e-2, 1e-3]), pmin=np.array([0.0, 1.0, -np.pi, 0.0, -1.0]),
              pmax=np.array([5.0, 5.0, np.pi, 10.0, 1.0]), nchains=256, sampler='demc', fepsilon=0.01,
              nzchain=20, seed=21)
    runs = []
    for env in (None, 'MC3B_NO_SEG'):
        if env:
            monkeypatch.setenv(env, '1')
        pop = Population(data, np.full(x.size, sigma), mc3.models.sinusoid, truth*1.001, [x], {}, **kw)
        assert (pop.seg is not None) == (env is None) and pop.grid == (env is None)
        assert pop.use_moment == (env is None)
        P = truth*(1 + 0.01*np.random.RandomState(1).standard_normal((128, 5)))
        got = pop.chisq(torch.as_tensor(P, device=pop.dev)).cpu().numpy()
        want = np.array([np.sum(((om.sinusoid(p, x) - data)/sigma)**2) for p in P])
        np.testing.assert_allclose(got, want, rtol=R64)
        pop.init_population('normal')
        pop.run(20)
        torch.cuda.synchronize()
        runs.append((pop.zchain.cpu().numpy(), pop.log_post.cpu().numpy(), pop.Z.cpu().numpy()))
        if env is None:
            assert int(pop.guard_hits.item()) == 0
    assert np.array_equal(runs[0][0], runs[1][0])
    np.testing.assert_allclose(runs[0][1], runs[1][1], rtol=1e-10)
    assert np.array_equal(runs[0][2], runs[1][2])


def test_population_on_a_gapped_series_with_per_point_uncertainties(mc3, monkeypatch):
    """k_sinegrid<USIG=false> with tile origins: weights travel with the reordered points."""
    from mc3_b200.engine import Population
    x, rs = _gapped(40000, 9)
    truth = np.array([1.0, 2.5, 0.3, 5.0, -0.2])
    unc = rs.uniform(0.3, 0.8, x.size)
    data = om.sinusoid(truth, x) + rs.normal(0, 1, x.size)*unc
    kw = dict(pstep=np.array([1e-2, 1e-3, 1e-2, 1e-2, 1e-3]), pmin=np.array([0.0, 1.0, -np.pi, 0.0, -1.0]),
              pmax=np.array([5.0, 5.0, np.pi, 10.0, 1.0]), nchains=256, sampler='demc', fepsilon=0.01,
              nzchain=12, seed=5)
    runs = []
    for env in (None, 'MC3B_NO_SEG'):
        if env:
            monkeypatch.setenv(env, '1')
        pop = Population(data, unc, mc3.models.sinusoid, truth*1.001, [x], {}, **kw)
        assert (pop.seg is not None) == (env is None) and not pop.usig and pop.d_fold is None
        P = truth*(1 + 0.01*np.random.RandomState(2).standard_normal((128, 5)))
        P[::5, 1] = np.random.RandomState(3).uniform(3, 12, P[::5].shape[0])*(x[1] - x[0])
        got = pop.chisq(torch.as_tensor(P, device=pop.dev)).cpu().numpy()
        want = np.array([ok.chisq(om.sinusoid(p, x), data, unc) for p in P])
        np.testing.assert_allclose(got, want, rtol=R64)
        pop.init_population('normal')
        pop.run(12)
        torch.cuda.synchronize()
        runs.append((pop.zchain.cpu().numpy(), pop.log_post.cpu().numpy(), pop.Z.cpu().numpy()))
    assert np.array_equal(runs[0][0], runs[1][0])
    np.testing.assert_allclose(runs[0][1], runs[1][1], rtol=1e-10)
    assert np.array_equal(runs[0][2], runs[1][2])


@pytest.mark.parametrize('snr', [2.0, 3000.0])
def test_unfused_moment_form_with_guarded_finish(mc3, snr):
    """Population.chisq() (initial population, statistics, data shard): k_sinefold<MOM> without
    the Metropolis epilogue, rows summed by mc3b_moment_finish -- which must re-evaluate what
    the expansion cannot deliver (high S/N) and leave the rest alone (low S/N)."""
    pop, x, data, sigma = _population(mc3, 30000 + 41, snr, 5.0, nchains=256)
    assert pop.use_moment
    rs = np.random.RandomState(4)
    truth = np.array([1.0, 2.5, 0.3, 5.0, -0.2])
    P = truth*(1 + 0.01/snr*rs.standard_normal((300, 5)))
    P[:20] = truth*(1 + 0.3*rs.standard_normal((20, 5)))          # far from the mode as well
    P[:, 1] = np.clip(P[:, 1], 1.2, 4.5)
    h0 = int(pop.guard_hits.item())
    got = pop.chisq(torch.as_tensor(P, device=pop.dev)).cpu().numpy()
    want = np.array([np.sum(((om.sinusoid(p, x) - data)/sigma)**2) for p in P])
    want += ((P[:, 1] - 2.5)/np.where(P[:, 1] > 2.5, 0.2, 0.1))**2
    np.testing.assert_allclose(got, want, rtol=R64)
    hits = int(pop.guard_hits.item()) - h0
    assert hits == 0 if snr < 10 else hits >= 250


def test_tensor_core_form_of_the_moment_kernel(mc3, monkeypatch):
    """MC3B_MOM_LAYOUT=1: k_sinemma (Pe, Po by mma.m8n8k4 on the B-fragment layout of the
    folded data) against the oracle, fused and unfused, low S/N and through the guard."""
    monkeypatch.setenv('MC3B_MOM_LAYOUT', '1')
    for n, snr in ((100000, 2.0), (20000 + 77, 3000.0), (1000, 1.0)):
        pop, x, data, sigma = _population(mc3, n, snr, 5.0)
        assert pop.use_moment and pop.moment.layout == 1
        pop.init_population('normal')                    # unfused launches + mc3b_moment_finish
        pop.use_moment = True
        pop.gen_dev.fill_(0)
        for gen in range(2):
            P, got, inb, before, after = _one_generation(pop, gen)
            want = np.array([np.sum(((om.sinusoid(p, x) - data)/sigma)**2) for p in P])
            prior = ((P[:, 1] - 2.5)/np.where(P[:, 1] > 2.5, 0.2, 0.1))**2
            if snr < 10:
                np.testing.assert_allclose(got[inb], want[inb], rtol=R64)
            moved = after != before
            assert moved.any()
            np.testing.assert_allclose(after[moved], (want + prior)[moved], rtol=R64)
        assert (int(pop.guard_hits.item()) == 0) == (snr < 10)
        rs = np.random.RandomState(0)
        Pq = P[rs.choice(P.shape[0], 64, replace=False)]
        got = pop.chisq(torch.as_tensor(Pq, device=pop.dev)).cpu().numpy()
        want = np.array([np.sum(((om.sinusoid(p, x) - data)/sigma)**2) for p in Pq])
        want += ((Pq[:, 1] - 2.5)/np.where(Pq[:, 1] > 2.5, 0.2, 0.1))**2
        np.testing.assert_allclose(got, want, rtol=R64)
