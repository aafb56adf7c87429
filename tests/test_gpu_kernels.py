"""GPU parity of the CUDA kernels, called through the C ABI, against the oracle
(oracle/) and the reference's golden outputs (tests/golden/kernels.npz).
fp64 tolerance 1e-10 relative, fp32 1e-5 (BASELINE.json north_star)."""
import ctypes
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import kernels as ok
from oracle import models as om
from oracle import problems as pb

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
K = np.load(os.path.join(GOLD, 'kernels.npz'))
R64, R32 = 1e-10, 1e-5


@pytest.fixture(scope='module')
def mc3():
    import mc3_b200
    return mc3_b200


# ---- reference known answers through the drop-in wrappers -------------------
def test_kat_chisq_residuals(mc3):
    ms = mc3.stats
    data = np.array([1.1, 1.2, 0.9, 1.0])
    model, unc = np.ones(4), np.full(4, 0.1)
    np.testing.assert_allclose(ms.chisq(model, data, unc), 6.0, rtol=1e-13)
    args = (np.array([2.5, 5.5]), np.array([2.0, 5.0]), np.array([0.0, 1.0]),
            np.array([0.0, 1.0]))
    np.testing.assert_allclose(ms.chisq(model, data, unc, *args), 6.25, rtol=1e-13)
    np.testing.assert_allclose(ms.residuals(model, data, unc),
                               [-1.0, -2.0, 1.0, 0.0], atol=1e-13)
    np.testing.assert_allclose(ms.residuals(model, data, unc, *args),
                               [-1.0, -2.0, 1.0, 0.0, 0.5], atol=1e-13)


def test_kat_dwt(mc3):
    ms = mc3.stats
    data = np.array([2.0, 0.0, 3.0, -2.0, -1.0, 2.0, 2.0, 0.0])
    params = np.array([1.0, 0.1, 0.1])
    np.testing.assert_allclose(ms.dwt_chisq(np.ones(8), data, params),
                               1693.22308882)
    c = ms.dwt_chisq(np.ones(8), data, params, np.array([1.0, 0.2, 0.3]),
                     np.array([0.0, 0.0, 0.1]), np.array([0.0, 0.0, 0.1]))
    np.testing.assert_allclose(c, 1697.2230888243134, rtol=1e-12)
    with pytest.raises(ValueError, match='at least three parameters'):
        ms.dwt_chisq(np.ones(8), data, params[:2])
    with pytest.raises(ValueError, match='2\\*\\*k'):
        ms.dwt_chisq(np.ones(7), data[:7], params)
    e4 = np.zeros(32)
    e4[4] = 1.0
    DAUB4_FWD, DAUB4_INV = pb.KAT_DAUB4_FWD, pb.KAT_DAUB4_INV
    inv = ms.dwt_daub4(e4, True)
    np.testing.assert_allclose(inv, DAUB4_INV, atol=1e-10)
    np.testing.assert_allclose(ms.dwt_daub4(e4), DAUB4_FWD, atol=1e-10)
    np.testing.assert_allclose(ms.dwt_daub4(inv), e4, atol=1e-8)


def test_kat_bin_array(mc3):
    ms = mc3.stats
    data = np.array([0, 1, 2, 3, 3, 3, 3, 3, 4])
    unc = np.array([3, 1, 1, 1, 2, 3, 2, 2, 4])
    np.testing.assert_allclose(ms.bin_array(data, 3), [1.0, 3.0, 10/3])
    bd, bs = ms.bin_array(data, 3, unc)
    np.testing.assert_allclose(bd, [1.42105263, 3.0, 3.11111111])
    np.testing.assert_allclose(bs, [0.68824720, 0.85714286, 1.33333333])


def test_kat_time_avg(mc3):
    RED_RMS, RED_RMSHI = pb.KAT_RED_RMS, pb.KAT_RED_RMSHI
    white, red = pb.teststats_series()
    rms, lo, hi, err, bsz = mc3.stats.time_avg(red, len(red)/10, 5)
    np.testing.assert_almost_equal(rms, RED_RMS)
    np.testing.assert_almost_equal(hi, RED_RMSHI)
    np.testing.assert_almost_equal(bsz, 1 + 5*np.arange(20))
    assert len(mc3.stats.time_avg(red)[0]) == 500
    assert len(mc3.stats.time_avg(list(red), 500, 2)[0]) == 250


# ---- golden outputs of the reference's C extensions -------------------------
def test_golden_chisq(mc3):
    ms = mc3.stats
    c = pb.chisq_case()
    np.testing.assert_allclose(ms.chisq(c['model'], c['data'], c['uncert']),
                               K['chisq_noprior'], rtol=R64)
    args = (c['params'], c['priors'], c['priorlow'], c['priorup'])
    np.testing.assert_allclose(ms.chisq(c['model'], c['data'], c['uncert'], *args),
                               K['chisq_prior'], rtol=R64)
    np.testing.assert_allclose(ms.residuals(c['model'], c['data'], c['uncert'], *args),
                               K['residuals_prior'], rtol=R64)


@pytest.mark.parametrize('n', [8, 1024, 16384])
def test_golden_dwt_chisq(mc3, n):
    d = pb.dwt_case(n)
    ms = mc3.stats
    np.testing.assert_allclose(ms.dwt_chisq(d['model'], d['data'], d['params']),
                               K[f'dwt_noprior_{n}'], rtol=R64)
    np.testing.assert_allclose(
        ms.dwt_chisq(d['model'], d['data'], d['params'], d['priors'],
                     d['priorlow'], d['priorup']), K[f'dwt_prior_{n}'], rtol=R64)


def test_golden_daub4(mc3):
    rs = np.random.RandomState(3)
    v1000, v4096 = rs.normal(0, 1, 1000), rs.normal(0, 1, 4096)
    for v, tag in ((v1000, '1000'), (v4096, '4096')):
        np.testing.assert_allclose(mc3.stats.dwt_daub4(v), K['daub4_fwd_' + tag],
                                   rtol=R64, atol=1e-12)
        np.testing.assert_allclose(mc3.stats.dwt_daub4(v, True),
                                   K['daub4_inv_' + tag], rtol=R64, atol=1e-12)


@pytest.mark.parametrize('key,args', [
    ('tavg_red', ('red', 100, 5)), ('tavg_white', ('white', 100, 5)),
    ('tavg_red_default', ('red', None, 1)), ('tavg_2000', (2000, 1000, 1)),
    ('tavg_50k', (50000, 300, 7)), ('tavg_50k_big', (50000, 25000, 997))])
def test_golden_time_avg(mc3, key, args):
    src, maxbins, binstep = args
    if src in ('red', 'white'):
        white, red = pb.teststats_series()
        data = red if src == 'red' else white
    else:
        data = pb.series_case(src, {2000: 5, 50000: 8}[src])
    out = np.array(mc3.stats.time_avg(data, maxbins, binstep))
    np.testing.assert_allclose(out, K[key], rtol=R64)


@pytest.mark.parametrize('bs', [100, 7, 4099])
def test_golden_bin_array(mc3, bs):
    d, u = pb.binarray_case()
    np.testing.assert_allclose(mc3.stats.bin_array(d, bs), K[f'bin_unw_{bs}'], rtol=R64)
    np.testing.assert_allclose(np.array(mc3.stats.bin_array(d, bs, u)),
                               K[f'bin_w_{bs}'], rtol=R64)


def test_golden_gelman(mc3):
    Z, zc, burn = pb.gelman_case()
    np.testing.assert_allclose(mc3.stats.gelman_rubin(Z, zc, burn), K['gelman'],
                               rtol=R64)


# ---- fused model + chi-squared kernel vs the oracle --------------------------
def _model_problem(name, n, seed):
    rs = np.random.RandomState(seed)
    if name == 'sinusoid':
        x = np.linspace(0, 10, n)
        p0 = np.array([1.0, 2.5, 0.3, 5.0, -0.2])
        sc = np.array([0.05, 0.02, 0.1, 0.1, 0.01])
    elif name == 'gaussian':
        x = np.linspace(-5, 5, n)
        p0 = np.array([2.0, 0.3, 1.2, 0.5])
        sc = np.array([0.1, 0.1, 0.05, 0.05])
    elif name == 'box':
        x = np.linspace(-0.5, 0.5, n)
        p0 = np.array([0.01, 0.0, 0.1, 1.0])
        sc = np.array([1e-3, 1e-2, 1e-2, 1e-3])
    else:
        deg = int(name[4:])
        x = np.linspace(-1, 1, n)
        p0 = rs.normal(0, 1, deg)
        sc = np.full(deg, 0.05)
        name = 'polynomial'
    uncert = rs.uniform(0.5, 1.5, n)*0.1
    data = om.MODELS[name](p0, x) + rs.normal(0, 1, n)*uncert
    return name, x, data, uncert, p0, sc


def _run_model_chisq(mc3, name, P, x, data, uncert, dtype, prior=None):
    from mc3_b200 import _lib
    dev = torch.device('cuda')
    model = mc3.models.BUILTIN[name]
    nb, npars = P.shape
    dP = torch.from_numpy(P).to(dev)
    tdt = torch.float64 if dtype == 'f64' else torch.float32
    dx = torch.from_numpy(x).to(dev).to(tdt)
    dd = torch.from_numpy(data).to(dev).to(tdt)
    dw = (1.0/torch.from_numpy(uncert).to(dev)).to(tdt)
    code = _lib.F64 if dtype == 'f64' else _lib.F32
    ns = ctypes.c_int(0)
    _lib.call('mc3b_model_chisq_plan', nb, x.size, code, ctypes.byref(ns))
    part = torch.empty((ns.value, nb), dtype=torch.float64, device=dev)
    _lib.call('mc3b_model_chisq', model.model_id, code, dP.data_ptr(), npars, nb,
              model.nmodel(npars), dx.data_ptr(), dd.data_ptr(), dw.data_ptr(),
              x.size, part.data_ptr(), nb, ns.value, _lib.stream_ptr())
    out = torch.empty(nb, dtype=torch.float64, device=dev)
    if prior is None:
        _lib.call('mc3b_chisq_finish', part.data_ptr(), nb, ns.value, nb, None, 0,
                  0, None, None, None, out.data_ptr(), _lib.stream_ptr())
    else:
        pr = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in prior]
        _lib.call('mc3b_chisq_finish', part.data_ptr(), nb, ns.value, nb,
                  dP.data_ptr(), npars, npars, pr[0].data_ptr(), pr[1].data_ptr(),
                  pr[2].data_ptr(), out.data_ptr(), _lib.stream_ptr())
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize('name', ['sinusoid', 'gaussian', 'box', 'poly1', 'poly3',
                                  'poly5', 'poly8'])
@pytest.mark.parametrize('nchains,n', [(1, 37), (7, 1000), (40, 4097), (300, 2560),
                                       (1000, 5000)])
def test_model_chisq_fp64(mc3, name, nchains, n):
    mname, x, data, uncert, p0, sc = _model_problem(name, n, 100 + nchains)
    rs = np.random.RandomState(nchains)
    P = p0 + rs.normal(0, 1, (nchains, p0.size))*sc
    got = _run_model_chisq(mc3, mname, P, x, data, uncert, 'f64')
    want = np.array([ok.chisq(om.MODELS[mname](p, x), data, uncert) for p in P])
    np.testing.assert_allclose(got, want, rtol=R64)


@pytest.mark.parametrize('name', ['sinusoid', 'gaussian', 'box', 'poly3'])
def test_model_chisq_fp32(mc3, name):
    mname, x, data, uncert, p0, sc = _model_problem(name, 20000, 7)
    P = p0 + np.random.RandomState(3).normal(0, 1, (256, p0.size))*sc
    got = _run_model_chisq(mc3, mname, P, x, data, uncert, 'f32')
    want = np.array([ok.chisq(om.MODELS[mname](p, x), data, uncert) for p in P])
    np.testing.assert_allclose(got, want, rtol=R32)


def test_model_chisq_priors_and_unaligned(mc3):
    """Two-sided Gaussian priors (stats.h:102-106) and inputs that are not
    16-byte aligned (the kernel must leave the TMA path)."""
    mname, x, data, uncert, p0, sc = _model_problem('sinusoid', 3001, 5)
    P = p0 + np.random.RandomState(1).normal(0, 1, (33, 5))*sc
    prior = (np.array([0.0, 2.5, 0.0, 5.0, 0.0]), np.array([0.0, 0.1, 0.0, 0.2, 0.0]),
             np.array([0.0, 0.1, 0.0, 0.4, 0.0]))
    got = _run_model_chisq(mc3, mname, P, x, data, uncert, 'f64', prior)
    want = np.array([ok.chisq(om.sinusoid(p, x), data, uncert, p, *prior) for p in P])
    np.testing.assert_allclose(got, want, rtol=R64)
    # odd element offset -> 8-byte aligned only
    got2 = _run_model_chisq(mc3, mname, P, x[1:], data[1:], uncert[1:], 'f64')
    want2 = np.array([ok.chisq(om.sinusoid(p, x[1:]), data[1:], uncert[1:]) for p in P])
    np.testing.assert_allclose(got2, want2, rtol=R64)


def test_model_chisq_full_size_properties(mc3):
    """BASELINE config-2 size (4096 chains x 1e5 points): identical chains give
    identical results, permuting chains permutes results (bit-exact), and a
    sample of chains matches the oracle."""
    mname, x, data, uncert, p0, sc = _model_problem('sinusoid', 100000, 11)
    rs = np.random.RandomState(2)
    P = p0 + rs.normal(0, 1, (4096, 5))*sc
    P[100] = P[7]
    got = _run_model_chisq(mc3, mname, P, x, data, uncert, 'f64')
    assert got[100] == got[7]
    perm = rs.permutation(4096)
    got_p = _run_model_chisq(mc3, mname, P[perm], x, data, uncert, 'f64')
    assert np.array_equal(got_p, got[perm])
    pick = rs.choice(4096, 16, replace=False)
    want = np.array([ok.chisq(om.sinusoid(p, x), data, uncert) for p in P[pick]])
    np.testing.assert_allclose(got[pick], want, rtol=R64)
    # run-to-run determinism
    assert np.array_equal(got, _run_model_chisq(mc3, mname, P, x, data, uncert, 'f64'))


def test_chisq_batch_and_nonfinite(mc3):
    rs = np.random.RandomState(4)
    n = 5003
    data, unc = rs.normal(0, 1, n), rs.uniform(0.5, 2, n)
    models = rs.normal(0, 1, (9, n))
    models[3, 17] = np.inf
    got = mc3.stats.chisq(models, data, unc)
    want = np.array([ok.chisq(m, data, unc) for m in models])
    assert np.isinf(got[3]) and np.isinf(want[3])
    keep = np.arange(9) != 3
    np.testing.assert_allclose(got[keep], want[keep], rtol=R64)


def test_builtin_models_evaluate_on_gpu(mc3):
    x = np.linspace(0, 10, 1001)
    for name, p in (('sinusoid', [1.0, 2.5, 0.3, 5.0, -0.2]), ('gaussian', [2.0, 5.0, 1.2, 0.5]),
                    ('box', [0.01, 5.0, 1.0, 1.0]), ('polynomial', [3.0, -2.4, 0.5])):
        got = mc3.models.BUILTIN[name](np.array(p), x)
        np.testing.assert_allclose(got, om.MODELS[name](np.array(p), x), rtol=1e-13,
                                   atol=1e-13)


# ---- wavelet likelihood with built-in model, batched ------------------------
@pytest.mark.parametrize('n', [64, 512, 2048, 8192, 65536, 1 << 17, 1 << 20])
def test_dwt_chisq_builtin_batched(mc3, n):
    from mc3_b200 import _lib
    dev = torch.device('cuda')
    rs = np.random.RandomState(n)
    x = np.linspace(-0.5, 0.5, n)
    ptrue = np.array([0.01, 0.0, 0.1, 1.0])
    data = om.box(ptrue, x) + rs.normal(0, 1e-3, n)
    nb = 13
    P = np.tile(np.array([0.01, 0.0, 0.1, 1.0, 1.0, 5e-3, 1e-3]), (nb, 1))
    P[:, :4] += rs.normal(0, 1, (nb, 4))*np.array([1e-3, 1e-2, 1e-2, 1e-4])
    P[:, 5:] *= rs.uniform(0.5, 2.0, (nb, 2))
    dP, dx, dd = (torch.from_numpy(a).to(dev) for a in (P, x, data))
    lib = _lib.load()
    ws = torch.empty(max(lib.mc3b_dwt_workspace(nb, n), 8)//8, dtype=torch.float64,
                     device=dev)
    out = torch.empty(nb, dtype=torch.float64, device=dev)
    _lib.call('mc3b_dwt_chisq', mc3.models.box.model_id, dP.data_ptr(), 7, nb, 7, 4,
              dx.data_ptr(), None, 0, dd.data_ptr(), n, ws.data_ptr(),
              out.data_ptr(), _lib.stream_ptr())
    torch.cuda.synchronize()
    want = np.array([ok.dwt_chisq(om.box(p[:4], x), data, p) for p in P])
    np.testing.assert_allclose(out.cpu().numpy(), want, rtol=R64)
    # same through the model-rows entry (user callables)
    rows = np.array([om.box(p[:4], x) for p in P])
    got2 = mc3.stats.dwt_chisq(rows, data, P)
    np.testing.assert_allclose(got2, want, rtol=R64)


def test_daub4_roundtrip_large(mc3):
    v = np.random.RandomState(8).normal(0, 1, 1 << 18)
    f = mc3.stats.dwt_daub4(v)
    np.testing.assert_allclose(f, ok.dwt_daub4(v), rtol=R64, atol=1e-12)
    np.testing.assert_allclose(mc3.stats.dwt_daub4(f, True), v, atol=1e-10)
    # Parseval: the D4 transform is orthogonal
    np.testing.assert_allclose(np.sum(f*f), np.sum(v*v), rtol=1e-12)


# ---- time series at scale: size-independent properties -----------------------
def test_bin_array_large_properties(mc3):
    n = 10_000_019
    rs = np.random.RandomState(6)
    d = rs.normal(1.0, 1.0, n)
    b = mc3.stats.bin_array(d, 100)
    assert b.size == n//100
    np.testing.assert_allclose(b, d[:n//100*100].reshape(-1, 100).mean(axis=1), rtol=1e-12)
    u = np.abs(rs.normal(0, 1, n)) + 0.5
    bw, bs = mc3.stats.bin_array(d, 1000, u)
    w = 1.0/u[:n//1000*1000].reshape(-1, 1000)**2
    np.testing.assert_allclose(bs, np.sqrt(1.0/w.sum(axis=1)), rtol=1e-12)
    np.testing.assert_allclose(
        bw, (d[:n//1000*1000].reshape(-1, 1000)*w).sum(axis=1)/w.sum(axis=1), rtol=1e-11)


def test_time_avg_large_matches_direct(mc3):
    n = 2_000_003
    d = pb.series_case(n, 21)
    rms, lo, hi, err, bsz = mc3.stats.time_avg(d, 1000, 37)
    for i in (0, 1, 13, len(rms) - 1):
        b = int(bsz[i])
        M = n//b
        means = d[:M*b].reshape(M, b).mean(axis=1)
        np.testing.assert_allclose(rms[i], np.sqrt(np.mean(means**2)), rtol=R64)
        np.testing.assert_allclose(err[i], np.std(d)*np.sqrt(M/(b*(M - 1.0))), rtol=R64)
        np.testing.assert_allclose(lo[i], rms[i]/np.sqrt(2.0*M), rtol=R64)
    small = ok.time_avg(d[:200000], 1000, 37)
    got = mc3.stats.time_avg(d[:200000], 1000, 37)
    np.testing.assert_allclose(np.array(got), np.array(small), rtol=R64)


def test_fast_sin_accuracy(mc3):
    """The kernel's branch-free fp64 sine (models.cuh fast_sin) against numpy:
    sinusoid with p = [1, 2 pi, 0, 0, 0] is y = sin(x)."""
    rs = np.random.RandomState(0)
    p = np.array([1.0, 2.0*np.pi, 0.0, 0.0, 0.0])
    xs = np.concatenate([
        np.linspace(-50, 50, 200001), rs.uniform(-1e4, 1e4, 200000),
        rs.uniform(-1e6, 1e6, 100000), np.pi*np.arange(-2000, 2000)/2.0,
        np.array([0.0, 1e-300, -1e-300, 1e-8, 0.5, 1.5707963267948966, 3.141592653589793]),
        rs.uniform(9e8, 1.1e9, 1000), rs.uniform(1e12, 1e15, 1000)])
    got = mc3.models.sinusoid(p, xs)
    want = np.sin(xs)
    small = np.abs(xs) <= 1e6
    assert np.max(np.abs(got - want)[small]) < 6e-16
    assert np.max(np.abs(got - want)) < 2e-15          # library path beyond 1e9
    assert got[xs == 0.0][0] == 0.0
    tiny = np.abs(xs) == 1e-300
    assert np.array_equal(got[tiny], xs[tiny])
    assert np.isnan(mc3.models.sinusoid(p, np.array([np.inf, np.nan, 1.0, 2.0])))[:2].all()


def test_time_avg_tile_and_prefix_paths_agree_with_oracle(mc3):
    """n large enough for the shared-memory tile kernel (bin sizes <= 4096) with
    maxbins beyond it (global-prefix kernel for the rest), odd n, ragged tiles."""
    n = 1_000_003
    d = pb.series_case(n, 33) + 0.7          # non-zero mean
    for maxbins, binstep in ((6000, 499), (1000, 1), (4096, 65), (50000, 4999)):
        got = np.array(mc3.stats.time_avg(d, maxbins, binstep))
        want = np.array(ok.time_avg(d, maxbins, binstep))
        np.testing.assert_allclose(got, want, rtol=R64)
    # unaligned input (8-byte offset): the bulk-copy path must be left
    import torch
    from mc3_b200 import _lib
    dev = torch.device('cuda')
    buf = torch.from_numpy(np.concatenate([[0.0], d])).to(dev)
    view = buf[1:]
    assert view.data_ptr() % 16 == 8
    nout = 10
    outs = [torch.empty(nout, dtype=torch.float64, device=dev) for _ in range(5)]
    lib = _lib.load()
    ws = torch.empty(lib.mc3b_binrms_workspace(n, 1000, 111)//8, dtype=torch.float64, device=dev)
    _lib.call('mc3b_binrms', view.data_ptr(), n, 1000, 111, ws.data_ptr(),
              *[o.data_ptr() for o in outs], _lib.stream_ptr())
    want = np.array(ok.time_avg(d, 1000, 111))
    np.testing.assert_allclose(np.array([o.cpu().numpy() for o in outs]), want, rtol=R64)
    bd = torch.empty(n//100, dtype=torch.float64, device=dev)
    _lib.call('mc3b_binarray', view.data_ptr(), n, 100, None, bd.data_ptr(), None,
              _lib.stream_ptr())
    np.testing.assert_allclose(bd.cpu().numpy(), ok.bin_array(d, 100), rtol=R64)


@pytest.mark.parametrize('bs', [1, 2, 3, 31, 32, 33, 100, 1023, 1024, 1025, 2048])
def test_bin_array_sizes(mc3, bs):
    n = 300_007
    rs = np.random.RandomState(bs)
    d, u = rs.normal(1.0, 1.0, n), np.abs(rs.normal(0, 1, n)) + 0.5
    np.testing.assert_allclose(mc3.stats.bin_array(d, bs), ok.bin_array(d, bs), rtol=R64)
    gw, gs = mc3.stats.bin_array(d, bs, u)
    ww, ws = ok.bin_array(d, bs, u)
    np.testing.assert_allclose(gw, ww, rtol=R64)
    np.testing.assert_allclose(gs, ws, rtol=R64)


def test_sinusoid_grid_recurrence_accuracy(mc3):
    """Uniform-grid sinusoid kernel (rotation recurrence re-anchored per tile):
    with data = 0 and unit weights chi-squared is sum(model^2); it must agree
    with numpy to ~1e-13 for benign and for stiff phases (many radians per
    step, long series), and match the plain kernel's chi-squared at 1e-10."""
    import torch
    from mc3_b200 import _lib
    rs = np.random.RandomState(17)
    for n, span, nch in ((100000, 10.0, 256), (1 << 20, 5000.0, 128), (4099, 1.0, 200)):
        x = np.linspace(-0.3*span, 0.7*span, n)
        P = np.column_stack([rs.uniform(0.5, 2, nch), rs.uniform(0.01, 3.0, nch),
                             rs.uniform(-3, 3, nch), rs.uniform(-1, 1, nch),
                             rs.uniform(-0.1, 0.1, nch)])
        data, unc = np.zeros(n), np.ones(n)
        dev = torch.device('cuda')
        dP, dx, dd, dw = (torch.from_numpy(a).to(dev) for a in (P, x, data, unc))
        outs = {}
        for mid in (_lib.load() and 1, 4):
            ns = ctypes.c_int(0)
            _lib.call('mc3b_model_chisq_plan', nch, n, _lib.F64, ctypes.byref(ns))
            part = torch.empty((ns.value, nch), dtype=torch.float64, device=dev)
            _lib.call('mc3b_model_chisq', mid, _lib.F64, dP.data_ptr(), 5, nch, 5,
                      dx.data_ptr(), dd.data_ptr(), dw.data_ptr(), n, part.data_ptr(),
                      nch, ns.value, _lib.stream_ptr())
            outs[mid] = part.sum(dim=0).cpu().numpy()
        want = np.array([np.sum(om.sinusoid(p, x)**2) for p in P[:24]])
        np.testing.assert_allclose(outs[4][:24], want, rtol=5e-13)
        np.testing.assert_allclose(outs[4], outs[1], rtol=1e-11)


def test_sinusoid_grid_with_data_and_weights(mc3):
    """Same kernel with real data and uncertainties (the residual is formed as
    y/sigma - d/sigma from a pre-scaled data tile): high S/N, offsets much
    larger than the noise, steps on both sides of the pi/2 fold."""
    import torch
    from mc3_b200 import _lib
    rs = np.random.RandomState(23)
    dev = torch.device('cuda')
    for n, span, snr in ((100000, 10.0, 10.0), (65536 + 77, 300.0, 1e4), (3001, 2.0, 1.0)):
        nch = 160
        x = np.linspace(0.1*span, 1.1*span, n)
        truth = np.array([1.3, 0.37*span/10, 0.4, 25.0, 0.02])
        unc = rs.uniform(0.5, 1.5, n)*truth[0]/snr
        data = om.sinusoid(truth, x) + rs.normal(0, 1, n)*unc
        P = truth + rs.normal(0, 1, (nch, 5))*np.array([0.05, 1e-4, 0.05, 0.05, 1e-3])/snr
        if snr <= 10:      # a few samples per period (at S/N 1e4 the rounding of 2 pi x / p1
            P[::7, 1] = rs.uniform(4, 12, P[::7].shape[0])*(x[1] - x[0])   # itself exceeds 1e-10)
        dP, dx, dd, dw = (torch.from_numpy(np.ascontiguousarray(a)).to(dev)
                          for a in (P, x, data, 1.0/unc))
        ns = ctypes.c_int(0)
        _lib.call('mc3b_model_chisq_plan', nch, n, _lib.F64, ctypes.byref(ns))
        part = torch.empty((ns.value, nch), dtype=torch.float64, device=dev)
        _lib.call('mc3b_model_chisq', 4, _lib.F64, dP.data_ptr(), 5, nch, 5,
                  dx.data_ptr(), dd.data_ptr(), dw.data_ptr(), n, part.data_ptr(),
                  nch, ns.value, _lib.stream_ptr())
        got = part.sum(dim=0).cpu().numpy()
        want = np.array([np.sum(((om.sinusoid(p, x) - data)/unc)**2) for p in P])
        np.testing.assert_allclose(got, want, rtol=R64)


@pytest.mark.parametrize('model_id,nch,n', [(4, 4096, 100000), (1, 4096, 100000 + 37),
                                            (0, 1024, 60000), (4, 512, 20011)])
def test_partial_rows_are_the_oracle_over_their_split(mc3, model_id, nch, n):
    """Every row of the partial workspace is the chi-squared over the data range
    mc3b_model_chisq_splits reports for it (decreasing split sizes included)."""
    from mc3_b200 import _lib
    rs = np.random.RandomState(model_id + n)
    dev = torch.device('cuda')
    x = np.linspace(0.0, 10.0, n)
    if model_id == 0:
        P = rs.normal(0, 1, (nch, 3))
        f, nm, npar = om.quad, 3, 3
    else:
        P = np.column_stack([rs.uniform(0.5, 2, nch), rs.uniform(0.3, 3.0, nch), rs.uniform(-3, 3, nch),
                             rs.uniform(-1, 1, nch), rs.uniform(-0.1, 0.1, nch)])
        f, nm, npar = om.sinusoid, 5, 5
    unc = rs.uniform(0.5, 1.5, n)
    data = f(P[0], x) + rs.normal(0, 1, n)*unc
    dP, dx, dd, dw = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (P, x, data, 1.0/unc))
    ns = ctypes.c_int(0)
    _lib.call('mc3b_model_chisq_plan', nch, n, _lib.F64, ctypes.byref(ns))
    bounds = (ctypes.c_int64*(ns.value + 1))()
    _lib.call('mc3b_model_chisq_splits', nch, n, _lib.F64, bounds, ns.value + 1, ctypes.byref(ns))
    b = np.array(bounds[:ns.value + 1])
    part = torch.empty((ns.value, nch), dtype=torch.float64, device=dev)
    _lib.call('mc3b_model_chisq', model_id, _lib.F64, dP.data_ptr(), npar, nch, nm, dx.data_ptr(),
              dd.data_ptr(), dw.data_ptr(), n, part.data_ptr(), nch, ns.value, _lib.stream_ptr())
    got = part.cpu().numpy()
    for c in (0, 1, nch//2, nch - 1):
        r2 = ((f(P[c], x) - data)/unc)**2
        want = np.array([r2[b[s]:b[s + 1]].sum() for s in range(ns.value)])
        np.testing.assert_allclose(got[:, c], want, rtol=1e-9, atol=1e-9*want.max())
        np.testing.assert_allclose(got[:, c].sum(), r2.sum(), rtol=R64)


def test_population_picks_grid_kernel_only_for_uniform_x(mc3):
    from mc3_b200.engine import Population
    p = pb.mcmc_case('sine')
    kw = dict(pstep=p['pstep'], pmin=p['pmin'], pmax=p['pmax'], nchains=128, sampler='demc')
    pop = Population(p['data'], p['uncert'], mc3.models.sinusoid, p['params'], [p['x']], {}, **kw)
    assert pop.grid
    xj = p['x'] + 1e-9*np.sin(np.arange(p['x'].size))
    pop2 = Population(p['data'], p['uncert'], mc3.models.sinusoid, p['params'], [xj], {}, **kw)
    assert not pop2.grid


# ---- the mirrored-pair form of the uniform-grid kernel (k_sinefold) ----------------
def _fold_np(d):
    """include/mc3b200.h mc3b_fold_data, in numpy."""
    nb = d.size//16
    b = d[:nb*16].reshape(nb, 16)
    lo, hi = b[:, 7::-1], b[:, 8:]
    out = np.empty((nb, 16))
    out[:, 0::2] = -0.5*(hi + lo)
    out[:, 1::2] = -0.5*(hi - lo)
    return out.ravel()


def _run_grid_usig(P, x, data, sigma, folded, rows=False, work=False):
    from mc3_b200 import _lib
    dev = torch.device('cuda')
    nch, n = P.shape[0], x.size
    dP, dx, dd = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (P, x, data))
    dw = torch.tensor([1.0/sigma], dtype=torch.float64, device=dev)
    o = _lib.ChisqOpts()
    o.uniform_sigma = 1
    if folded:
        df = torch.full((n,), float('nan'), dtype=torch.float64, device=dev)
        _lib.call('mc3b_fold_data', dd.data_ptr(), n, df.data_ptr(), _lib.stream_ptr())
        o.folded = df.data_ptr()
        if work:
            wk = torch.empty((_lib.FOLD_WORK, nch), dtype=torch.float64, device=dev)
            o.work = wk.data_ptr()
    ns = ctypes.c_int(0)
    _lib.call('mc3b_model_chisq_plan', nch, n, _lib.F64, ctypes.byref(ns))
    part = torch.empty((ns.value, nch), dtype=torch.float64, device=dev)
    _lib.call('mc3b_model_chisq_ex', 4, _lib.F64, dP.data_ptr(), 5, nch, 5, dx.data_ptr(),
              dd.data_ptr(), dw.data_ptr(), n, part.data_ptr(), nch, ns.value,
              ctypes.byref(o), _lib.stream_ptr())
    torch.cuda.synchronize()
    if rows:
        return part.cpu().numpy()
    return part.sum(dim=0).cpu().numpy()


def test_fold_data_is_the_documented_layout(mc3):
    from mc3_b200 import _lib
    dev = torch.device('cuda')
    for n in (16, 100, 4099, 100000):
        d = np.random.RandomState(n).normal(3, 2, n)
        dd = torch.from_numpy(d).to(dev)
        out = torch.full((n,), 7.0, dtype=torch.float64, device=dev)
        _lib.call('mc3b_fold_data', dd.data_ptr(), n, out.data_ptr(), _lib.stream_ptr())
        got = out.cpu().numpy()
        nb = n//16*16
        assert np.array_equal(got[:nb], _fold_np(d))
        assert np.all(got[nb:] == 7.0)


@pytest.mark.parametrize('n,span,offset,snr', [(100000, 10.0, 5.0, 2.0), (65536 + 77, 300.0, 25.0, 1e4),
                                               (3001, 2.0, 5e4, 1.0), (1 << 20, 5000.0, -3.0, 30.0),
                                               (130, 1.0, 1.0, 5.0)])
def test_sinusoid_folded_kernel_matches_oracle(mc3, n, span, offset, snr):
    """k_sinefold (point pairs mirrored about block centres, one uncertainty for all
    points) against the per-point evaluation: benign and short periods, offsets far
    above the noise, high S/N, ragged tails; and against k_sinegrid at 1e-11."""
    rs = np.random.RandomState(n % 1000 + 5)
    nch = 192
    x = np.linspace(0.1*span, 1.1*span, n)
    truth = np.array([1.3, 0.37*span/10, 0.4, offset, 0.02])
    sigma = truth[0]/snr
    data = om.sinusoid(truth, x) + rs.normal(0, sigma, n)
    P = truth + rs.normal(0, 1, (nch, 5))*np.array([0.05, 1e-4, 0.05, 0.05, 1e-3])/snr
    if snr <= 10:
        P[::7, 1] = rs.uniform(2.2, 12, P[::7].shape[0])*(x[1] - x[0])      # a few samples per period
        P[1::7, 1] = rs.uniform(0.5, 3, P[1::7].shape[0])*span                # less than a period in all
    got = _run_grid_usig(P, x, data, sigma, True)
    want = np.array([np.sum(((om.sinusoid(p, x) - data)/sigma)**2) for p in P])
    np.testing.assert_allclose(got, want, rtol=R64)
    ref = _run_grid_usig(P, x, data, sigma, False)
    # (at S/N 1e4 an error of 1e-15 A in the sine is already 1e-11 sigma: both kernels sit
    # ~1e-11 from the oracle there, on different sides)
    np.testing.assert_allclose(got, ref, rtol=1e-11 if snr <= 100 else R64)
    # run-to-run determinism; constants derived once per chain (opts.work) or by every CTA:
    # same bits; and the library-sine guard for huge arguments
    assert np.array_equal(got, _run_grid_usig(P, x, data, sigma, True))
    assert np.array_equal(got, _run_grid_usig(P, x, data, sigma, True, work=True))
    Pbig = P[:40].copy()
    Pbig[:, 1] = 1e-9*(x[1] - x[0])*rs.uniform(1, 2, 40)
    m = min(n, 2000)                 # (the argument's own rounding is ~1e-3 rad here: kernel against kernel)
    gb = _run_grid_usig(Pbig, x[:m], data[:m], sigma, True)
    wb = _run_grid_usig(Pbig, x[:m], data[:m], sigma, False)
    np.testing.assert_allclose(gb, wb, rtol=1e-11)


def test_folded_partial_rows_are_the_oracle_over_their_split(mc3):
    from mc3_b200 import _lib
    rs = np.random.RandomState(77)
    nch, n, sigma = 4096, 100000 + 37, 0.5
    x = np.linspace(0.0, 10.0, n)
    P = np.column_stack([rs.uniform(0.5, 2, nch), rs.uniform(0.3, 3.0, nch), rs.uniform(-3, 3, nch),
                         rs.uniform(-1, 1, nch), rs.uniform(-0.1, 0.1, nch)])
    data = om.sinusoid(P[0], x) + rs.normal(0, sigma, n)
    got = _run_grid_usig(P, x, data, sigma, True, rows=True)
    ns = ctypes.c_int(0)
    bounds = (ctypes.c_int64*(got.shape[0] + 1))()
    _lib.call('mc3b_model_chisq_splits', nch, n, _lib.F64, bounds, got.shape[0] + 1, ctypes.byref(ns))
    b = np.array(bounds[:ns.value + 1])
    for c in (0, 1, nch//2, nch - 1):
        r2 = ((om.sinusoid(P[c], x) - data)/sigma)**2
        want = np.array([r2[b[s]:b[s + 1]].sum() for s in range(ns.value)])
        np.testing.assert_allclose(got[:, c], want, rtol=1e-9, atol=1e-9*want.max())
        np.testing.assert_allclose(got[:, c].sum(), r2.sum(), rtol=R64)


def test_population_uses_the_folded_kernel_for_one_uncertainty(mc3, monkeypatch):
    """Population prepares the paired copy of the data when the grid kernel applies and
    the uncertainties are all equal; the run is the same Markov chain as with the
    per-point kernel up to the rounding of chi-squared (same decisions on this case)."""
    from mc3_b200.engine import Population
    from mc3_b200 import workloads
    w = workloads.config2(n=20000)
    kw = dict(pstep=w['pstep'], pmin=w['pmin'], pmax=w['pmax'], prior=w['prior'], priorlow=w['priorlow'],
              priorup=w['priorup'], nchains=256, sampler='demc', fepsilon=w['fepsilon'], nzchain=12, seed=3)
    pop = Population(w['data'], w['uncert'], mc3.models.sinusoid, w['params'], [w['x']], {}, **kw)
    assert pop.grid and pop.usig and pop.d_fold is not None
    pop.init_population('normal')
    pop.run(12)
    monkeypatch.setenv('MC3B_NO_FOLD', '1')
    pop2 = Population(w['data'], w['uncert'], mc3.models.sinusoid, w['params'], [w['x']], {}, **kw)
    assert pop2.d_fold is None
    pop2.init_population('normal')
    pop2.run(12)
    torch.cuda.synchronize()
    assert torch.equal(pop.zchain, pop2.zchain)
    np.testing.assert_allclose(pop.log_post.cpu().numpy(), pop2.log_post.cpu().numpy(), rtol=1e-11)
    np.testing.assert_allclose(pop.Z.cpu().numpy(), pop2.Z.cpu().numpy(), rtol=0, atol=0)
    unc = w['uncert'].copy()
    unc[5] *= 1.5
    pop3 = Population(w['data'], unc, mc3.models.sinusoid, w['params'], [w['x']], {}, **kw)
    assert pop3.grid and not pop3.usig and pop3.d_fold is None


def test_config4_full_size_properties(mc3):
    """BASELINE config 4 at its stated size (1e8 points): time_avg (1000 bin sizes) and
    bin_array against direct numpy sums on a sample of bin sizes -- size-independent
    properties, since the oracle's O(N x sizes) loop does not finish in seconds here."""
    n = 100_000_000
    rs = np.random.RandomState(44)
    d = rs.standard_normal(n)
    d += 0.05*np.sin(np.arange(n)*1e-4)                   # a little red noise
    rms, lo, hi, err, bsz = mc3.stats.time_avg(d, 1000, 1)
    assert len(rms) == len(bsz) and np.all(np.diff(bsz) >= 0)
    sd = np.std(d)
    for i in (0, 1, 57, len(rms)//2, len(rms) - 2, len(rms) - 1):
        b = int(bsz[i])
        M = n//b
        means = d[:M*b].reshape(M, b).mean(axis=1)
        np.testing.assert_allclose(rms[i], np.sqrt(np.mean(means**2)), rtol=R64)
        np.testing.assert_allclose(err[i], sd*np.sqrt(M/(b*(M - 1.0))), rtol=R64)
    b = mc3.stats.bin_array(d, 100)
    np.testing.assert_allclose(b, d.reshape(-1, 100).mean(axis=1), rtol=1e-12, atol=1e-15)
    u = 0.5 + np.abs(d)
    bw, bs = mc3.stats.bin_array(d, 100, u)
    w = 1.0/u.reshape(-1, 100)**2
    np.testing.assert_allclose(bs, np.sqrt(1.0/w.sum(axis=1)), rtol=1e-12)
    np.testing.assert_allclose(bw, (d.reshape(-1, 100)*w).sum(axis=1)/w.sum(axis=1), rtol=1e-11, atol=1e-14)
