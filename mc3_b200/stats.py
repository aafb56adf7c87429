"""mc3.stats look-alike: numpy in, numpy out, same names / argument order /
error behaviour as the reference wrappers, with the arithmetic done by the CUDA
library through the C ABI (no CPU fallback).

    chisq, residuals        mc3/stats/stats.py:94-216   -> mc3b_chisq_batch / mc3b_residuals
    dwt_chisq, dwt_daub4    mc3/stats/stats.py:219-284, 577-611 -> mc3b_dwt_chisq / mc3b_daub4
    bin_array               mc3/stats/stats.py:36-91    -> mc3b_binarray
    time_avg                mc3/stats/time_averaging.py:17-60 -> mc3b_binrms
    gelman_rubin            mc3/stats/gelman.py:12-61   -> mc3b_gelman_rubin

Host-side post-processing kept in numpy/scipy (not hot):
    log_prior, cred_region, marginal_statistics, calc_sample_statistics.
"""
import numpy as np
import torch

from . import _lib

__all__ = ['bin_array', 'residuals', 'chisq', 'dwt_chisq', 'dwt_daub4',
           'time_avg', 'gelman_rubin', 'log_prior', 'cred_region',
           'marginal_statistics', 'calc_sample_statistics', 'hpd_statistics',
           'summary_stats']


def _dev():
    if not torch.cuda.is_available():
        raise _lib.Mc3bError('mc3_b200 needs a CUDA device (no CPU fallback)')
    return torch.device('cuda', torch.cuda.current_device())


def _up(a, dtype=np.double):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).to(_dev())


def _prior_terms(params, priors, priorlow, priorup):
    """stats.py:208-216: Gaussian priors are those with low > 0 and up > 0."""
    if params is None or priors is None or priorlow is None or priorup is None:
        return None
    params, priors = np.asarray(params, float), np.asarray(priors, float)
    priorlow, priorup = np.asarray(priorlow, float), np.asarray(priorup, float)
    ip = (priorlow > 0) & (priorup > 0)
    return (params - priors)[ip], priorlow[ip], priorup[ip]


def _finish(part, nb, params, priors, priorlow, priorup):
    """sum of one-row partials + prior terms via mc3b_chisq_finish."""
    out = torch.empty(nb, dtype=torch.float64, device=part.device)
    st = _lib.stream_ptr()
    if params is None or priors is None or priorlow is None or priorup is None:
        _lib.call('mc3b_chisq_finish', part.data_ptr(), nb, 1, nb, None, 0, 0,
                  None, None, None, out.data_ptr(), st)
    else:
        P = _up(np.atleast_2d(np.asarray(params, float)))
        if P.shape[0] != nb:
            P = P.expand(nb, -1).contiguous()
        pr, lo, up = _up(priors), _up(priorlow), _up(priorup)
        _lib.call('mc3b_chisq_finish', part.data_ptr(), nb, 1, nb, P.data_ptr(),
                  _lib.ld(P), P.shape[1], pr.data_ptr(), lo.data_ptr(),
                  up.data_ptr(), out.data_ptr(), st)
    return out


def chisq(model, data, uncert, params=None, priors=None, priorlow=None,
          priorup=None):
    """sum(((model-data)/uncert)^2) + Gaussian-prior terms.  `model` may be 2-D
    [nchains, N] (then params may be [nchains, npars]) -> array of chisq."""
    m = _up(np.atleast_2d(model))
    d, u = _up(data), _up(uncert)
    nb, n = m.shape
    part = torch.empty(nb, dtype=torch.float64, device=m.device)
    _lib.call('mc3b_chisq_batch', m.data_ptr(), _lib.ld(m), nb, d.data_ptr(),
              u.data_ptr(), n, part.data_ptr(), _lib.stream_ptr())
    out = _finish(part, nb, params, priors, priorlow, priorup).cpu().numpy()
    return float(out[0]) if np.ndim(model) == 1 else out


def residuals(model, data, uncert, params=None, priors=None, priorlow=None,
              priorup=None):
    m, d, u = _up(model), _up(data), _up(uncert)
    n = m.numel()
    pt = _prior_terms(params, priors, priorlow, priorup)
    npr = 0 if pt is None else pt[0].size
    out = torch.empty(n + npr, dtype=torch.float64, device=m.device)
    if npr:
        off, lo, up = (_up(a) for a in pt)
        args = (off.data_ptr(), lo.data_ptr(), up.data_ptr(), npr)
    else:
        args = (None, None, None, 0)
    _lib.call('mc3b_residuals', m.data_ptr(), d.data_ptr(), u.data_ptr(), n,
              *args, out.data_ptr(), _lib.stream_ptr())
    return out.cpu().numpy()


def _is_pow2(n):
    return n >= 4 and (n & (n - 1)) == 0


def dwt_chisq(model, data, params, priors=None, priorlow=None, priorup=None):
    """Carter & Winn (2009) wavelet pseudo chi-squared.  N must be 2^k >= 4: the
    reference's result for other sizes is undefined."""
    if np.shape(params)[-1] < 3:
        raise ValueError('Wavelet chisq should have at least three parameters')
    m = _up(np.atleast_2d(model))
    d = _up(data)
    nb, n = m.shape
    if not _is_pow2(n):
        raise ValueError(
            f'dwt_chisq needs a data size of the form 2**k >= 4, got {n}')
    P = _up(np.atleast_2d(np.asarray(params, float)))
    if P.shape[0] != nb:
        P = P.expand(nb, -1).contiguous()
    lib = _lib.load()
    ws = torch.empty(max(lib.mc3b_dwt_workspace(nb, n), 8)//8,
                     dtype=torch.float64, device=m.device)
    part = torch.empty(nb, dtype=torch.float64, device=m.device)
    _lib.call('mc3b_dwt_chisq', -1, P.data_ptr(), _lib.ld(P), nb, P.shape[1], 0,
              None, m.data_ptr(), _lib.ld(m), d.data_ptr(), n, ws.data_ptr(),
              part.data_ptr(), _lib.stream_ptr())
    out = _finish(part, nb, params if priors is not None else None, priors,
                  priorlow, priorup).cpu().numpy()
    return float(out[0]) if np.ndim(model) == 1 else out


def dwt_daub4(array, inverse=False):
    """Daubechies-4 DWT (or inverse) of a 1-D array, zero-padded to 2^M."""
    a = np.asarray(array, dtype=np.double)
    n2 = 1 << int(np.ceil(np.log2(a.size)))
    pad = np.zeros(n2)
    pad[:a.size] = a
    if n2 < 4:
        return pad
    src = _up(pad)
    out = torch.empty_like(src)
    ws = torch.empty(n2, dtype=torch.float64, device=src.device)
    _lib.call('mc3b_daub4', src.data_ptr(), n2, -1 if inverse else 1,
              ws.data_ptr(), out.data_ptr(), _lib.stream_ptr())
    return out.cpu().numpy()


def bin_array(data, binsize, uncert=None):
    d = _up(data)
    binsize = int(binsize)
    nb = d.numel()//binsize
    bd = torch.empty(nb, dtype=torch.float64, device=d.device)
    if uncert is None:
        _lib.call('mc3b_binarray', d.data_ptr(), d.numel(), binsize, None,
                  bd.data_ptr(), None, _lib.stream_ptr())
        return bd.cpu().numpy()
    u = _up(uncert)
    bs = torch.empty(nb, dtype=torch.float64, device=d.device)
    _lib.call('mc3b_binarray', d.data_ptr(), d.numel(), binsize, u.data_ptr(),
              bd.data_ptr(), bs.data_ptr(), _lib.stream_ptr())
    return [bd.cpu().numpy(), bs.cpu().numpy()]


def time_avg(data, maxbins=None, binstep=1):
    """rms-vs-bin-size curve: returns [rms, rmslo, rmshi, stderr, binsz]."""
    if isinstance(data, (list, tuple)):
        data = np.array(data)
    if maxbins is None:
        maxbins = len(data)//2
    maxbins, binstep = int(maxbins), int(binstep)
    d = _up(data)
    nout = (maxbins - 1)//binstep + 1
    outs = [torch.empty(nout, dtype=torch.float64, device=d.device)
            for _ in range(5)]
    lib = _lib.load()
    ws = torch.empty(max(lib.mc3b_binrms_workspace(d.numel(), maxbins, binstep), 8)//8,
                     dtype=torch.float64, device=d.device)
    _lib.call('mc3b_binrms', d.data_ptr(), d.numel(), maxbins, binstep,
              ws.data_ptr(), *[o.data_ptr() for o in outs], _lib.stream_ptr())
    return [o.cpu().numpy() for o in outs]


def gelman_rubin(Z, Zchain, burnin):
    """PSRF per parameter (gelman.py:36-61).  The row-index table that orders
    each chain's samples is bookkeeping done on the host; the statistics run on
    the device."""
    Z = np.asarray(Z, dtype=np.double)
    Zchain = np.asarray(Zchain)
    nchains = int(np.amax(Zchain)) + 1
    npars = Z.shape[1]
    order = np.argsort(Zchain, kind='stable')
    zs = Zchain[order]
    starts = np.searchsorted(zs, np.arange(nchains), side='left')
    ends = np.searchsorted(zs, np.arange(nchains), side='right')
    niter = int(np.amin(ends - starts)) - int(burnin)
    if niter < 1:
        print('Not enough samples for Gelman-Rubin test.')
        return np.zeros(npars)
    rows = np.stack([order[s + burnin:s + burnin + niter] for s in starts])
    dZ, drows = _up(Z), _up(rows, np.int64)
    work = torch.empty((2, nchains, npars), dtype=torch.float64, device=dZ.device)
    psrf = torch.empty(npars, dtype=torch.float64, device=dZ.device)
    _lib.call('mc3b_gelman_rubin', dZ.data_ptr(), npars, nchains, 0,
              drows.data_ptr(), niter, 0, niter, work.data_ptr(), psrf.data_ptr(),
              _lib.stream_ptr())
    return psrf.cpu().numpy()


# ---------------------------------------------------------------------------
# Host-side post-processing (numpy / scipy)
# ---------------------------------------------------------------------------
def log_prior(posterior, prior, priorlow, priorup, pstep):
    """-0.5 * sum of squared, width-scaled prior offsets of the free parameters
    (stats.py:367-392); 2 log(p) terms where priorlow < 0."""
    post = np.atleast_2d(np.asarray(posterior, dtype=float))
    ifree = np.where(np.asarray(pstep) > 0)[0]
    terms = np.zeros_like(post)
    for i, k in enumerate(ifree):
        if priorlow[k] > 0 and priorup[k] > 0:
            d = post[:, i] - prior[k]
            terms[:, i] = np.where(d < 0, d/priorlow[k],
                                   np.where(d > 0, d/priorup[k], d))
        elif priorlow[k] < 0:
            terms[:, i] = 2.0*np.log(post[:, i])
    logp = -0.5*np.sum(terms**2, axis=1)
    return logp[0] if logp.size == 1 else logp


def cred_region(posterior=None, quantile=0.6827, pdf=None, xpdf=None):
    """Highest-posterior-density region from a Gaussian-KDE trace of the
    marginal (stats.py:433-467): KDE on <= 120k samples, 100-point grid inside
    +-6 sigma, linear resample to 3000 points, density threshold at `quantile`."""
    import scipy.interpolate as si
    import scipy.stats as ss
    if pdf is None and xpdf is None:
        thin = max(1, int(np.size(posterior)/120000))
        kde = ss.gaussian_kde(posterior[::thin])
        mu, sd = np.mean(posterior), np.std(posterior)
        lo = max(mu - 6*sd, np.amin(posterior))
        hi = min(mu + 6*sd, np.amax(posterior))
        x = np.linspace(lo, hi, 100)
        xpdf = np.linspace(lo, hi, 3000)
        pdf = si.interp1d(x, kde.evaluate(x))(xpdf)
    if quantile is None:
        return pdf, xpdf, 0.0
    ip = np.argsort(pdf)[::-1]
    cdf = np.cumsum(pdf[ip])
    ihpd = np.where(cdf >= quantile*cdf[-1])[0][0]
    return pdf, xpdf, np.amin(pdf[ip][0:ihpd])


def marginal_statistics(posterior, statistics='med_central', quantile=0.683,
                        pdf=None, xpdf=None):
    """Per-parameter estimate and interval (stats.py:764-802)."""
    nsamples, npars = np.shape(posterior)
    values = np.tile(np.nan, npars)
    low, high = np.tile(np.nan, npars), np.tile(np.nan, npars)
    if statistics is None:
        return values, low, high
    if pdf is None or xpdf is None:
        pdf, xpdf = [None]*npars, [None]*npars
    if statistics.startswith('med_'):
        values = np.median(posterior, axis=0)
    elif statistics.startswith('max_'):
        for i in range(npars):
            pdf[i], xpdf[i], _ = cred_region(posterior[:, i], quantile, pdf[i], xpdf[i])
            values[i] = xpdf[i][np.argmax(pdf[i])]
    if quantile is None:
        return values, low, high
    if statistics.endswith('_central'):
        low = np.percentile(posterior, 100*0.5*(1 - quantile), axis=0)
        high = np.percentile(posterior, 100*0.5*(1 + quantile), axis=0)
    elif statistics.endswith('_like'):
        for i in range(npars):
            pdf[i], xpdf[i], hmin = cred_region(posterior[:, i], quantile, pdf[i], xpdf[i])
            sel = xpdf[i][pdf[i] > hmin]
            low[i], high[i] = np.amin(sel), np.amax(sel)
    return values, low, high


def hpd_statistics(posterior, quantile=0.683):
    """mode, hpd_low, hpd_high of every column of posterior [n, nfree] on the GPU
    (mc3b_hpd): the reference's cred_region + marginal_statistics('max_like',
    stats.py:433-467, 764-802) -- Gaussian KDE on 100 points, linear resample to
    3000, density threshold at `quantile`."""
    post = np.ascontiguousarray(posterior, dtype=np.double)
    n, nfree = post.shape
    dP = _up(post)
    lib = _lib.load()
    ws = torch.empty(max(lib.mc3b_hpd_workspace(nfree), 8)//8, dtype=torch.float64,
                     device=dP.device)
    out = torch.empty((3, nfree), dtype=torch.float64, device=dP.device)
    _lib.call('mc3b_hpd', dP.data_ptr(), n, nfree, float(quantile), ws.data_ptr(),
              out.data_ptr(), _lib.stream_ptr())
    res = out.cpu().numpy()
    return res[0], res[1], res[2]


def expand_free_stats(free_stats, bestp, pstep):
    """Spread per-free-parameter statistics (median, mean, std, low, high[, mode,
    hpd_low, hpd_high]) over the full parameter vector: fixed parameters keep
    bestp (std 0), shared ones copy their source (stats.py:908-964)."""
    pstep = np.asarray(pstep)
    ifree = np.where(pstep > 0)[0]
    ishare = np.where(pstep < 0)[0]
    out = []
    for idx, vals in enumerate(free_stats):
        full = np.zeros(len(pstep)) if idx == 2 else np.copy(bestp).astype(float)
        full[ifree] = vals
        for i in ishare:
            full[i] = full[-int(pstep[i]) - 1]
        out.append(full)
    return tuple(out)


def calc_sample_statistics(posterior, bestp, pstep, quantile=0.683,
                           calc_hpd=False, pdf=None, xpdf=None, device=False):
    """median, mean, std, central bounds (+ mode and HPD bounds), expanded to
    the full parameter vector: fixed parameters keep bestp with zero std,
    shared ones copy their source (stats.py:876-964).  device=True takes the
    HPD statistics from the GPU (hpd_statistics) instead of scipy."""
    pstep = np.asarray(pstep)
    npars = len(pstep)
    ifree = np.where(pstep > 0)[0]
    ishare = np.where(pstep < 0)[0]

    def expand(free_vals, fill=None):
        out = np.copy(bestp).astype(float) if fill is None else np.full(npars, fill)
        out[ifree] = free_vals
        for i in ishare:
            out[i] = out[-int(pstep[i]) - 1]
        return out

    med, mlo, mhi = marginal_statistics(posterior, 'med_central', quantile)
    res = [expand(med), expand(np.mean(posterior, axis=0)),
           expand(np.std(posterior, axis=0), fill=0.0), expand(mlo), expand(mhi)]
    if calc_hpd:
        if device and pdf is None and xpdf is None:
            mode, hlo, hhi = hpd_statistics(posterior, quantile)
        else:
            mode, hlo, hhi = marginal_statistics(posterior, 'max_like', quantile,
                                                 pdf=pdf, xpdf=xpdf)
        res += [expand(mode), expand(hlo), expand(hhi)]
    return tuple(res)


def summary_stats(posterior, mc3_output, filename=None, device=True):
    """The reference's `<root>_statistics.txt` (mc3/stats/stats.py:967-1112): medians,
    means, best fit and marginal modes, standard deviations, 1/2-sigma central and
    highest-posterior-density intervals (machine readable, then LaTeX), and the fit
    statistics.  `posterior` is the burned (and thinned to <= 20000 rows) sample of
    the free parameters; HPD statistics come from the GPU unless device=False."""
    import sys
    from . import utils as mu
    bestp, pstep = mc3_output['bestp'], np.asarray(mc3_output['pstep'])
    pnames, texnames = mc3_output['pnames'], mc3_output['texnames']
    npars = len(bestp)
    s1 = calc_sample_statistics(posterior, bestp, pstep, 0.683, calc_hpd=True, device=device)
    s2 = calc_sample_statistics(posterior, bestp, pstep, 0.9545, calc_hpd=True, device=device)
    median, mean, std, mode = s1[0], s1[1], s1[2], s1[5]
    cols = '2sigma_low     1sigma_low     1sigma_up      2sigma_up      Parameter'
    lines = ['Summary of posterior statistics:', '', 'Parameter estimates:',
             ' Median         Mean           Max-posterior  Mode           Parameter']
    lines += [f'{median[i]:14.7e} {mean[i]:14.7e} {bestp[i]:14.7e} {mode[i]:14.7e}  {pnames[i]}'
              for i in range(npars)]
    lines += ['', ' Std_deviation  Parameter']
    lines += [f'{std[i]:14.7e}  {pnames[i]}' for i in range(npars)]
    for title, lo2, lo1, hi1, hi2 in (
            ('Central quantile credible intervals:', s2[3], s1[3], s1[4], s2[4]),
            ('Highest-posterior-density credible intervals:', s2[6], s1[6], s1[7], s2[7])):
        lines += ['', title, ' ' + cols]
        lines += [f'{lo2[i]:14.7e} {lo1[i]:14.7e} {hi1[i]:14.7e} {hi2[i]:14.7e}  {pnames[i]}'
                  for i in range(npars)]
    lines += ['', '', 'LaTeX format']
    for k, (title, val, lo, hi) in enumerate((
            ('Median and 1sigma central-quantile statistics', median, s1[3], s1[4]),
            ('Median and 2sigma central-quantile statistics', median, s2[3], s2[4]),
            ('Marginal max_posterior (mode) and 1sigma-HPD statistics', mode, s1[6], s1[7]),
            ('Marginal max_posterior (mode) and 2sigma-HPD statistics', mode, s2[6], s2[7]))):
        tex = mu.tex_parameters(val, lo, hi, significant_digits=2)
        lines += ([] if k == 0 else ['']) + [title]
        lines += [f'{texnames[i]}  &  {tex[i]}' for i in range(npars)]
    bic = mc3_output['BIC']
    fmt = len(f'{bic:.4f}')
    lines += ['', '',
              f"Best-parameter's chi-squared:       {mc3_output['best_chisq']:{fmt}.4f}",
              f"Best-parameter's -2*log(posterior): {-2.0*mc3_output['best_log_post']:{fmt}.4f}",
              f"Bayesian Information Criterion:     {bic:{fmt}.4f}",
              f"Reduced chi-squared:                {mc3_output['red_chisq']:{fmt}.4f}",
              f"Standard deviation of residuals:  {mc3_output['stddev_residuals']:.6g}", '', '', '']
    text = '\n'.join(lines)
    if filename is None:
        sys.stdout.write(text)
    else:
        with open(filename, 'w') as f:
            f.write(text)
    return text
