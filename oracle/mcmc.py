"""ORACLE (test infrastructure only): numpy restatement of the reference's
MCMC loop in its single-process form (ncpu=1), the only form in which the
reference is reproducible (SURVEY.md finding 4).

Follows, statement by statement, the order in which the reference consumes
its legacy numpy random stream, so that with the same two seeds it reproduces
the reference's posterior byte for byte (tests/test_oracle.py pins this against
fixtures made from the real reference by oracle/make_golden.py):

  sizes, initial population, best-of-initial     mc3/mcmc_driver.py:116-134, 229-278
  per-generation loop                            mc3/chain.py:158-299
  model evaluation / chi-squared                 mc3/chain.py:302-340
  output keys                                    mc3/stats/stats.py:805-852

It also records every random draw in a slot-indexed log (the "replay log")
that the CUDA replay mode consumes.
"""
import numpy as np

from . import kernels as ok


class DrawLog:
    """Slot-indexed record of the child's random stream.

    normal[g, nfree]              shared support draw of generation g   chain.py:185
    a[g, j], b[g, j]              snooker: iR1, iR2 (after the ==iR1 -> 0 rule) chain.py:197-200
                                  demc:    r1, r2 (after both != ID rules)       chain.py:223-229
    usj[g, j]                     snooker: uniform compared with 0.1     chain.py:201
    iz[g, j], gs[g, j]            snooker jump: z index, U(1.2, 2.2)     chain.py:203-213
    u[g, j]                       Metropolis uniform; NaN = not drawn    chain.py:257-260
    inb, acc [g, j]               in-bounds / accepted flags
    chisq[g, j]                   proposal chi-squared (NaN if not evaluated)
    ngen                          generations started (the last may be partial)
    """

    def __init__(self, nchains, nfree, sampler):
        self.nchains, self.nfree, self.sampler = nchains, nfree, sampler
        self.rows = []
        self.normal = []

    def start_generation(self, normal):
        n = self.nchains
        self.normal.append(np.array(normal, dtype=float))
        self.rows.append(dict(
            a=np.zeros(n, np.int64), b=np.zeros(n, np.int64),
            usj=np.full(n, np.nan), iz=np.full(n, -1, np.int64),
            gs=np.full(n, np.nan), u=np.full(n, np.nan),
            inb=np.zeros(n, bool), acc=np.zeros(n, bool),
            chisq=np.full(n, np.nan), done=np.zeros(n, bool)))
        return self.rows[-1]

    def arrays(self):
        out = {'normal': np.array(self.normal), 'ngen': len(self.rows)}
        for k in self.rows[0]:
            out[k] = np.array([r[k] for r in self.rows])
        return out


def _eval_chisq(func, params, indparams, indparams_dict, wlike, data, uncert,
                prior, priorlow, priorup, chisq_fn, dwt_fn, ret='chisq'):
    """chain.py:302-340."""
    if wlike:
        model = func(params[0:-3], *indparams, **indparams_dict)
    else:
        model = func(params, *indparams, **indparams_dict)
    if np.any(model == np.inf):
        chisq = np.inf
    elif wlike:
        chisq = dwt_fn(model, data, params, prior, priorlow, priorup)
    else:
        chisq = chisq_fn(model, data, uncert, params, prior, priorlow, priorup)
    if ret == 'both':
        return model, chisq
    return chisq


def mcmc(data, uncert, func, params, indparams=(), indparams_dict=None,
         pmin=None, pmax=None, pstep=None,
         prior=None, priorlow=None, priorup=None,
         nchains=7, nsamples=None, sampler='snooker', wlike=False,
         grtest=False, grbreak=0.0, grnmin=0.5, burnin=0, thinning=1,
         fgamma=1.0, fepsilon=0.0, hsize=10, kickoff='normal',
         parent_seed=None, child_seed=None, record=True,
         chisq_fn=None, dwt_fn=None, max_generations=None):
    """Single-process MCMC exactly as reference mcmc()+Chain.run() with ncpu=1.

    parent_seed seeds numpy before the initial population (the caller of the
    reference does this with np.random.seed); child_seed is what the reference
    child feeds to np.random.seed at chain.py:180.
    """
    indparams_dict = indparams_dict or {}
    chisq_fn = chisq_fn or ok.chisq
    dwt_fn = dwt_fn or ok.dwt_chisq
    data = np.asarray(data, float)
    uncert = np.asarray(uncert, float)
    params = np.array(params, float)
    npars = params.size
    pstep = np.asarray(pstep, float)
    pmin = np.full(npars, -np.inf) if pmin is None else np.asarray(pmin, float)
    pmax = np.full(npars, np.inf) if pmax is None else np.asarray(pmax, float)
    if prior is None or priorlow is None or priorup is None:
        prior = priorlow = priorup = np.zeros(npars)
    prior, priorlow, priorup = (np.asarray(a, float) for a in
                                (prior, priorlow, priorup))

    def ev(p, ret='chisq'):
        return _eval_chisq(func, p, indparams, indparams_dict, wlike, data,
                           uncert, prior, priorlow, priorup, chisq_fn, dwt_fn,
                           ret)

    # mcmc_driver.py:116-134
    ifree = np.where(pstep > 0)[0]
    ishare = np.where(pstep < 0)[0]
    nfree = ifree.size
    M0 = hsize*nchains
    nzchain = int(np.ceil(nsamples/nchains/thinning))
    zlen = M0 + nzchain*nchains
    zburn = int(int(burnin)/thinning)

    freepars = np.zeros((nchains, nfree))
    Z = np.zeros((zlen, nfree))
    log_post = np.zeros(zlen)
    zchain = -np.ones(zlen, int)
    chainsize = np.tile(hsize, nchains)
    outbounds = np.zeros(nfree, int)
    bestp = np.copy(params)
    numaccept = 0

    # mcmc_driver.py:186-198
    if grnmin >= 1:
        grnmin = int(grnmin/thinning)
    elif grnmin > 0:
        grnmin = int(grnmin*nchains*(nzchain - zburn))
    grnmin += int(M0 + zburn*nchains)

    # mcmc_driver.py:229-278 -- initial population, serial, parent stream.
    if parent_seed is not None:
        np.random.seed(parent_seed)
    values = np.copy(params)
    x0, sigma = params[ifree].copy(), pstep[ifree].copy()
    i = j = 0
    while i < M0 and j < 100*M0:
        if kickoff == 'normal':
            trial = np.random.normal(x0, sigma)
        else:
            trial = np.random.uniform(pmin[ifree], pmax[ifree])
        values[ifree] = trial
        if np.any(values > pmax) or np.any(values < pmin):
            j += 1
            continue
        for s in ishare:
            values[s] = values[-int(pstep[s]) - 1]
        lp = -0.5*ev(values)
        if not np.isfinite(lp):
            j += 1
            continue
        Z[i] = values[ifree]
        log_post[i] = lp
        i += 1
    if i < M0 - 1:
        raise ValueError('Cannot populate an initial sample set of parameters')
    izbest = np.argmax(log_post[0:M0])
    best_log_post = log_post[izbest]
    bestp[ifree] = Z[izbest]

    # chain.py:163-180
    IDs = np.arange(nchains)
    index = M0 + IDs
    freepars[:] = Z[IDs]
    chisq = -2.0*log_post[IDs]
    nextp = np.copy(params)
    njump = 0
    zsize = M0
    gamma = fgamma*2.38/np.sqrt(2*nfree)
    chainlen = int(zlen/nchains)
    if child_seed is not None:
        np.random.seed(child_seed)

    log = DrawLog(nchains, nfree, sampler) if record else None
    gr_history = []
    report = (nzchain*nchains)/10
    intsteps = report
    gen = 0
    stop = False
    while not stop:
        if max_generations is not None and gen >= max_generations:
            break
        njump += 1
        normal = np.random.normal(0, pstep[ifree], nfree)
        row = log.start_generation(normal) if record else None
        for jc in range(nchains):
            ID = jc
            mrfactor = 1.0
            sjump = False
            if sampler == 'snooker':
                iR1 = np.random.randint(0, zsize)
                iR2 = np.random.randint(1, zsize)
                if iR2 == iR1:
                    iR2 = 0
                usj = np.random.uniform()
                sjump = usj < 0.1
                if record:
                    row['a'][jc], row['b'][jc], row['usj'][jc] = iR1, iR2, usj
                if sjump:
                    iz = np.random.randint(zsize)
                    z = Z[iz]
                    gs = np.random.uniform(1.2, 2.2)
                    if record:
                        row['iz'][jc], row['gs'][jc] = iz, gs
                    if np.all(z == freepars[ID]):
                        jump = gs*(Z[iR2] - Z[iR1])
                    else:
                        dz = freepars[ID] - z
                        zp1 = np.dot(Z[iR1], dz)
                        zp2 = np.dot(Z[iR2], dz)
                        jump = gs*(zp1 - zp2)*dz/np.dot(dz, dz)
                else:
                    jump = gamma*(Z[iR1] - Z[iR2]) + fepsilon*normal
            elif sampler == 'mrw':
                jump = normal
            elif sampler == 'demc':
                r1 = np.random.randint(1, nchains)
                if r1 == ID:
                    r1 = 0
                r2 = (r1 + np.random.randint(2, nchains)) % nchains
                if r2 == ID:
                    r2 = (r1 + 1) % nchains
                if record:
                    row['a'][jc], row['b'][jc] = r1, r2
                jump = gamma*(freepars[r1] - freepars[r2]) + fepsilon*normal
            else:
                raise ValueError(sampler)

            nextp[ifree] = np.copy(freepars[ID]) + jump
            outpars = ((nextp < pmin) | (nextp > pmax))[ifree]
            if np.any(outpars):
                outbounds += outpars
            else:
                for s in ishare:
                    nextp[s] = nextp[-int(pstep[s]) - 1]
                nextchisq = ev(nextp)
                if sampler == 'snooker' and sjump:
                    cnorm = np.dot(freepars[ID] - z, freepars[ID] - z)
                    nnorm = np.dot(nextp[ifree] - z, nextp[ifree] - z)
                    mrfactor = (nnorm/cnorm)**(0.5*(nfree - 1))
                with np.errstate(all='ignore'):
                    u = np.random.uniform()
                    metro = np.exp(0.5*(chisq[jc] - nextchisq))*mrfactor > u
                if record:
                    row['inb'][jc], row['u'][jc] = True, u
                    row['chisq'][jc], row['acc'][jc] = nextchisq, metro
                if metro:
                    freepars[ID] = np.copy(nextp[ifree])
                    chisq[jc] = nextchisq
                    numaccept += 1
                    if chisq[jc] < -2*best_log_post:
                        bestp[ifree] = freepars[ID]
                        for s in ishare:
                            bestp[s] = bestp[-int(pstep[s]) - 1]
                        best_log_post = -0.5*chisq[jc]
            if record:
                row['done'][jc] = True
            if njump == thinning:
                if zsize == zlen:           # chain.py:279-280
                    stop = True
                    break
                if sampler == 'snooker':
                    index[jc] = zsize
                zsize += 1
                zchain[index[jc]] = ID
                Z[index[jc]] = np.copy(freepars[ID])
                log_post[index[jc]] = -0.5*chisq[jc]
                index[jc] += nchains
                chainsize[ID] += 1
        gen += 1
        if stop:
            break
        if njump == thinning:
            njump = 0
        if sampler in ('mrw', 'demc') and chainsize[0] == chainlen:
            break

        # Hub, checked at generation boundaries (the reference hub runs
        # concurrently, mcmc_driver.py:309-348; used for statistics only).
        if (zsize - M0 >= report) or zsize == zlen:
            report += intsteps
            if grtest and np.all(chainsize > (zburn + hsize)):
                psrf = ok.gelman_rubin(Z, zchain, zburn)
                gr_history.append((zsize, psrf))
                if grbreak > 0.0 and np.all(psrf < grbreak) and zsize > grnmin:
                    break

    zvalid = zchain >= 0
    nsample = np.sum(zvalid)*thinning
    lpr = ok.log_prior(Z[zvalid], prior, priorlow, priorup, pstep) \
        if np.any(zvalid) else np.zeros(0)
    out = {
        'posterior': Z[zvalid], 'zchain': zchain[zvalid],
        'log_post': log_post[zvalid],
        'chisq': -2.0*(log_post[zvalid] - lpr),
        'acceptance_rate': numaccept*100.0/max(nsample, 1),
        'bestp': bestp, 'best_log_post': best_log_post,
        'burnin': zburn, 'ifree': ifree, 'pstep': pstep,
        # raw state (not reference keys):
        'Z': Z, 'zchain_full': zchain, 'log_post_full': log_post,
        'numaccept': numaccept, 'outbounds': outbounds, 'M0': M0,
        'zsize': zsize, 'chainsize': chainsize, 'freepars': freepars,
        'chisq_cur': chisq, 'generations': gen, 'gr_history': gr_history,
    }
    best_model, opt = ev(bestp, 'both')
    out['best_model'] = best_model
    if record:
        out['draws'] = log.arrays()
    return out
