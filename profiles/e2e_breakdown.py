"""Where the end-to-end time of mcmc() goes (config 2, K generations)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mc3_b200 as mc3
from mc3_b200 import workloads, mcmc_driver as md, utils as mu, stats as ms
from mc3_b200.engine import Population

K = int(sys.argv[1]) if len(sys.argv) > 1 else 200
w = workloads.config2()
def sync(): torch.cuda.synchronize()
for rep in range(2):
    t = [time.perf_counter()]
    def lap(name):
        sync(); t.append(time.perf_counter()); print(f'  {name:28s} {1e3*(t[-1]-t[-2]):8.2f} ms')
    print('rep', rep)
    pop = Population(w['data'], w['uncert'], mc3.models.sinusoid, w['params'], [w['x']], {},
                     w['pstep'], w['pmin'], w['pmax'], w['prior'], w['priorlow'], w['priorup'],
                     nchains=4096, sampler='demc', fepsilon=0.01, thinning=1, nzchain=K, seed=3)
    lap('Population() + H2D')
    pop.init_population('normal'); lap('init_population')
    pop.run(K); lap(f'run({K}) incl. graph capture')
    c = pop.counters(); lap('counters')
    out = {'burnin': 0}
    n = pop.zsize()
    Z = pop.Z[:n].cpu().numpy(); zc = pop.zchain[:n].cpu().numpy().astype(int); lp = pop.log_post[:n].cpu().numpy(); lap('D2H history')
    zv = zc >= 0
    lpr = ms.log_prior(Z[zv], pop.prior, pop.priorlow, pop.priorup, pop.pstep); lap('log_prior')
    best = md.calc_bestfit_statistics(c['bestp'], pop); lap('bestfit stats')
    post, _, zm = mu.burn(Z=Z[zv], zchain=zc[zv], burnin=0); lap('burn')
    st = ms.calc_sample_statistics(post, c['bestp'], pop.pstep); lap('sample statistics')
    print('  total', 1e3*(t[-1]-t[0]), 'ms')

# ---- the hub call itself under cProfile (second call: warm) -----------------
import cProfile, pstats
def hub():
    return md.mcmc(w['data'], w['uncert'], mc3.models.sinusoid, w['params'], [w['x']], {},
                   w['pmin'], w['pmax'], w['pstep'], w['prior'], w['priorlow'], w['priorup'],
                   4096, None, 4096*K, 'demc', False, None, False, 0.0, 0.5, 0, 1, 1.0, 0.01,
                   10, 'normal', None, False, mc3.Log(verb=-1), None, None, seed=77)
hub(); sync()
t0 = time.perf_counter(); hub(); sync(); print('mcmc() warm wall', 1e3*(time.perf_counter()-t0), 'ms')
pr = cProfile.Profile(); pr.enable(); hub(); sync(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(45)
pstats.Stats(pr).sort_stats('tottime').print_stats(25)
