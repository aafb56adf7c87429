import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cProfile, pstats
import mc3_b200 as mc3
from mc3_b200.engine import Population
from oracle import problems as pb
p = pb.mcmc_case('quad')
for rep in range(2):
    pop = Population(p['data'], p['uncert'], mc3.models.polynomial, p['params'], [p['x']], {}, p['pstep'],
                     nchains=7, sampler='snooker', thinning=1, nzchain=14286, seed=3)
    pop.init_population('normal'); torch.cuda.synchronize()
    t0 = time.perf_counter(); pop.run(14286); torch.cuda.synchronize(); t1 = time.perf_counter()
    print('run(14286) small kernel:', 1e3*(t1-t0), 'ms ->', 1e6*(t1-t0)/14286, 'us/gen')
q = mc3.Log(verb=-1)
def call():
    return mc3.sample(p['data'], p['uncert'], func=mc3.models.polynomial, params=p['params'], indparams=[p['x']],
                      pstep=p['pstep'], sampler='snooker', nsamples=1e5, burnin=1000, nchains=7, seed=3, log=q)
call()
pr = cProfile.Profile(); pr.enable(); call(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
