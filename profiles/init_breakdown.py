"""Host-side cost of Population() + init_population() + update_output for config 2 (second call: warm)."""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mc3_b200 as mc3
from mc3_b200 import workloads
from mc3_b200.engine import Population

w = workloads.config2()
def make():
    return Population(w['data'], w['uncert'], mc3.models.sinusoid, w['params'], [w['x']], {},
                      w['pstep'], w['pmin'], w['pmax'], w['prior'], w['priorlow'], w['priorup'],
                      nchains=4096, sampler='demc', fepsilon=0.01, thinning=1, nzchain=20, seed=3)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pop = make()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    pop.init_population('normal')
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f'rep {rep}: Population() {1e3*(t1-t0):.2f} ms, init_population {1e3*(t2-t1):.2f} ms')
pr = cProfile.Profile(); pr.enable(); pop = make(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(22)
pr = cProfile.Profile(); pr.enable(); pop.init_population('normal'); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(14)
