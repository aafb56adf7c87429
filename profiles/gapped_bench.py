"""Config 2 with gaps in the time series (a constant cadence, three gaps, the cadence resuming off
the original grid): generation time with the piecewise-uniform layout (tile origins) and with the
plain sinusoid kernel (MC3B_NO_SEG=1).  CUDA-event timing of 200 replays of the generation graph."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mc3_b200 as mc3
from mc3_b200 import workloads
from mc3_b200.engine import Population

w = workloads.config2()
x, d = w['x'], w['data']
keep = np.ones(x.size, bool)
for lo, wd in ((20000, 3000), (50000, 10), (75000, 1), (99300, 300)):
    keep[lo:lo + wd] = False
x, d = x[keep].copy(), d[keep].copy()
x[x > 6.0] += 0.37*(w['x'][1] - w['x'][0])
out = {'n': int(x.size)}
unc_pp = 0.5*np.random.RandomState(1).uniform(0.8, 1.25, x.size)
for name, env, unc in (('segmented', None, np.full(x.size, 0.5)), ('segmented_per_point_sigma', None, unc_pp),
                       ('plain', 'MC3B_NO_SEG', np.full(x.size, 0.5)), ('plain_per_point_sigma', 'MC3B_NO_SEG', unc_pp)):
    os.environ.pop('MC3B_NO_SEG', None)
    if env:
        os.environ[env] = '1'
    pop = Population(d, unc, mc3.models.sinusoid, w['params'], [x], {},
                     w['pstep'], w['pmin'], w['pmax'], w['prior'], w['priorlow'], w['priorup'],
                     nchains=4096, sampler='demc', fepsilon=0.01, thinning=1, nzchain=400, seed=3)
    pop.init_population('normal')
    pop.run(10, use_graph=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    pop.run(200, use_graph=True)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)/200
    out[name] = {'ms_per_generation': ms, 'chain_steps_per_s': 4096/(ms*1e-3),
                 'tiles': None if pop.seg is None else int(pop.seg['starts'].size),
                 'leftover_points': None if pop.seg is None else pop.seg['nleft'],
                 'moment_form': bool(pop.use_moment),
                 'guard_hits': int(pop.guard_hits.item()) if pop.moment is not None else None}
print(json.dumps(out))
