"""Where one generation of config 2 goes (CUDA-event timing of back-to-back launches):
proposal kernel alone, model kernel alone, model kernel with the fused Metropolis
epilogue, and the captured generation graph."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mc3_b200 as mc3
from mc3_b200 import _lib, workloads
from mc3_b200.engine import Population

w = workloads.config2()
pop = Population(w['data'], w['uncert'], mc3.models.sinusoid, w['params'], [w['x']], {},
                 w['pstep'], w['pmin'], w['pmax'], w['prior'], w['priorlow'], w['priorup'],
                 nchains=4096, sampler='demc', fepsilon=0.01, thinning=1, nzchain=4000, seed=3)
pop.init_population('normal')
pop.run(5, use_graph=True)
torch.cuda.synchronize()
st = _lib.stream_ptr()


def timeit(fn, reps=300):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return 1e3*a.elapsed_time(b)/reps      # us per call


P = pop.nextp[:4096]
g = pop.gen
print('k_propose (host-driven gen)      %8.2f us' % timeit(lambda: _lib.call('mc3b_propose', ctypes.byref(pop.S), g, pop.zsize(), 0, 4096, st)))
print('moment form in use:', getattr(pop, 'use_moment', False))
print('pair kernel alone (k_fold_consts + k_sinefold) %8.2f us' % timeit(lambda: pop.data_chisq(P)))
zr = pop.M0 + 10*4096


def fused():
    pop.data_chisq(P, fuse=(0, g, -1, False))
print('generation kernel + fused Metropolis  %8.2f us' % timeit(fused))
if getattr(pop, 'use_moment', False):
    pop.use_moment = False
    print('pair kernel + fused Metropolis       %8.2f us' % timeit(fused))
    pop.use_moment = True
part, ld, ns = pop.data_chisq(P)
print('k_metropolis alone               %8.2f us' % timeit(lambda: _lib.call('mc3b_metropolis', ctypes.byref(pop.S), part.data_ptr(), ld, ns, 0, g, -1, 0, 4096, st)))
print('k_advance alone                  %8.2f us' % timeit(lambda: _lib.call('mc3b_advance', ctypes.byref(pop.S), st)))
pop.gen_dev.fill_(pop.gen)
print('generation graph replay          %8.2f us' % timeit(lambda: pop._graph.replay(), reps=300))
print('nsplit', ns)
