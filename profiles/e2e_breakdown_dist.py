"""cProfile of mcmc() on rank 0 under torchrun (config 2, 4096 chains per GPU, K generations)."""
import os, sys, time, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import mc3_b200 as mc3
from mc3_b200 import workloads, mcmc_driver as md

K = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('WORLD_SIZE', 1), ('LOCAL_RANK', 0)))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
w = workloads.config2()
n = 4096*world
def hub(seed):
    with contextlib.redirect_stdout(sys.stderr):
        return md.mcmc(w['data'], w['uncert'], mc3.models.sinusoid, w['params'], [w['x']], {},
                       w['pmin'], w['pmax'], w['pstep'], w['prior'], w['priorlow'], w['priorup'],
                       n, None, n*K, 'demc', False, None, False, 0.0, 0.5, 0, 1, 1.0, 0.01,
                       10, 'normal', None, False, mc3.Log(verb=-1), None, None, seed=seed,
                       rank=rank, world=world)
def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
hub(1); out = None
for rep in range(2):
    barrier(); t0 = time.perf_counter(); out = hub(2 + rep); barrier()
    if rank == 0: print('mcmc() wall', 1e3*(time.perf_counter() - t0), 'ms', file=sys.stderr)
    out = None
import cProfile, pstats
barrier()
pr = cProfile.Profile(); pr.enable(); out = hub(9); barrier(); pr.disable()
if rank == 0:
    pstats.Stats(pr, stream=sys.stderr).sort_stats('cumulative').print_stats(30)
    pstats.Stats(pr, stream=sys.stderr).sort_stats('tottime').print_stats(14)
sys.stderr.flush()
os._exit(0)
