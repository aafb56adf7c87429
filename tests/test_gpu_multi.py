"""Two-GPU tests (skipped on a single-GPU box): chain partition with the NCCL
population all-gather and data sharding with the chi-squared all-gather must
reproduce the single-GPU run."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import problems as pb


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem():
    p = pb.mcmc_case('sine')
    n = 200                      # < one tile: every chi-squared sum has one fixed order
    return {k: (v[:n] if isinstance(v, np.ndarray) and v.shape == (512,) else v)
            for k, v in p.items()}


def _run(rank, world, port, sampler, shard, q):
    import torch.distributed as dist
    import mc3_b200 as mc3
    from mc3_b200.mcmc_driver import mcmc
    torch.cuda.set_device(rank)
    if world > 1:
        os.environ['MASTER_ADDR'] = '127.0.0.1'
        os.environ['MASTER_PORT'] = str(port)
        dist.init_process_group('nccl', rank=rank, world_size=world,
                                device_id=torch.device('cuda', rank))
    p = _problem()
    out = mcmc(p['data'], p['uncert'], mc3.models.sinusoid, p['params'], [p['x']], {},
               p['pmin'], p['pmax'], p['pstep'], p['prior'], p['priorlow'], p['priorup'],
               256, None, 256*30, sampler, False, None, True, 0.0, 0.5, 6, 2, 1.0, 0.01,
               4, 'normal', None, False, mc3.Log(verb=-1), None, None, seed=21,
               rank=rank, world=world, shard=shard)
    if rank == 0:
        q.put({k: out[k] for k in ('posterior', 'zchain', 'log_post', 'bestp',
                                   'acceptance_rate', 'medianp')})
        q.close()
        q.join_thread()              # flush the payload before leaving without destructors
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        os._exit(0)


def _launch(world, sampler, shard):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_run, args=(r, world, port, sampler, shard, q))
             for r in range(world)]
    for pr in procs:
        pr.start()
    res = q.get(timeout=100)
    for pr in procs:
        pr.join(60)
    return res


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
@pytest.mark.parametrize('sampler', ['demc', 'snooker', 'mrw'])
def test_chain_partition_equals_single_gpu(sampler):
    one = _launch(1, sampler, 'chains')
    two = _launch(2, sampler, 'chains')
    for k in ('posterior', 'zchain', 'log_post', 'bestp'):
        assert np.array_equal(one[k], two[k]), k
    assert one['acceptance_rate'] == two['acceptance_rate']


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_data_sharding_matches_single_gpu():
    one = _launch(1, 'demc', 'chains')
    two = _launch(2, 'demc', 'data')
    assert np.array_equal(one['zchain'], two['zchain'])
    # the chi-squared is summed in a different order (two halves): same trajectory
    # unless a decision sat within rounding of its threshold
    same = np.all(one['posterior'] == two['posterior'], axis=1).mean()
    assert same > 0.98
    n = 256*3
    np.testing.assert_allclose(one['log_post'][:n], two['log_post'][:n], rtol=1e-12)
