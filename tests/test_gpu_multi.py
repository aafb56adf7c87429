"""Two-GPU tests (skipped on a single-GPU box): the chain partition (peer-memory
exchange with generation flags by default, NCCL all-gather with MC3B_P2P=0) and
data sharding with the chi-squared all-gather must reproduce the single-GPU run --
at sizes that take the TMA path and the decreasing split schedule (n >= 20 000),
with the launch shape planned for the whole population (plan_chains) so that the
bits do not depend on the number of devices.  Semantics: chain.py:221-232 (demc
reads the population), :276-289 (history rows), gelman.py:36-92."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

NCH = 1024


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(rank, world, port, sampler, shard, n, opts, q):
    import torch.distributed as dist
    import mc3_b200 as mc3
    from mc3_b200 import workloads
    from mc3_b200.mcmc_driver import mcmc
    for k, v in opts.get('env', {}).items():
        os.environ[k] = v
    torch.cuda.set_device(rank)
    if world > 1:
        os.environ['MASTER_ADDR'] = '127.0.0.1'
        os.environ['MASTER_PORT'] = str(port)
        dist.init_process_group('nccl', rank=rank, world_size=world,
                                device_id=torch.device('cuda', rank))
    w = workloads.config2(n=n)
    if opts.get('ragged_sigma'):
        w['uncert'] = w['uncert']*(1 + 0.3*np.sin(np.arange(n)))
    ngen = opts.get('ngen', 24)
    out = mcmc(w['data'], w['uncert'], mc3.models.sinusoid, w['params'], [w['x']], {},
               w['pmin'], w['pmax'], w['pstep'], w['prior'], w['priorlow'], w['priorup'],
               NCH, None, NCH*ngen, sampler, False, None, opts.get('grtest', True), 0.0, 0.5, 6, 2,
               1.0, 0.01, 4, 'normal', opts.get('savefile'), False, mc3.Log(verb=-1), None, None,
               seed=21, rank=rank, world=world, shard=shard, plan_chains=NCH,
               use_graph=opts.get('use_graph'), return_population=True)
    pop = out.pop('_population')
    psrf = pop.gelman_rubin(2)
    if rank == 0:
        res = {k: out[k] for k in ('posterior', 'zchain', 'log_post', 'bestp', 'acceptance_rate',
                                   'medianp', 'stdp')}
        res['psrf'] = psrf
        res['p2p'] = pop.p2p is not None
        q.put(res)
        q.close()
        q.join_thread()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        pop.close()                      # graphs first, then the process group
        del pop, out
        dist.destroy_process_group()


def _launch(world, sampler, shard='chains', n=20011, **opts):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_run, args=(r, world, port, sampler, shard, n, opts, q))
             for r in range(world)]
    for pr in procs:
        pr.start()
    res = q.get(timeout=240)
    for pr in procs:
        pr.join(120)
        assert pr.exitcode == 0, 'a rank did not shut down cleanly'
    return res


def _same(one, two):
    for k in ('posterior', 'zchain', 'log_post', 'bestp', 'medianp', 'stdp', 'psrf'):
        assert np.array_equal(one[k], two[k]), k
    assert one['acceptance_rate'] == two['acceptance_rate']


needs2 = pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')


@needs2
@pytest.mark.parametrize('sampler', ['demc', 'snooker', 'mrw'])
def test_chain_partition_equals_single_gpu(sampler):
    """Peer-memory exchange (default): byte-identical history, statistics and
    Gelman-Rubin values on 1 and 2 GPUs; graphs captured (snooker included)."""
    one = _launch(1, sampler)
    two = _launch(2, sampler)
    assert two['p2p'] == (sampler != 'mrw')
    _same(one, two)


@needs2
@pytest.mark.parametrize('sampler', ['demc', 'snooker'])
def test_chain_partition_nccl_fallback(sampler):
    one = _launch(1, sampler)
    two = _launch(2, sampler, env={'MC3B_P2P': '0'})
    assert not two['p2p']
    _same(one, two)


@needs2
def test_chain_partition_per_point_uncertainties_and_eager():
    one = _launch(1, 'demc', ragged_sigma=True, use_graph=False)
    two = _launch(2, 'demc', ragged_sigma=True, use_graph=False)
    _same(one, two)


@needs2
def test_savefile_on_two_gpus(tmp_path):
    """Every rank takes part in the intermediate saves (collective history gather),
    rank 0 writes: no mismatched collectives, same results as without a savefile."""
    ref = _launch(2, 'demc')
    sv = _launch(2, 'demc', savefile=str(tmp_path/'mc.npz'))
    _same(ref, sv)
    with np.load(str(tmp_path/'mc.npz')) as f:
        assert np.array_equal(f['posterior'], sv['posterior'])


@needs2
def test_data_sharding_matches_single_gpu():
    """N = 1e6 points split over two devices (config 5's decomposition): the
    chi-squared is summed in a different order (two halves), so trajectories agree
    except where a decision sat within rounding of its threshold."""
    one = _launch(1, 'demc', n=1_000_000, ngen=12)
    two = _launch(2, 'demc', shard='data', n=1_000_000, ngen=12)
    assert np.array_equal(one['zchain'], two['zchain'])
    same = np.all(one['posterior'] == two['posterior'], axis=1).mean()
    assert same > 0.98
    np.testing.assert_allclose(one['log_post'][:NCH*3], two['log_post'][:NCH*3], rtol=1e-11)
    np.testing.assert_allclose(one['psrf'], two['psrf'], rtol=1e-3)
