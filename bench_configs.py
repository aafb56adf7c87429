#!/usr/bin/env python
"""Measurements of the other BASELINE.json configurations (bench.py stays on
config 2, the configuration the headline metric is quoted on).

    python bench_configs.py config3 [--chains 16384] [--steps 20]
    python bench_configs.py config4 [--n 100000000]
    python bench_configs.py config1

Each prints one JSON line: device-timed throughput (CUDA events), the roofline
of the dominant kernel, and a bounded CPU sample of the oracle / reference C code.
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        return {'hbm_gbs': 6650.0, 'note': 'fallback'}


def fp64_peak(torch, _lib):
    sink = torch.zeros(8, dtype=torch.float64, device='cuda')
    fl = ctypes.c_double()
    best = 0.0
    for it in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.call('mc3b_fma_peak', _lib.F64, 20000, sink.data_ptr(), ctypes.byref(fl),
                  _lib.stream_ptr())
        b.record()
        torch.cuda.synchronize()
        if it:
            best = max(best, fl.value/(a.elapsed_time(b)*1e-3))
    return best


def timed(torch, fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def config3(args):
    """snooker + wavelet likelihood, 16384 chains on N = 2^20 points (launch under
    torchrun for several GPUs: chains partitioned, history rows stored into every
    device over NVLink, generation flags, captured graphs)."""
    import torch
    import torch.distributed as dist
    import mc3_b200 as mc3
    from mc3_b200 import _lib, workloads
    from mc3_b200.engine import Population
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    w = workloads.config3()
    n = w['x'].size
    nch = args.chains
    K = args.steps
    pop = Population(w['data'], w['uncert'], mc3.models.box, w['params'], [w['x']], {},
                     w['pstep'], w['pmin'], w['pmax'], w['prior'], w['priorlow'],
                     w['priorup'], nchains=nch, sampler='snooker', wlike=True,
                     thinning=1, nzchain=K + 4, seed=5, hsize=args.hsize, rank=rank, world=world)
    t0 = time.perf_counter()
    pop.init_population('normal')
    torch.cuda.synchronize()
    t_init = time.perf_counter() - t0
    pop.run(3)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    pop.run(K)
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    P = pop.nextp[pop.chain0:pop.chain0 + pop.nlocal]
    kms, _ = timed(torch, lambda: pop.data_chisq(P), reps=5, warm=1)
    peak = fp64_peak(torch, _lib)
    flops = float(w['flops_per_point'])*pop.nlocal*n
    c = pop.counters()
    if rank == 0:
        # CPU: dwt_chisq of one chain with the numpy box model -- the reference's own
        # C extension (oracle/_ref) when it is built, else the oracle port
        from oracle import kernels as ok, models as om, ref
        p = w['params']
        kind = 'port'
        fn = lambda: ok.dwt_chisq(om.box(p[:4], w['x']), w['data'], p)
        if ref.have_ref_ext():
            rdwt = ref.ref_ext()[1]
            fn = lambda: rdwt.chisq(p, om.box(p[:4], w['x']), w['data'])
            kind = 'reference'
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            fn()
        cpu_eval = (time.perf_counter() - t0)/reps
        line = {
            'metric': 'chain-steps/s', 'value': nch*K/(ms*1e-3), 'unit': 'chain-steps/s',
            'n_gpus': world, 'steps': K, 'ms_per_step': ms/K, 'dtype': 'f64', 'data': 'synthetic',
            'scaling': 'strong' if world > 1 else None,
            'config': {'workload': w['name'], 'nchains': nch, 'ndata': n, 'sampler': 'snooker',
                       'wlike': True, 'hsize': args.hsize,
                       'exchange': 'peer-memory stores of the new history rows + generation flags'
                                   if pop.p2p is not None else ('NCCL all-gather' if world > 1 else 'single GPU'),
                       'graph': pop._graph is not None},
            'chisq_evals_per_s': nch*K/(ms*1e-3)*n,
            'init_population_s': t_init,
            'acceptance_rate_pct': 100.0*c['numaccept']/(nch*(K + 3)),
            'roofline': {'bound': 'fp64',
                         'kernel': 'k_dwt_reg_model (4 levels in registers) + k_dwt_reg + k_dwt_pass + k_dwt_last',
                         'achieved': flops/(kms*1e-3)/1e12, 'peak': peak/1e12, 'unit': 'TFLOP/s',
                         'frac': flops/(kms*1e-3)/peak, 'ms_per_launch_set': kms,
                         'chains_per_launch': pop.nlocal,
                         'algorithmic_flops_per_chain_point': w['flops_per_point']},
            'cpu_baseline': {'value': 1.0/cpu_eval, 'unit': 'chain-steps/s', 'cores': 1,
                             'kind': kind, 'sample': f'{reps} evaluations of dwt_chisq ('
                             f'{"reference _dwt.chisq from oracle/_ref" if kind == "reference" else "oracle C port"}'
                             f') + numpy box model, one chain, N=2^20: {1e3*cpu_eval:.1f} ms each'},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        pop.close()
        del pop
        torch.cuda.synchronize()
        dist.destroy_process_group()


def config4(args):
    import torch
    import mc3_b200 as mc3
    from mc3_b200 import _lib
    from oracle import problems as pb, kernels as ok
    n = args.n
    rs = np.random.RandomState(20260104)
    t0 = time.perf_counter()
    white = rs.normal(0.0, 1.0, n)
    from scipy.signal import lfilter
    red = lfilter([1.0], [1.0, -0.95], rs.normal(0.0, 0.2, n))
    series = white + red
    unc = np.abs(rs.normal(0.0, 1.0, n)) + 0.5
    gen_s = time.perf_counter() - t0
    dev = torch.device('cuda')
    d = torch.from_numpy(series).to(dev)
    u = torch.from_numpy(unc).to(dev)
    st = _lib.stream_ptr
    hbm = peaks().get('hbm_gbs', 6650.0)
    out = {'metric': 'time_avg + bin_array', 'n': n, 'unit': 'ms', 'data': 'synthetic',
           'config': {'workload': 'config4: rms-vs-binsize (maxbins 1000) + bin_array(100) on a '
                                  f'{n:.0e}-point white+AR(1) series', 'host_generation_s': gen_s}}
    # bin_array
    nb = n//100
    bd = torch.empty(nb, dtype=torch.float64, device=dev)
    bs = torch.empty(nb, dtype=torch.float64, device=dev)
    for name, fn, nbytes in (
        ('bin_array_unweighted', lambda: _lib.call('mc3b_binarray', d.data_ptr(), n, 100, None,
                                                   bd.data_ptr(), None, st()), 8.0*n + 8.0*nb),
        ('bin_array_weighted', lambda: _lib.call('mc3b_binarray', d.data_ptr(), n, 100, u.data_ptr(),
                                                 bd.data_ptr(), bs.data_ptr(), st()), 16.0*n + 16.0*nb)):
        med, mn = timed(torch, fn)
        out[name] = {'ms': med, 'ms_min': mn,
                     'roofline': {'bound': 'hbm', 'achieved': nbytes/(med*1e-3)/1e9, 'peak': hbm,
                                  'unit': 'GB/s', 'frac': nbytes/(med*1e-3)/1e9/hbm,
                                  'algorithmic_bytes': nbytes}}
    # time_avg
    maxbins, binstep = 1000, 1
    nout = (maxbins - 1)//binstep + 1
    lib = _lib.load()
    ws = torch.empty(lib.mc3b_binrms_workspace(n, maxbins, binstep)//8, dtype=torch.float64, device=dev)
    outs = [torch.empty(nout, dtype=torch.float64, device=dev) for _ in range(5)]
    med, mn = timed(torch, lambda: _lib.call('mc3b_binrms', d.data_ptr(), n, maxbins, binstep,
                                             ws.data_ptr(), *[o.data_ptr() for o in outs], st()))
    # algorithmic traffic of this formulation: read x twice (prefix, deviations), write + gather P
    nbins_total = float(sum(n//b for b in range(1, maxbins + 1, binstep)))
    peak64 = fp64_peak(torch, _lib)                    # flops/s; one add per FMA slot = peak64/2 adds/s
    direct_ms = 1e3*float(n)*nout/(peak64/2.0)
    hbm_ms = 1e3*8.0*n/(hbm*1e9)
    out['time_avg'] = {'ms': med, 'ms_min': mn, 'bin_sizes': nout, 'bins_evaluated': nbins_total,
                       # report against max(8N / BW_HBM, N * nsizes / FP64 add peak), the
                       # bound of the reference's direct algorithm (one add per point and bin size) with
                       # all sizes produced in one pass; the prefix formulation does N ln(maxbins) bin
                       # evaluations instead of N * nsizes adds, so it can (and does) beat that bound
                       'roofline': {'bound': 'max(hbm one pass, fp64 adds of the direct algorithm)',
                                    'bound_ms': max(direct_ms, hbm_ms), 'direct_algorithm_adds_ms': direct_ms,
                                    'hbm_one_pass_ms': hbm_ms, 'achieved_ms': med,
                                    'frac': max(direct_ms, hbm_ms)/med,
                                    'note': 'frac > 1: faster than the direct algorithm at the FP64 add peak',
                                    'hbm': {'achieved': (16.0*n)/(med*1e-3)/1e9, 'peak': hbm, 'unit': 'GB/s',
                                            'frac': (16.0*n)/(med*1e-3)/1e9/hbm,
                                            'algorithmic_bytes': 16.0*n,
                                            'passes': 'two reads of the series (moments; tile kernel)'}}}
    # parity at full size against direct sums (a few bin sizes) and oracle on a prefix
    rms = outs[0].cpu().numpy()
    err = outs[3].cpu().numpy()
    chk = {}
    for i in (0, 9, 99, 499, 999):
        b = 1 + i*binstep
        M = n//b
        means = series[:M*b].reshape(M, b).mean(axis=1)
        chk[b] = abs(rms[i]/np.sqrt(np.mean(means**2)) - 1.0)
    out['time_avg']['max_rel_err_vs_direct'] = max(chk.values())
    # CPU baselines (bounded): reference C binarray on the full series, binrms on 2e6 points
    t0 = time.perf_counter(); ok.bin_array(series, 100); t_ba = time.perf_counter() - t0
    t0 = time.perf_counter(); ok.bin_array(series, 100, unc); t_baw = time.perf_counter() - t0
    ns = min(n, 2_000_000)
    t0 = time.perf_counter(); ok.time_avg(series[:ns], 1000, 1); t_ta = time.perf_counter() - t0
    out['cpu_baseline'] = {'kind': 'port', 'cores': 1,
                           'bin_array_unweighted_ms': 1e3*t_ba, 'bin_array_weighted_ms': 1e3*t_baw,
                           'time_avg_ms_extrapolated': 1e3*t_ta*n/ns,
                           'sample': f'oracle C binarray on all {n} points; oracle C binrms on the first '
                                     f'{ns} points ({t_ta:.2f} s), scaled linearly in N'}
    print(json.dumps(out), flush=True)


def config1(args):
    """examples/get_started.py as written: snooker, 7 chains, 1e5 samples."""
    import torch
    import mc3_b200 as mc3
    from oracle import problems as pb
    p = pb.mcmc_case('quad')
    quiet = mc3.Log(verb=-1)
    res = {}
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = mc3.sample(p['data'], p['uncert'], func=mc3.models.polynomial, params=p['params'],
                         indparams=[p['x']], pstep=p['pstep'], sampler='snooker', nsamples=1e5,
                         burnin=1000, nchains=7, seed=3, log=quiet)
        torch.cuda.synchronize()
        res[rep] = time.perf_counter() - t0
    line = {'metric': 'chain-steps/s', 'value': 100002/res[1], 'unit': 'chain-steps/s',
            'config': {'workload': 'config1: get_started quadratic, snooker, 7 chains, N=100, 1e5 samples'},
            'seconds': res[1], 'seconds_first_call': res[0],
            'bestp': out['bestp'].tolist(), 'best_chisq': float(out['best_chisq']),
            'medianp': out['medianp'].tolist(), 'acceptance_rate': float(out['acceptance_rate']),
            'reference_docs': {'bestp': [3.0768, -2.5000, 0.5089], 'best_chisq': 112.5923,
                               'acceptance_rate': 28.36, 'seconds': 6.0}}
    print(json.dumps(line), flush=True)


def config5(args):
    """MRW + DEMC, 65536 chains over the GPUs of one box (launch under torchrun), with
    the Gelman-Rubin test on the device every 10 generations.
      --shard chains: chains partitioned, data replicated, population exchange
      --shard data:   data sharded, chi-squared all-gather, proposals replicated
      --scaling weak:   N = --n5 points PER GPU (BASELINE config 5: 1e6), so the work
                        per GPU (65536 x 1e6 chain-points per generation) is fixed
      --scaling strong: N = --n5 points in TOTAL, fixed problem
    One JSON line: per-sampler device-timed throughput (CUDA events, max over ranks),
    the model kernel timed alone against the FP64 FMA peak, and on rank 0 a bounded CPU
    sample of the same chain-step (numpy model + oracle C chi-squared)."""
    import torch
    import torch.distributed as dist
    import mc3_b200 as mc3
    from mc3_b200 import _lib, workloads
    from mc3_b200.engine import Population
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    n_total = args.n5*world if args.scaling == 'weak' else args.n5
    w = workloads.config2(n=n_total, seed=20260105)
    w['x'] = np.linspace(0, 10.0*n_total/1e6, n_total)
    w['data'] = workloads.sinusoid_np(np.array([1.0, 2.5, 0.3, 5.0, -0.2]), w['x']) + \
        np.random.RandomState(20260105).normal(0, 0.5, n_total)
    K = args.steps
    lines = []
    roof = None
    for sampler in ('mrw', 'demc'):
        pop = Population(w['data'], w['uncert'], mc3.models.sinusoid, w['params'], [w['x']], {},
                         w['pstep'], w['pmin'], w['pmax'], w['prior'], w['priorlow'], w['priorup'],
                         nchains=args.chains5, sampler=sampler, fepsilon=0.01, thinning=1,
                         nzchain=K + 4, seed=9, hsize=args.hsize, rank=rank, world=world,
                         shard=args.shard)
        t0 = time.perf_counter()
        pop.init_population('normal')
        torch.cuda.synchronize()
        t_init = time.perf_counter() - t0
        pop.run(3)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        done, psrf_h = 0, []
        while done < K:                              # Gelman-Rubin on the device every 10 generations
            step = min(10, K - done)
            pop.run(step)
            done += step
            psrf_h.append(pop.gelman_rubin_async(0))
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        psrf = psrf_h[-1][0].numpy() if psrf_h[-1] is not None else np.zeros(pop.nfree)
        c = pop.counters()
        if roof is None:
            P = pop.nextp[pop.chain0:pop.chain0 + pop.nlocal]
            moment = bool(getattr(pop, 'use_moment', False))
            # as the generation launches it: fused with the Metropolis step (chains partitioned) or
            # followed by the guarded finish (data sharded)
            fuse = (pop.chain0, pop.gen, -1, False) if moment and pop.fused else None
            kms, _ = timed(torch, lambda: pop._data_chisq_local(P, fuse, moment), reps=5, warm=1)
            peak = fp64_peak(torch, _lib)
            flops = float(w['flops_per_point'])*pop.nlocal*pop.ndata
            kname = 'k_model_chisq'
            if pop.grid:
                kname = ('k_fold_consts + k_sinefold<MOM> + Metropolis epilogue' if moment and pop.fused else
                         'k_fold_consts + k_sinefold<MOM> (guard in k_moment_finish, not timed)' if moment else
                         'k_fold_consts + k_sinefold' if getattr(pop, 'd_fold', None) is not None else
                         'k_sinegrid<USIG=%s>' % ('true' if pop.usig else 'false'))
            roof = {'bound': 'fp64', 'kernel': kname,
                    'achieved': flops/(kms*1e-3)/1e12, 'peak': peak/1e12, 'unit': 'TFLOP/s',
                    'frac': flops/(kms*1e-3)/peak, 'ms_per_launch': kms,
                    'chains_per_launch': pop.nlocal, 'points_per_launch': pop.ndata,
                    'algorithmic_flops_per_chain_point': w['flops_per_point'],
                    'hbm_stream_GBs': 8.0*pop.ndata/(kms*1e-3)/1e9}
        lines.append({'sampler': sampler, 'ms_per_step': ms/K,
                      'chain_steps_per_s': args.chains5*K/(ms*1e-3),
                      'chisq_evals_per_s': args.chains5*K/(ms*1e-3)*n_total,
                      'init_population_s': t_init,
                      'exchange': ('none (independent chains)' if sampler == 'mrw' and args.shard == 'chains'
                                   else 'chi-squared all-gather' if args.shard == 'data' and world > 1
                                   else 'peer-memory stores + generation flags' if pop.p2p is not None
                                   else 'NCCL all-gather' if world > 1 else 'single GPU'),
                      'acceptance_pct': 100.0*c['numaccept']/(args.chains5*(K + 3)),
                      'gelman_rubin_every': 10, 'gelman_rubin_max': float(np.max(psrf))})
        pop.close()
        del pop
        torch.cuda.empty_cache()
    cpu = None
    if rank == 0:
        from oracle import kernels as ok, models as om
        ns = min(n_total, 1_000_000)
        xs, ds, us = w['x'][:ns], w['data'][:ns], w['uncert'][:ns]
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < 5.0:
            ok.chisq(om.sinusoid(w['params'], xs), ds, us, w['params'], w['prior'], w['priorlow'], w['priorup'])
            reps += 1
        per = (time.perf_counter() - t0)/reps*(n_total/ns)
        cpu = {'value': 1.0/per, 'unit': 'chain-steps/s', 'cores': 1, 'kind': 'port',
               'sample': f'{reps} chain-steps of numpy sinusoid + oracle C chi-squared on {ns} points '
                         f'(scaled linearly to N={n_total}): {1e3*per:.1f} ms per chain-step'}
        print(json.dumps({'metric': 'chain-steps/s', 'unit': 'chain-steps/s',
                          'value': max(r['chain_steps_per_s'] for r in lines),
                          'n_gpus': world, 'shard': args.shard, 'scaling': args.scaling, 'dtype': 'f64',
                          'data': 'synthetic',
                          'config': {'workload': 'config5: MRW + DEMC, 65536 chains, sinusoid+line, '
                                     f'N={n_total:.0e} points in total ({n_total//world:.0e} per GPU), '
                                     'Gelman-Rubin on device every 10 generations',
                                     'nchains': args.chains5, 'ndata_total': n_total,
                                     'hsize': args.hsize, 'steps': K},
                          'runs': lines, 'roofline': roof, 'cpu_baseline': cpu}), flush=True)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('which', choices=['config1', 'config3', 'config4', 'config5'])
    ap.add_argument('--shard', default='chains', choices=['chains', 'data'])
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'])
    ap.add_argument('--chains5', type=int, default=65536)
    ap.add_argument('--n5', type=int, default=1_000_000)
    ap.add_argument('--chains', type=int, default=16384)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--hsize', type=int, default=10)
    ap.add_argument('--n', type=int, default=100_000_000)
    a = ap.parse_args()
    {'config1': config1, 'config3': config3, 'config4': config4, 'config5': config5}[a.which](a)
