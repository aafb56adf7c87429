"""CPU-only, world_size 2 over gloo: the chain partition and the exchanges the
multi-GPU path uses (mc3_b200/parallel.py) -- population all-gather, history
gather in reference row order, owned-slice reductions."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mc3_b200 import parallel as par


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        nchains, nfree, M0, K = 8, 3, 16, 5
        c0, nl = par.chain_slice(nchains, rank, world)
        full = torch.arange(nchains*nfree, dtype=torch.float64).view(nchains, nfree)
        # population all-gather: every rank only knows its own rows
        X = torch.full((nchains, nfree), -1.0, dtype=torch.float64)
        X[c0:c0 + nl] = full[c0:c0 + nl]
        par.allgather_rows(X, c0, nl)
        ok_x = torch.equal(X, full)
        # history gather, row M0 + k*nchains + c
        zlen = M0 + K*nchains
        truth = torch.arange(zlen*nfree, dtype=torch.float64).view(zlen, nfree)
        Z = torch.zeros_like(truth)
        Z[:M0] = truth[:M0]
        zc = torch.full((zlen,), -1, dtype=torch.int32)
        for k in range(K):
            r = M0 + k*nchains + c0
            Z[r:r + nl] = truth[r:r + nl]
            zc[r:r + nl] = torch.arange(c0, c0 + nl, dtype=torch.int32)
        par.gather_history(Z, M0, K, nchains, rank, world)
        par.gather_history(zc, M0, K, nchains, rank, world)
        ok_z = torch.equal(Z, truth)
        want = torch.cat([torch.full((M0,), -1, dtype=torch.int32),
                          torch.arange(nchains, dtype=torch.int32).repeat(K)])
        ok_c = torch.equal(zc, want)
        # history gather to rank 0 only, in two instalments (rows [0, 3) then [3, 5))
        Z2 = torch.zeros_like(truth)
        Z2[:M0] = truth[:M0]
        for k in range(K):
            r = M0 + k*nchains + c0
            Z2[r:r + nl] = truth[r:r + nl]
        mine_before = Z2.clone()
        par.gather_history(Z2, M0, 3, nchains, rank, world, dst=0)
        par.gather_history(Z2, M0 + 3*nchains, K - 3, nchains, rank, world, dst=0)
        ok_z = ok_z and (torch.equal(Z2, truth) if rank == 0 else torch.equal(Z2, mine_before))
        # packed report counters: one row per rank, combined on the host
        # (what Population.counters_async all-gathers; layout of mc3b_pack_counters)
        nf = 3
        row = torch.tensor([10.0 + rank, 5.0 - rank, 7.0, float(c0 + 1)] +
                           [float(rank)]*nf + [1.0, 2.0, 3.0 + rank], dtype=torch.float64)
        parts = [torch.empty_like(row) for _ in range(world)]
        dist.all_gather(parts, row)
        h = torch.stack(parts).numpy()
        ok_z = ok_z and h[:, 0].sum() == 21.0 and h[:, 4 + nf:].sum(axis=0).tolist() == [2.0, 4.0, 7.0] \
            and min(h, key=lambda r: (r[1], r[2], r[3]))[3] == 5.0
        # owned-slice reduction
        acc = torch.zeros(nchains, dtype=torch.int32)
        acc[c0:c0 + nl] = rank + 1
        acc[(c0 + nl) % nchains] = 99          # garbage outside the owned slice
        tot = par.sum_owned(acc, c0, nl)
        ok_s = torch.equal(tot, torch.tensor([1]*4 + [2]*4, dtype=torch.int32))
        q.put((rank, ok_x, ok_z, ok_c, ok_s))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_exchanges():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    for r in res:
        assert all(r[1:]), r


def test_chain_slice_rules():
    assert par.chain_slice(4096, 3, 8) == (1536, 512)
    with pytest.raises(ValueError, match='multiple of the number of devices'):
        par.chain_slice(10, 0, 4)
