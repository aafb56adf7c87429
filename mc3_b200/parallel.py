"""Multi-GPU plumbing: chains are partitioned across the GPUs of one box
(contiguous blocks, one process per GPU) and exchanged with NCCL through
torch.distributed.  Replaces the reference's shared ctypes
arrays + pipes (mc3/mcmc_driver.py:143-221, 302-307).

What travels, per generation:
  mrw      nothing (chains are independent)
  demc     all-gather of the current states X [nchains, nfree]   (chain.py:231)
  snooker  all-gather of the history rows written this generation (chain.py:197-217)
and at report points the thinned history / counters.  Every function here works
on plain tensors, so the same code runs under gloo on CPU in the tests.
"""
import ctypes

import torch
import torch.distributed as dist


def chain_slice(nchains, rank, world):
    """(first chain, number of chains) owned by `rank`: contiguous blocks."""
    if nchains % world != 0:
        raise ValueError(f'nchains ({nchains}) must be a multiple of the '
                         f'number of devices ({world})')
    nlocal = nchains // world
    return rank*nlocal, nlocal


def allgather_rows(block, chain0, nlocal, group=None):
    """In-place all-gather of `block` [nchains, w]: every rank contributes rows
    [chain0, chain0+nlocal) and receives all the others."""
    mine = block[chain0:chain0 + nlocal]
    if dist.get_backend(group) == 'nccl':
        dist.all_gather_into_tensor(block.view(-1), mine.reshape(-1), group=group)
    else:                                   # gloo: no in-place flat gather
        parts = [torch.empty_like(mine) for _ in range(dist.get_world_size(group))]
        dist.all_gather(parts, mine.contiguous(), group=group)
        block.copy_(torch.cat(parts).view_as(block))


def gather_history(t, M0, K, nchains, rank, world, group=None, dst=None):
    """Complete the thinned history tensor t ([zlen] or [zlen, w]) on every rank
    (dst=None) or on rank `dst` only (the other ranks just send their rows).
    Row M0 + k*nchains + c belongs to the owner of chain c; K thinned steps."""
    if world == 1 or K == 0:
        return
    nlocal = nchains // world
    w = 1 if t.dim() == 1 else t.shape[1]
    v = t[M0:M0 + K*nchains].view(K, world, nlocal*w)
    mine = v[:, rank].contiguous()
    if dst is None:
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)
        v.copy_(torch.stack(parts, dim=1))
    else:
        parts = [torch.empty_like(mine) for _ in range(world)] if rank == dst else None
        dist.gather(mine, parts, dst=dist.get_global_rank(group, dst) if group is not None else dst,
                    group=group)
        if rank == dst:
            v.copy_(torch.stack(parts, dim=1))


def sum_owned(t, chain0, nlocal, group=None):
    """All-reduce a per-chain tensor of which each rank only owns a slice."""
    m = torch.zeros_like(t)
    m[chain0:chain0 + nlocal] = t[chain0:chain0 + nlocal]
    dist.all_reduce(m, group=group)
    return m


# ---------------------------------------------------------------------------
# Peer memory: one buffer per device, mapped into every process of the box
# ---------------------------------------------------------------------------
class _RawCuda:
    """CUDA array interface over memory the C library allocated."""
    def __init__(self, ptr, nelem, typestr):
        self.__cuda_array_interface__ = {'shape': (nelem,), 'typestr': typestr,
                                         'data': (ptr, False), 'version': 2}


class PeerBuffer:
    """`nbytes` of zeroed device memory on every rank of `group`, each mapped into
    every process with CUDA IPC (mc3b_peer_alloc / mc3b_peer_open).  `local` is this
    rank's buffer as a float64 tensor; `ptrs` a device int64 tensor of all ranks'
    addresses in THIS process (what the kernels take as X_peers / Z_peers /
    F_peers).  Buffers are cached per (group, size, tag): the handle exchange is paid
    once per process, not once per mcmc() call."""
    _cache = {}

    @classmethod
    def get(cls, nbytes, tag, owner, rank, world, group, dev):
        """A cached buffer no live object owns, else a new one (every rank runs the
        same program, so every rank takes the same branch)."""
        import weakref
        key = (id(group) if group is not None else 0, int(nbytes), tag, dev.index)
        for pb in cls._cache.setdefault(key, []):
            if pb.owner is None or pb.owner() is None:
                pb.owner = weakref.ref(owner)
                return pb
        pb = cls(nbytes, rank, world, group, dev)
        pb.owner = weakref.ref(owner)
        cls._cache[key].append(pb)
        return pb

    def __init__(self, nbytes, rank, world, group, dev):
        from . import _lib
        nbytes = (int(nbytes) + 255)//256*256
        ptr = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        _lib.call('mc3b_peer_alloc', nbytes, ctypes.byref(ptr), handle)
        handles = [None]*world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        addrs = []
        for r in range(world):
            if r == rank:
                addrs.append(ptr.value)
            else:
                q = ctypes.c_void_p()
                _lib.call('mc3b_peer_open', handles[r], ctypes.byref(q))
                addrs.append(q.value)
        self.nbytes, self.addrs = nbytes, addrs
        self.local = torch.as_tensor(_RawCuda(ptr.value, nbytes//8, '<f8'), device=dev)
        self.ptrs = torch.tensor(addrs, dtype=torch.int64, device=dev)
