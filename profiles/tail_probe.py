"""Kernel time per (chain, point) of k_model_chisq<grid> against the series length:
separates steady-state loop efficiency from prologue/tail effects."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mc3_b200 import _lib
dev = torch.device('cuda')
nch = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rs = np.random.RandomState(0)
P = np.column_stack([rs.uniform(0.5, 2, nch), rs.uniform(0.5, 3.0, nch), rs.uniform(-3, 3, nch),
                     rs.uniform(-1, 1, nch), rs.uniform(-0.1, 0.1, nch)])
dP = torch.from_numpy(P).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for n in (100000, 200000, 400000, 1000000, 4000000):
    x = torch.linspace(0, 10, n, dtype=torch.float64, device=dev)
    d = torch.randn(n, dtype=torch.float64, device=dev); w = torch.ones(n, dtype=torch.float64, device=dev)
    ns = ctypes.c_int(0)
    _lib.call('mc3b_model_chisq_plan', nch, n, _lib.F64, ctypes.byref(ns))
    part = torch.empty((ns.value, nch), dtype=torch.float64, device=dev)
    def run():
        _lib.call('mc3b_model_chisq', 4, _lib.F64, dP.data_ptr(), 5, nch, 5, x.data_ptr(), d.data_ptr(),
                  w.data_ptr(), n, part.data_ptr(), nch, ns.value, _lib.stream_ptr())
    for _ in range(3): run()
    ts = []
    for _ in range(10):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    print(f'n={n:8d} nsplit={ns.value:4d} ms={ms:.4f} ps/chain-point={1e9*ms/(nch*n):.4f} '
          f'fp64-pipe-frac(6 instr)={(12.0*nch*n/(ms*1e-3))/36.3e12:.3f}')
