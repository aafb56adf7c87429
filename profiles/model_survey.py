"""Throughput of the fused model+chi-squared kernel over models, shapes and dtypes."""
import ctypes, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mc3_b200 import _lib

dev = torch.device('cuda')
P0 = {0: [1.0, 0.5, -0.3], 1: [1.0, 2.5, 0.3, 5.0, -0.2], 4: [1.0, 2.5, 0.3, 5.0, -0.2],
      2: [2.0, 5.0, 1.2, 0.5], 3: [0.01, 5.0, 1.0, 1.0]}
NAMES = {0: 'polynomial(3)', 1: 'sinusoid', 4: 'sinusoid(grid)', 2: 'gaussian', 3: 'box'}
rs = np.random.RandomState(0)
rows = []
for (nch, n) in ((4096, 100000), (65536, 100000), (512, 1000000), (64, 1000000), (7, 1000000), (7, 10000), (7, 1000)):
    x = np.linspace(0, 10, n)
    data = rs.normal(0, 1, n); w = np.ones(n)
    for dt, code, tdt in (('f64', _lib.F64, torch.float64), ('f32', _lib.F32, torch.float32)):
        dx, dd, dw = (torch.from_numpy(a).to(dev).to(tdt) for a in (x, data, w))
        for mid in (0, 1, 4, 2, 3):
            if mid == 4 and dt == 'f32':
                continue
            p = np.array(P0[mid]); P = p + rs.normal(0, 1e-3, (nch, p.size))
            dP = torch.from_numpy(P).to(dev)
            ns = ctypes.c_int(0)
            _lib.call('mc3b_model_chisq_plan', nch, n, code, ctypes.byref(ns))
            part = torch.empty((ns.value, nch), dtype=torch.float64, device=dev)
            def run():
                _lib.call('mc3b_model_chisq', mid, code, dP.data_ptr(), p.size, nch, p.size, dx.data_ptr(),
                          dd.data_ptr(), dw.data_ptr(), n, part.data_ptr(), nch, ns.value, _lib.stream_ptr())
            for _ in range(3): run()
            torch.cuda.synchronize()
            ts = []
            for _ in range(7):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); run(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
            ms = float(np.median(ts))
            rows.append((nch, n, dt, NAMES[mid], ms, nch*n/(ms*1e-3)))
            print(f'{nch:6d} x {n:8d} {dt} {NAMES[mid]:15s} {ms:9.4f} ms  {nch*n/(ms*1e-3):.3e} chain-points/s', flush=True)
json.dump(rows, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'r1_model_survey.json'), 'w'))
