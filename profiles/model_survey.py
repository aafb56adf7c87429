"""Roofline lines of the fused model + chi-squared kernel for every built-in model at
the config-2 shape (4096 chains x 1e5 points): fp64 and fp32, per-point and uniform
uncertainties, uniform and jittered abscissa.  One JSON line per case on stdout
(kept under profiles/r2_model_survey.jsonl).

    python profiles/model_survey.py > profiles/r2_model_survey.jsonl

Algorithmic flops per chain-point follow SURVEY 8(d): FMA = 2, a transcendental = 1;
chi-squared core 3 + model (quadratic 4, sinusoid+line 7, Gaussian line 7, box 3).
The denominator is the FMA peak of the arithmetic type measured live (mc3b_fma_peak).
"""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mc3_b200 import _lib

dev = torch.device('cuda')
NCH, N = 4096, 100000
P0 = {0: [1.0, 0.5, -0.3], 1: [1.0, 2.5, 0.3, 5.0, -0.2], 4: [1.0, 2.5, 0.3, 5.0, -0.2],
      2: [2.0, 5.0, 1.2, 0.5], 3: [0.01, 5.0, 1.0, 1.0]}
NAMES = {0: 'polynomial(3)', 1: 'sinusoid', 4: 'sinusoid on a uniform grid', 2: 'gaussian', 3: 'box'}
FLOPS = {0: 7, 1: 10, 4: 10, 2: 10, 3: 6}
rs = np.random.RandomState(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def peak(code):
    sink = torch.zeros(8, dtype=torch.float64, device=dev)
    fl = ctypes.c_double(0.0)
    best = 0.0
    for it in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.call('mc3b_fma_peak', code, 20000, sink.data_ptr(), ctypes.byref(fl), _lib.stream_ptr())
        b.record()
        torch.cuda.synchronize()
        if it:
            best = max(best, fl.value/(a.elapsed_time(b)*1e-3))
    return best


PEAK = {_lib.F64: peak(_lib.F64), _lib.F32: peak(_lib.F32)}
x_grid = np.linspace(0, 10, N)
x_jit = np.sort(x_grid + rs.uniform(-0.3, 0.3, N)*(x_grid[1] - x_grid[0]))
data = rs.normal(0, 1, N)
for dt, code, tdt in (('f64', _lib.F64, torch.float64), ('f32', _lib.F32, torch.float32)):
    for mid in (4, 1, 0, 2, 3):
        if mid == 4 and dt == 'f32':
            continue
        for usig in (True, False):
            x = x_grid if mid == 4 else x_jit
            w = np.full(N, 2.0) if usig else rs.uniform(0.5, 1.5, N)
            dx, dd, dw = (torch.from_numpy(a).to(dev).to(tdt) for a in (x, data, w))
            p = np.array(P0[mid])
            P = p + rs.normal(0, 1e-3, (NCH, p.size))
            dP = torch.from_numpy(P).to(dev)
            ns = ctypes.c_int(0)
            _lib.call('mc3b_model_chisq_plan', NCH, N, code, ctypes.byref(ns))
            part = torch.empty((ns.value, NCH), dtype=torch.float64, device=dev)
            o = _lib.ChisqOpts()
            o.uniform_sigma = 1 if usig else 0

            def run():
                _lib.call('mc3b_model_chisq_ex', mid, code, dP.data_ptr(), p.size, NCH, p.size,
                          dx.data_ptr(), dd.data_ptr(), dw.data_ptr(), N, part.data_ptr(), NCH,
                          ns.value, ctypes.byref(o), _lib.stream_ptr())
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            ts = []
            for k in range(20):
                flush.fill_(k)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                run()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            ms = float(np.mean(ts))
            ach = FLOPS[mid]*NCH*N/(ms*1e-3)
            print(json.dumps({
                'kernel': 'k_sinegrid' if mid == 4 else 'k_model_chisq', 'model': NAMES[mid], 'dtype': dt,
                'uncertainties': 'uniform' if usig else 'per point',
                'abscissa': 'uniform grid' if mid == 4 else 'jittered (general)',
                'nchains': NCH, 'ndata': N, 'ms_per_launch': ms,
                'chain_points_per_s': NCH*N/(ms*1e-3),
                'roofline': {'bound': 'fp64' if dt == 'f64' else 'fp32', 'achieved': ach/1e12,
                             'peak': PEAK[code]/1e12, 'unit': 'TFLOP/s', 'frac': ach/PEAK[code],
                             'algorithmic_flops_per_chain_point': FLOPS[mid],
                             'peak_source': 'measured live: mc3b_fma_peak'},
                'l2': 'flushed between launches (256 MB write)'}), flush=True)
