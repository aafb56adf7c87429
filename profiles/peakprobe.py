import ctypes, torch, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mc3_b200 import _lib
sink = torch.zeros(8, dtype=torch.float64, device='cuda'); fl = ctypes.c_double()
def t(fn, *a):
    best = 0
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); _lib.call(fn, *a, sink.data_ptr(), ctypes.byref(fl), _lib.stream_ptr()); e1.record(); torch.cuda.synchronize()
        if it: best = max(best, fl.value/(e0.elapsed_time(e1)*1e-3))
    return best/1e12
print('regs 8ch x 8w/SMSP', t('mc3b_fma_peak', 0, 20000))
for v in (1,2,3,4,5,6):
    print('variant', v, t('mc3b_fma_peak_variant', v, 3000))
