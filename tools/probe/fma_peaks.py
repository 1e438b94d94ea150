import ctypes as C, torch, sys
sys.path.insert(0, '.')
from predict_pv_yield_b200 import lib
L = lib.load()
sink = torch.zeros(4, device='cuda'); flops = C.c_double(0.0); st = torch.cuda.current_stream().cuda_stream
for name in ('pvb200_probe_fp32_fma', 'pvb200_probe_fp32_fma2'):
    best = 0
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); lib.check(getattr(L, name)(sink.data_ptr(), 4096, C.byref(flops), st), 'p'); e1.record(); torch.cuda.synchronize()
        best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    print(name, round(best, 2), 'TFLOP/s')
