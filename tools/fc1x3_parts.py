"""Backward of the fp32 head with parts of the tensor-core fc1 kernels switched off (run under ncu for per-kernel times)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from predict_pv_yield_b200 import lib, ops  # noqa: E402
L = lib.load()
L.pvb200_debug_set_fc1x3.argtypes = [ctypes.c_int]
L.pvb200_debug_set_fc1x3.restype = None
dev = torch.device("cuda:0")
B, K1, F1 = 32, 32 * 11 * 56 * 56, 128
g = torch.Generator(device=dev).manual_seed(0)
feats = torch.relu(torch.randn((B, K1), device=dev, generator=g)).requires_grad_(True)
mk = lambda o, i: (torch.randn((o, i), device=dev, generator=g) / i ** 0.5).requires_grad_(True)  # noqa: E731
mb = lambda o: torch.zeros((o,), device=dev).requires_grad_(True)  # noqa: E731
params = [mk(F1, K1), mb(F1), mk(128, F1), mb(128), None, None, mk(64, 128), mb(64), mk(12, 64), mb(12)]
gout = torch.randn((B, 12), device=dev, generator=g)
for flags in [int(a) for a in sys.argv[1:]] or [0]:
    L.pvb200_debug_set_fc1x3(1 | (flags << 1))
    for it in range(2):
        ops.HeadFn.apply(feats, None, None, *params).backward(gout)
    torch.cuda.synchronize()
L.pvb200_debug_set_fc1x3(1)
