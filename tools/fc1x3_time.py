"""Time the fp32 head (fc1 forward / backward) with fc1 on the tensor cores (fc1_bf16x3.cu) and on the FMA-pipe kernels:
python tools/fc1x3_time.py [B]"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from predict_pv_yield_b200 import lib, ops  # noqa: E402

L = lib.load()
L.pvb200_debug_set_fc1x3.argtypes = [ctypes.c_int]
L.pvb200_debug_set_fc1x3.restype = None
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
K1, F1 = 32 * 11 * 56 * 56, 128
g = torch.Generator(device=dev).manual_seed(0)
feats = torch.relu(torch.randn((B, K1), device=dev, generator=g)).requires_grad_(True)
mk = lambda o, i: (torch.randn((o, i), device=dev, generator=g) / i ** 0.5).requires_grad_(True)  # noqa: E731
mb = lambda o: torch.zeros((o,), device=dev).requires_grad_(True)  # noqa: E731
params = [mk(F1, K1), mb(F1), mk(128, F1), mb(128), None, None, mk(64, 128), mb(64), mk(12, 64), mb(12)]
gout = torch.randn((B, 12), device=dev, generator=g)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for mode in (0, 1):
    L.pvb200_debug_set_fc1x3(mode)
    tf, tb = [], []
    for it in range(6):
        flush.zero_()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        out = ops.HeadFn.apply(feats, None, None, *params)
        e[1].record()
        out.backward(gout)
        e[2].record()
        torch.cuda.synchronize()
        tf.append(e[0].elapsed_time(e[1]))
        tb.append(e[1].elapsed_time(e[2]))
    tf, tb = sorted(tf[2:]), sorted(tb[2:])
    gb = K1 * F1 * 4 / 1e9
    print(f"fc1 {'tensor cores' if mode else 'FMA pipe'}: head fwd {tf[1]:.3f} ms ({gb / tf[1] * 1e3:.0f} GB/s of W), "
          f"head bwd {tb[1]:.3f} ms (W read + dW write {2 * gb / tb[1] * 1e3:.0f} GB/s)")
