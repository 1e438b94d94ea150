"""Time the tensor-core weight gradient (conv3d_wgrad_bf16x3) with parts of the kernel switched off (tools only):
which of bulk copies / split arithmetic / MMAs / accumulator drain bounds a step.  Run on the GPU box."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from predict_pv_yield_b200 import lib, ops  # noqa: E402

L = lib.load()
L.pvb200_debug_set_wgrad_flags.argtypes = [C.c_int]
dev = torch.device("cuda:0")
for (B, Ci, T, H, W, Co) in [(32, 32, 17, 62, 62, 32), (32, 12, 19, 64, 64, 32)]:
    x = torch.randn((B, Ci, T, H, W), device=dev)
    gz = torch.randn((B, Co, T - 2, H - 2, W - 2), device=dev)
    xb, gzb = ops.to_blocked_f32(x), ops.to_blocked_f32(gz, pad=2)
    steps = B * T * (H - 2)
    amax = torch.zeros(2, device=dev)
    ops.absmax_f32(xb, amax[0:1]); ops.absmax_f32(gzb, amax[1:2])
    for variant, kw in (("bf16x3", {}), ("f16x2", dict(amax=(amax[0:1], amax[1:2])))):
        for flags, name in [(0, "full"), (4, "no MMA"), (2 | 4, "copies + drain only"), (1 | 4, "split + drain only"), (1 | 2, "MMA + drain only"),
                            (1 | 2 | 8, "MMA only"), (8, "no drain"), (16, "drain: ld only"), (1 | 2 | 16, "MMA + ld only"), (1 | 2 | 4 | 8, "barriers only")]:
            L.pvb200_debug_set_wgrad_flags(flags)
            for _ in range(2):
                ops.conv3d_wgrad_bf16x3(xb, gzb, Ci, Co, gz_pad=2, **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ops.conv3d_wgrad_bf16x3(xb, gzb, Ci, Co, gz_pad=2, **kw)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"Ci={Ci:2d} {variant} {name:22s} {ms:7.3f} ms  = {ms * 1e-3 * 1.965e9 * 148 / steps:7.0f} clk per step")
        L.pvb200_debug_set_wgrad_flags(0)
    for dyn in (0, 1):
        L.pvb200_set_dynamic_tiles(dyn)
        for _ in range(2):
            ops.conv3d_wgrad_bf16x3(xb, gzb, Ci, Co, gz_pad=2, amax=(amax[0:1], amax[1:2]))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.conv3d_wgrad_bf16x3(xb, gzb, Ci, Co, gz_pad=2, amax=(amax[0:1], amax[1:2]))
        e1.record()
        torch.cuda.synchronize()
        print(f"Ci={Ci:2d} f16x2 {'dynamic chunks' if dyn else 'static split'}: {e0.elapsed_time(e1) / 5:.3f} ms")
    L.pvb200_set_dynamic_tiles(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.absmax_f32(xb, amax[0:1])
    e1.record()
    torch.cuda.synchronize()
    print(f"absmax of x ({xb.numel() * 4 / 1e6:.0f} MB): {e0.elapsed_time(e1) / 5:.3f} ms")
