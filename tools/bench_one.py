"""Run one bf16 conv kernel repeatedly at a full layer shape (for ncu captures): python tools/bench_one.py fwd|dgrad [layer]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from predict_pv_yield_b200 import lib, ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
layer = int(sys.argv[2]) if len(sys.argv) > 2 else 1
lib.load()
dev = torch.device("cuda:0")
B = 32
T, S, C = 19, 64, 12
for _ in range(layer):
    C, T, S = 32, T - 2, S - 2
w = torch.randn(32, C, 3, 3, 3, device=dev) / (C * 27) ** 0.5
b = torch.randn(32, device=dev)
x = torch.randn(B, C, T, S, S, device=dev)
gz = torch.randn(B, 32, T - 2, S - 2, S - 2, device=dev)
xb = ops.to_blocked_bf16(x)
gzp = ops.to_blocked_bf16(gz, pad=2)
for _ in range(4):
    if which == "fwd":
        ops.conv3d_fwd_bf16(xb, w, b)
    else:
        ops.conv3d_dgrad_bf16(gzp, w, xb)
torch.cuda.synchronize()
