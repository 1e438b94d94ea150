"""Debug helper: run the bf16 igemm forward / data gradient for one shape and compare with torch (one process per case so a
hang can be bounded by `timeout`).  python tools/ig_debug.py fwd|dgrad B Ci T H W Co"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from predict_pv_yield_b200 import lib, ops  # noqa: E402

which = sys.argv[1]
B, Ci, T, H, W, Co = [int(v) for v in sys.argv[2:8]]
lib.load()
dev = torch.device("cuda:0")
r16 = lambda t: t.bfloat16().float()  # noqa: E731
g = torch.Generator().manual_seed(1)
w = torch.randn((Co, Ci, 3, 3, 3), generator=g) / np.sqrt(Ci * 27)
if which == "fwd":
    x = r16(torch.randn((B, Ci, T, H, W), generator=g))
    b = torch.randn((Co,), generator=g) * 0.1
    want = F.relu(F.conv3d(x.double(), r16(w).double(), b.double()))
    got = ops.from_blocked_bf16(ops.conv3d_fwd_bf16(ops.to_blocked_bf16(x.to(dev)), w.to(dev), b.to(dev)), Co)
else:
    gz = r16(torch.randn((B, Co, T - 2, H - 2, W - 2), generator=g))
    xd = torch.zeros((B, Ci, T, H, W), dtype=torch.float64, requires_grad=True)
    F.conv3d(xd, r16(w).double(), None).backward(gz.double())
    want = xd.grad
    got = ops.from_blocked_bf16(ops.conv3d_dgrad_bf16(ops.to_blocked_bf16(gz.to(dev), pad=2), w.to(dev), None), Ci)
torch.cuda.synchronize()
err = float((got.cpu().double() - want).abs().max() / want.abs().max())
print(which, sys.argv[2:8], "err", f"{err:.3e}")
