"""Accuracy of the tensor-core weight gradient at the full conv1 shape (32 -> 32 channels, 17x62x62 input, batch 32) against
torch fp64 on the GPU (checker only): the flush window of the toward-zero accumulators matters only when a CTA runs many steps."""
import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from predict_pv_yield_b200 import ops  # noqa: E402
dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
B, Ci, T, S, Co = 32, 32, 17, 62, 32
g = torch.Generator(device=dev).manual_seed(0)
x = torch.relu(torch.randn(B, Ci, T, S, S, device=dev, generator=g))
gz = torch.randn(B, Co, T - 2, S - 2, S - 2, device=dev, generator=g)
wd = torch.zeros(Co, Ci, 3, 3, 3, device=dev, dtype=torch.float64, requires_grad=True)
bd = torch.zeros(Co, device=dev, dtype=torch.float64, requires_grad=True)
for b0 in range(0, B, 4):  # fp64 reference in slices (memory)
    F.conv3d(x[b0:b0 + 4].double(), wd, bd).backward(gz[b0:b0 + 4].double())
am = torch.zeros(2, device=dev)
xb, gzb = ops.to_blocked_f32(x, amax=am[0:1]), ops.to_blocked_f32(gz, pad=2, amax=am[1:2])
nerr = lambda a, b: float((a.double() - b).abs().max() / b.abs().max())  # noqa: E731
for name, kw in (("f16x2", dict(amax=(am[0:1], am[1:2]))), ("bf16x3", {})):
    dw, db = ops.conv3d_wgrad_bf16x3(xb, gzb, Ci, Co, gz_pad=2, **kw)
    print(f"{name}: dw {nerr(dw, wd.grad):.2e}  db {nerr(db, bd.grad):.2e}")
x32 = x.clone().requires_grad_(False)
w32 = torch.zeros(Co, Ci, 3, 3, 3, device=dev, requires_grad=True)
b32 = torch.zeros(Co, device=dev, requires_grad=True)
F.conv3d(x32, w32, b32).backward(gz)
print(f"torch fp32 (cuDNN, TF32 off): dw {nerr(w32.grad, wd.grad):.2e}  db {nerr(b32.grad, bd.grad):.2e}")
