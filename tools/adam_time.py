"""Time the multi-tensor Adam kernel on an fc1.weight-sized tensor (141 M parameters, 28 B each)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from predict_pv_yield_b200 import ops  # noqa: E402
dev = torch.device("cuda:0")
n = 128 * 32 * 11 * 56 * 56
p, g, m, v = (torch.randn(n, device=dev) for _ in range(4))
v.abs_()
for _ in range(3):
    ops.adam_step([p], [g], [m], [v], 5e-4, 0.9, 0.999, 1e-8, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(10):
    ops.adam_step([p], [g], [m], [v], 5e-4, 0.9, 0.999, 1e-8, 2 + i)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"adam_step_f32 on {n / 1e6:.0f} M parameters: {ms:.3f} ms = {28 * n / ms / 1e6:.0f} GB/s")
