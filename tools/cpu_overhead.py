"""How long does the HOST need to enqueue one train step (no synchronisation), against the GPU time of the step?
python tools/cpu_overhead.py [fp32|bf16]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import conv3d_oracle as O  # noqa: E402
from predict_pv_yield_b200.models.conv3d.model import Model  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
dev = torch.device("cuda:0")
kw = dict(include_pv_yield=False, include_nwp=False, forecast_minutes=60, history_minutes=30)
torch.manual_seed(0)
m = Model(**kw, precision=precision).to(dev)
m.batch_size = 32
opt = m.configure_optimizers()
batch = O.batch_to(O.make_synthetic_batch(32, seed=1, include_legacy_keys=False), dev)


def step(i):
    opt.zero_grad()
    loss = m.training_step(batch, i)
    loss.backward()
    opt.step()


for i in range(5):
    step(i)
torch.cuda.synchronize()
n = 30
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for i in range(n):
    step(i)
e1.record()
t_enqueue = (time.perf_counter() - t0) / n * 1e3
torch.cuda.synchronize()
print(f"{precision}: host enqueue {t_enqueue:.2f} ms/step, GPU {e0.elapsed_time(e1) / n:.2f} ms/step")
