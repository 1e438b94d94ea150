"""Time one train step (forward, L1 loss, backward, Adam) of the two-tower conv3d_sat_nwp model at the reference's
default constructor sizes (12 x 19 x 64 x 64 satellite, 10 x 3 x 64 x 64 NWP, batch 32) and check the loss against
the CPU oracle.  python tools/bench_sat_nwp.py [--batch 32] [--steps 10] [--no-oracle]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--no-oracle", action="store_true")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"])
    args = ap.parse_args()
    from predict_pv_yield_b200 import ops
    from predict_pv_yield_b200.models.conv3d.model_sat_nwp import Model

    dev = torch.device("cuda:0")
    B = args.batch
    kw = dict(forecast_minutes=30, history_minutes=60)  # defaults: T = 6 + 12 + 1 = 19, NWP T = 0 + 1 + 1 = 2
    torch.manual_seed(0)
    m = Model(**kw, precision=args.precision).to(dev)
    m.batch_size = B
    rs = np.random.RandomState(0)
    sat = torch.from_numpy(rs.randint(0, 1024, size=(B, 12, 19, 64, 64)).astype(np.int16))
    nwp = torch.from_numpy(rs.randn(B, 10, 2, 64, 64).astype(np.float32))
    pv = torch.from_numpy(rs.rand(B, 19, 128).astype(np.float32))
    ids = torch.from_numpy(rs.randint(0, 940, size=(B, 128)).astype(np.int64))
    batch = {"satellite": {"data": sat}, "nwp": {"data": nwp}, "pv": {"pv_yield": pv, "pv_system_row_number": ids}}
    dbatch = {k: {kk: vv.to(dev) for kk, vv in v.items()} for k, v in batch.items()}
    opt = m.configure_optimizers()
    if not args.no_oracle:
        from oracle.sat_nwp_oracle import OracleSatNwpModel

        o = OracleSatNwpModel(**kw)
        o.batch_size = B
        o.load_state_dict({k: v.cpu() for k, v in m.state_dict().items()})
        t0 = time.perf_counter()
        ref = o.step_losses(batch)["nmae"]
        ref.backward()
        t_cpu = time.perf_counter() - t0
        loss = m.training_step(dbatch, 0)
        print(f"loss {float(loss):.7f}  oracle {float(ref):.7f}  rel err {abs(float(loss) - float(ref)) / abs(float(ref)):.2e}  "
              f"(oracle fwd+bwd {t_cpu:.2f} s on {torch.get_num_threads()} threads)")

    def step(i):
        opt.zero_grad()
        loss = m.training_step(dbatch, i)
        loss.backward()
        opt.step()

    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    tm = ops.KernelTimer()
    ops.set_timer(tm)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ops.set_timer(None)
    ms = e0.elapsed_time(e1) / args.steps
    print(f"conv3d_sat_nwp [{args.precision}] train step: {ms:.2f} ms  ({B / ms * 1e3:.0f} samples/s), params {sum(p.numel() for p in m.parameters())}")
    for k, v in sorted(tm.summary().items(), key=lambda kv: -kv[1]["ms"]):
        print(f"  {k:<34}{v['ms'] / args.steps:8.3f} ms/step  {v['flops'] / max(v['ms'], 1e-9) / 1e9:8.1f} TF  {v['bytes'] / max(v['ms'], 1e-9) / 1e6:7.0f} GB/s")


if __name__ == "__main__":
    main()
