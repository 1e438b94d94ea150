"""BASELINE configs[4]: the deep Conv3d variant (8 conv3d layers, 32 channels, 128x128 crops) -- one train step on one
B200, both precisions, loss checked against the CPU oracle on a small batch.  python tools/bench_deep.py [--batch 16]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    from oracle import conv3d_oracle as O
    from predict_pv_yield_b200 import ops
    from predict_pv_yield_b200.models.conv3d.model import Model

    dev = torch.device("cuda:0")
    kw = dict(include_pv_yield=False, include_nwp=False, forecast_minutes=60, history_minutes=30, number_of_conv3d_layers=8,
              conv3d_channels=32, image_size_pixels=128, number_sat_channels=12)
    B = args.batch
    batch = O.make_synthetic_batch(B, image_size_pixels=128, seed=9, include_legacy_keys=False)
    dbatch = O.batch_to(batch, dev)
    ref = None
    for precision in ("fp32", "bf16"):
        torch.manual_seed(1)
        m = Model(**kw, precision=precision).to(dev)
        m.batch_size = B
        if ref is None:  # CPU oracle on 2 samples (the full batch would take minutes)
            om = O.OracleModel(**kw)
            om.batch_size = 2
            om.load_state_dict({k: v.cpu() for k, v in m.state_dict().items()})
            small = O.make_synthetic_batch(2, image_size_pixels=128, seed=9, include_legacy_keys=False)
            with torch.no_grad():
                ref = float(om.step_losses(small)["nmae"])
        m.batch_size = 2
        with torch.no_grad():
            got = float(m.training_step(O.batch_to(O.make_synthetic_batch(2, image_size_pixels=128, seed=9, include_legacy_keys=False), dev), 0))
        m.batch_size = B
        opt = m.configure_optimizers()

        def step(i):
            opt.zero_grad()
            loss = m.training_step(dbatch, i)
            loss.backward()
            opt.step()

        for i in range(2):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            step(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        print(f"deep variant (8 layers, 128x128, batch {B}) {precision}: loss(2 samples) {got:.6f} vs oracle {ref:.6f} "
              f"(rel {abs(got - ref) / abs(ref):.1e}); train step {ms:.2f} ms = {B / ms * 1e3:.0f} samples/s; "
              f"params {sum(p.numel() for p in m.parameters())}; peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
        del m, opt


if __name__ == "__main__":
    main()
