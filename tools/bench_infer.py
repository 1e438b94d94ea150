"""BASELINE configs[3]: inference sweep (forecast all GB PV systems), batch 512..8192 on one B200, both precisions.
Inputs resident in HBM, no_grad forward through Model.forward (micro-batched), CUDA-event timed.
python tools/bench_infer.py [--batches 512 1024 2048 4096 8192]"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", type=int, nargs="+", default=[512, 1024, 2048, 4096, 8192])
    args = ap.parse_args()
    from predict_pv_yield_b200.models.conv3d.model import Model

    dev = torch.device("cuda:0")
    kw = dict(include_pv_yield=False, include_nwp=False, forecast_minutes=60, history_minutes=30)
    rs = np.random.RandomState(0)
    base = torch.from_numpy(rs.randint(0, 1024, size=(512, 12, 19, 64, 64)).astype(np.int16)).to(dev)
    print(f"{'precision':<10}{'batch':>7}{'ms':>10}{'samples/s':>12}")
    for precision in ("fp32", "bf16"):
        torch.manual_seed(0)
        m = Model(**kw, precision=precision).to(dev).eval()
        for B in args.batches:
            sat = base.repeat(B // 512, 1, 1, 1, 1) if B > 512 else base[:B]
            batch = {"satellite": {"data": sat}}
            with torch.no_grad():
                m(batch)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                y = m(batch)
                e1.record()
                torch.cuda.synchronize()
            assert y.shape == (B, 12) and torch.isfinite(y).all()
            ms = e0.elapsed_time(e1)
            print(f"{precision:<10}{B:>7}{ms:>10.2f}{B / ms * 1e3:>12.0f}")
            del sat, batch, y


if __name__ == "__main__":
    main()
