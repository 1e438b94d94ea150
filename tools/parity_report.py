"""Print a per-tensor parity table of the CUDA step against (a) the reference golden vectors and (b) the
fp64 oracle, next to torch-fp32's own distance from fp64 (the noise floor of end-to-end gradients:
ReLU-mask flips).  Diagnostic tool; run on the GPU box:

    python tools/parity_report.py [--out profiles/parity_r01.txt]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import conv3d_oracle as O  # noqa: E402
from oracle.golden_cases import CASES, golden_batch, golden_state_dict, thin  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from predict_pv_yield_b200.models.conv3d.model import Model

    dev = torch.device("cuda:0")
    lines = []
    for name, case in CASES.items():
        g = dict(np.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz")))
        m = Model(**case["model"]).to(dev)
        m.batch_size = case["batch"]
        sd = golden_state_dict(m)
        m.load_state_dict(sd)
        o32 = O.OracleModel(**case["model"])
        o32.batch_size = case["batch"]
        o32.load_state_dict(sd)
        o64 = O.OracleModel(**case["model"]).double()
        o64.batch_size = case["batch"]
        o64.load_state_dict({k: v.double() for k, v in sd.items()})
        batch = golden_batch(name)
        loss = m.training_step(O.batch_to(batch, dev), 0)
        loss.backward()
        r32 = o32.step_losses(batch)
        r32["nmae"].backward()
        r64 = o64.step_losses(O.batch_to(batch, float_dtype=torch.float64))
        r64["nmae"].backward()
        with torch.no_grad():
            y = m(O.batch_to(batch, dev))
        lines.append(f"== {name}: loss cuda {float(loss):.8f} ref-golden {float(g['nmae']):.8f} fp64 {float(r64['nmae']):.10f}")
        lines.append(f"   y_hat: cuda-vs-fp64 {O.normalised_max_err(y, r64['y_hat']):.2e}   torch32-vs-fp64 "
                     f"{O.normalised_max_err(r32['y_hat'], r64['y_hat']):.2e}")
        lines.append(f"   {'grad':<18}{'cuda-vs-fp64':>14}{'torch32-vs-fp64':>17}{'cuda-vs-golden':>16}")
        for (k, p), (_, q32), (_, q64) in zip(m.named_parameters(), o32.named_parameters(), o64.named_parameters()):
            e_c = O.normalised_max_err(p.grad, q64.grad)
            e_t = O.normalised_max_err(q32.grad, q64.grad)
            gg = g["grad." + k]
            e_g = float(np.abs(thin(p.grad).astype(np.float64) - gg).max()) / max(float(np.abs(gg).max()), 1e-30)
            lines.append(f"   {k:<18}{e_c:>14.2e}{e_t:>17.2e}{e_g:>16.2e}")
    text = "\n".join(lines)
    print(text)
    if args.out:
        os.makedirs(os.path.dirname(os.path.join(ROOT, args.out)), exist_ok=True)
        with open(os.path.join(ROOT, args.out), "w") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main()
