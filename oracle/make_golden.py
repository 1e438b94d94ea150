"""TEST INFRASTRUCTURE ONLY -- generate ``tests/golden/*.npz`` by running the UNMODIFIED reference.

Run in the build container (where ``/root/reference`` exists):

    python oracle/make_golden.py [case ...]

For every case in ``oracle/golden_cases.py`` it imports the reference ``Model``
(``/root/reference/predict_pv_yield/models/conv3d/model.py``, through ``oracle/ref_shims.py``),
loads the deterministic numpy weights, and records, in fp32 on CPU:

* ``y_hat``             -- reference ``forward`` applied to the normalised cube (the reference model only
                           casts, ``model.py:113``; normalisation is ``netcdf_dataset.py:96-101``)
* ``nmae/mse/mse_exp/mae_exp`` -- the four logged losses of ``base_model.py:98-103``
* ``grad.<param>``      -- gradients of the returned loss (``nmae``) after ``backward()``
* ``adam2.<param>``     -- parameters after TWO ``torch.optim.Adam(lr=5e-4)`` steps (``base_model.py:255-257``)

Large tensors are thinned (every 97th element) to keep the fixtures small.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_shims  # noqa: E402
from oracle.conv3d_oracle import sat_constants, sat_normalise_numpy  # noqa: E402
from oracle.golden_cases import (CASES, MAXPOOL_CASES, SAT_NWP_CASES, golden_batch, golden_state_dict,  # noqa: E402
                                 maxpool_inputs, sat_nwp_batch, thin)

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def run_case(name: str) -> dict:
    if name in SAT_NWP_CASES:  # models/conv3d/model_sat_nwp.py
        Model, case, batch = ref_shims.import_reference_sat_nwp_model(), SAT_NWP_CASES[name], sat_nwp_batch(name)
    else:
        Model, case, batch = ref_shims.import_reference_model(), CASES[name], golden_batch(name)
    torch.manual_seed(0)
    model = Model(**case["model"])
    model.batch_size = case["batch"]  # base_model.py:30,95 (class default 32 truncates targets)
    model.load_state_dict(golden_state_dict(model))
    # normalise exactly as netcdf_dataset.py:96-101 does, in numpy, then hand the float cube to the reference
    sat = batch["satellite"]["data"].numpy()
    mean, std = sat_constants(sat.shape[1])
    batch["satellite"]["data"] = torch.from_numpy(sat_normalise_numpy(sat, mean, std))

    out = {}
    opt = model.configure_optimizers()
    for step in range(2):
        opt.zero_grad()
        loss = model.training_step(batch, step)
        loss.backward()
        if step == 0:
            with torch.no_grad():
                out["y_hat"] = model(batch).numpy().copy()
            logged = model._last_logged
            out["nmae"] = np.float32(logged["NMAE/Train"].item())
            out["mse"] = np.float32(logged["MSE/Train"].item())
            out["mse_exp"] = np.float32(logged["MSE_EXP/Train"].item())
            out["mae_exp"] = np.float32(logged["MAE_EXP/Train"].item())
            out["loss"] = np.float32(loss.item())
            for k, p in model.named_parameters():
                out["grad." + k] = thin(p.grad)
        opt.step()
    for k, p in model.named_parameters():
        out["adam2." + k] = thin(p)
    return out


def run_maxpool_case(name: str) -> dict:
    """The unmodified reference Conv3dMaxPool (perceiver_conv3d_nwp_sat.py:42-57): output, input / weight / bias gradients."""
    Block = ref_shims.import_reference_conv3d_maxpool()
    B, Ci, T, H, W, Co = MAXPOOL_CASES[name]
    m = Block(out_channels=Co, in_channels=Ci)
    m.load_state_dict(golden_state_dict(m))
    x, g = maxpool_inputs(name)
    x.requires_grad_(True)
    y = m(x)
    y.backward(g)
    return {"y": y.detach().numpy().copy(), "gx": x.grad.numpy().copy(), "dw": m.sat_conv3d.weight.grad.numpy().copy(),
            "db": m.sat_conv3d.bias.grad.numpy().copy()}


def normalise_digest() -> str:
    """sha256 over the fp32 bits of the normalisation of ALL 65536 int16 values x 12 channels."""
    x = np.arange(-32768, 32768, dtype=np.int32).astype(np.int16)
    cube = np.broadcast_to(x.reshape(1, 1, 1, 256, 256), (1, 12, 1, 256, 256)).copy()
    y = sat_normalise_numpy(cube, *sat_constants(12))
    return hashlib.sha256(y.tobytes()).hexdigest()


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)  # the reference's conv wgrad is thread-count dependent at the 1e-5 level
    only = sys.argv[1:]  # optional: names of the cases to (re)generate; default = every fixture
    if only:
        for name in only:
            res = run_maxpool_case(name) if name in MAXPOOL_CASES else run_case(name)
            np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **res)
            print(name, "written")
        return
    for name in list(CASES) + list(SAT_NWP_CASES):
        res = run_case(name)
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **res)
        print(name, "y_hat", res["y_hat"].shape, "loss", float(res["loss"]))
    for name in MAXPOOL_CASES:
        res = run_maxpool_case(name)
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **res)
        print(name, "y", res["y"].shape)
    with open(os.path.join(OUT, "normalise_sha256.txt"), "w") as f:
        f.write(normalise_digest() + "\n")
    print("normalise digest written")


if __name__ == "__main__":
    main()
