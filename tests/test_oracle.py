"""CPU: pin the oracle (oracle/conv3d_oracle.py) against outputs of the UNMODIFIED reference.

The fixtures in tests/golden/ were produced by oracle/make_golden.py, which runs the reference
``Model`` (predict_pv_yield/models/conv3d/model.py) and step (base_model.py:78-153,255-257).
"""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import conv3d_oracle as O
from oracle.golden_cases import CASES, golden_batch, golden_state_dict, thin


@pytest.fixture(autouse=True)
def _one_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)  # fixtures were generated single-threaded (conv wgrad summation order)
    yield
    torch.set_num_threads(n)


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, f"{name}.npz")))


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_forward_and_losses(golden_dir, name):
    g = _load(golden_dir, name)
    case = CASES[name]
    m = O.OracleModel(**case["model"])
    m.batch_size = case["batch"]
    m.load_state_dict(golden_state_dict(m))
    batch = golden_batch(name)
    with torch.no_grad():
        r = m.step_losses(batch)
    # same torch ops in the same order on the same bits -> bit-identical
    assert np.array_equal(r["y_hat"].numpy(), g["y_hat"])
    for k in ("nmae", "mse", "mse_exp", "mae_exp"):
        assert np.float32(r[k].item()) == g[k], k
    assert g["loss"] == g["nmae"]  # the returned loss is the L1 one (base_model.py:99,146)


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_grads_and_adam(golden_dir, name):
    g = _load(golden_dir, name)
    case = CASES[name]
    m = O.OracleModel(**case["model"])
    m.batch_size = case["batch"]
    m.load_state_dict(golden_state_dict(m))
    batch = golden_batch(name)
    opt = m.configure_optimizers()
    for step in range(2):
        opt.zero_grad()
        loss = m.training_step(batch, step)
        loss.backward()
        if step == 0:
            for k, p in m.named_parameters():
                ref = g["grad." + k]
                got = thin(p.grad)
                scale = max(float(np.abs(ref).max()), 1e-30)
                assert float(np.abs(got - ref).max()) / scale <= 1e-6, k
        opt.step()
    for k, p in m.named_parameters():
        ref = g["adam2." + k]
        got = thin(p)
        assert float(np.abs(got - ref).max()) <= 1e-6 * max(float(np.abs(ref).max()), 1e-30) + 1e-9, k


def test_state_dict_keys_and_shapes_default_model():
    """SURVEY.md section 8b state_dict contract (probe of the reference): names, shapes, total params."""
    m = O.OracleModel(include_pv_yield=False, include_nwp=False, forecast_minutes=60, history_minutes=30)
    sd = m.state_dict()
    assert tuple(sd["sat_conv0.weight"].shape) == (32, 12, 3, 3, 3)
    assert tuple(sd["conv3d_3.weight"].shape) == (32, 32, 3, 3, 3)
    assert tuple(sd["fc1.weight"].shape) == (128, 1103872)
    assert m.cnn_output_size == 1103872
    assert sum(p.numel() for p in m.parameters()) == 141414732


def test_derived_sizes():
    d = O.derived_sizes(30, 60, "pv_yield")
    assert (d["history_len_5"], d["forecast_len_5"], d["history_len_30"], d["forecast_len_30"]) == (6, 12, 1, 2)
    assert (d["history_len_60"], d["forecast_len_60"], d["forecast_len"]) == (1, 1, 12)
    assert d["number_of_samples_per_batch"] == 128
    d = O.derived_sizes(30, 120, "gsp_yield")
    assert d["forecast_len"] == 4 and d["number_of_samples_per_batch"] == 32
    assert O.derived_sizes(90, 60)["history_len_60"] == 2  # ceil (base_model.py:57)


def test_normalise_exhaustive_digest(golden_dir):
    """All 65536 int16 values x 12 channels; numpy (netcdf_dataset.py:96-101 verbatim) == torch, and
    both equal the committed digest."""
    x = np.arange(-32768, 32768, dtype=np.int32).astype(np.int16)
    cube = np.broadcast_to(x.reshape(1, 1, 1, 256, 256), (1, 12, 1, 256, 256)).copy()
    mean, std = O.sat_constants(12)
    y_np = O.sat_normalise_numpy(cube, mean, std)
    y_t = O.sat_normalise(torch.from_numpy(cube), torch.from_numpy(mean), torch.from_numpy(std)).numpy()
    assert np.array_equal(y_np.view(np.uint32), y_t.view(np.uint32))
    want = open(os.path.join(golden_dir, "normalise_sha256.txt")).read().strip()
    assert hashlib.sha256(y_np.tobytes()).hexdigest() == want
    # spot values: (x - mean) / std with two fp32 roundings
    c, v = 3, 517
    exp = np.float32(np.float32(np.float32(v) - mean[c]) / std[c])
    assert y_np[0, c, 0, (v + 32768) // 256, (v + 32768) % 256] == exp


def test_adam_numpy_matches_torch():
    rs = np.random.RandomState(0)
    p0 = rs.randn(1000).astype(np.float32)
    p = torch.nn.Parameter(torch.from_numpy(p0.copy()))
    opt = torch.optim.Adam([p], lr=5e-4)
    pn, m, v = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    for step in range(1, 4):
        g = rs.randn(1000).astype(np.float32)
        p.grad = torch.from_numpy(g.copy())
        opt.step()
        O.adam_step_numpy(pn, g, m, v, step)
        assert np.abs(pn - p.detach().numpy()).max() <= 2e-7
