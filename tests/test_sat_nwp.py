"""SURVEY.md section 8f rank 1 -- the two-tower model (predict_pv_yield/models/conv3d/model_sat_nwp.py).

CPU: the oracle restatement (oracle/sat_nwp_oracle.py) is pinned bit-for-bit against outputs of the UNMODIFIED reference
(tests/golden/sat_nwp_*.npz, oracle/make_golden.py).  GPU: the CUDA path (time-padded fp32 convolutions, generic Linear
/ embedding / history kernels through the C ABI) against those goldens and the oracle, fp32 tolerance 1e-5 (normalised).
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import conv3d_oracle as O
from oracle.golden_cases import SAT_NWP_CASES, golden_state_dict, sat_nwp_batch, thin
from oracle.sat_nwp_oracle import OracleSatNwpModel


@pytest.fixture(autouse=True)
def _one_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)


def _nerr_np(a, b):
    scale = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max()) / scale


# ---------------------------------------------------------------------------------------------------------------------
# CPU
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(SAT_NWP_CASES))
def test_oracle_matches_reference(golden_dir, name):
    g = dict(np.load(os.path.join(golden_dir, f"{name}.npz")))
    case = SAT_NWP_CASES[name]
    m = OracleSatNwpModel(**case["model"])
    m.batch_size = case["batch"]
    m.load_state_dict(golden_state_dict(m))
    batch = sat_nwp_batch(name)
    opt = m.configure_optimizers()
    for step in range(2):
        opt.zero_grad()
        r = m.step_losses(batch)
        r["nmae"].backward()
        if step == 0:
            assert np.array_equal(r["y_hat"].detach().numpy(), g["y_hat"])  # same torch ops, same order, same bits
            for k in ("nmae", "mse", "mse_exp", "mae_exp"):
                assert np.float32(r[k].item()) == g[k], k
            for k, p in m.named_parameters():
                assert _nerr_np(thin(p.grad), g["grad." + k]) <= 1e-6, k
        opt.step()
    for k, p in m.named_parameters():
        ref = g["adam2." + k]
        assert float(np.abs(thin(p) - ref).max()) <= 1e-6 * max(float(np.abs(ref).max()), 1e-30) + 1e-9, k


def test_mirror_has_the_reference_state_dict_layout():
    """Same keys, shapes and derived sizes as the oracle (itself constructed like the reference), no CUDA needed."""
    from predict_pv_yield_b200.models.conv3d.model_sat_nwp import Model

    for case in SAT_NWP_CASES.values():
        m, o = Model(**case["model"]), OracleSatNwpModel(**case["model"])
        assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, tuple(v.shape)) for k, v in o.state_dict().items()]
        assert (m.cnn_output_size, m.nwp_cnn_output_size, m.forecast_len) == (o.cnn_output_size, o.nwp_cnn_output_size, o.forecast_len)
    d = Model()  # reference defaults (model_sat_nwp.py:18-37)
    assert d.name == "conv3d_sat_nwp" and d.cnn_output_size == 32 * 56 * 56 * 19 and d.nwp_cnn_output_size == 32 * 56 * 56 * 2
    with pytest.raises(RuntimeError, match="CUDA"):
        Model(**SAT_NWP_CASES["sat_nwp_gsp"]["model"])(sat_nwp_batch("sat_nwp_gsp"))


# ---------------------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 5, 4, 9, 10, 16), (1, 32, 3, 12, 12, 32), (2, 12, 1, 8, 8, 8)])
def test_time_padded_conv_kernels(dev, shape):
    """Conv3d 3x3x3 with padding (1, 0, 0): forward, data gradient (with ReLU mask), weight / bias gradient vs torch fp64."""
    from predict_pv_yield_b200 import ops

    B, Ci, T, H, W, Co = shape
    g = torch.Generator().manual_seed(7)
    x = torch.randn((B, Ci, T, H, W), generator=g)
    w = torch.randn((Co, Ci, 3, 3, 3), generator=g) / np.sqrt(Ci * 27)
    b = torch.randn((Co,), generator=g) * 0.1
    gz = torch.randn((B, Co, T, H - 2, W - 2), generator=g)
    xd = x.double().requires_grad_(True)
    wd = w.double().requires_grad_(True)
    bd = b.double().requires_grad_(True)
    pre = F.conv3d(xd, wd, bd, padding=(1, 0, 0))
    pre.backward(gz.double())
    got = ops.conv3d_fwd(x.to(dev), w.to(dev), b.to(dev), relu=True, pad_t=1)
    assert O.normalised_max_err(got, F.relu(pre.detach())) <= 1e-5
    gx = ops.conv3d_dgrad(gz.to(dev), w.to(dev), x.to(dev), x.shape, pad_t=1)
    assert O.normalised_max_err(gx, xd.grad * (x > 0).double()) <= 1e-4
    dw, db = ops.conv3d_wgrad(x.to(dev), gz.to(dev), pad_t=1)
    assert O.normalised_max_err(dw, wd.grad) <= 1e-4
    assert O.normalised_max_err(db, bd.grad) <= 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("case", [(3, 40, 24, True, False), (2, 700, 128, True, True), (4, 64, 12, False, False),
                                  (3, 14400, 32, True, True), (5, 9000, 20, True, True)])
def test_linear_fn(dev, case):
    """Generic Linear (+ReLU) forward / backward through the C ABI (small path and weight-streaming path) vs torch fp64."""
    from predict_pv_yield_b200 import ops

    B, K, N, relu, mask_input = case
    g = torch.Generator().manual_seed(8)
    x = torch.randn((B, K), generator=g)
    if mask_input:
        x = F.relu(x)
    w = torch.randn((N, K), generator=g) / np.sqrt(K)
    b = torch.randn((N,), generator=g) * 0.1
    gy = torch.randn((B, N), generator=g)
    xd, wd, bd = x.double().requires_grad_(True), w.double().requires_grad_(True), b.double().requires_grad_(True)
    yd = F.linear(xd, wd, bd)
    yd = F.relu(yd) if relu else yd
    yd.backward(gy.double())
    xg = x.to(dev).requires_grad_(True)
    wg, bg = w.to(dev).requires_grad_(True), b.to(dev).requires_grad_(True)
    y = ops.LinearFn.apply(xg, wg, bg, relu, mask_input)
    y.backward(gy.to(dev))
    assert O.normalised_max_err(y.detach(), yd.detach()) <= 1e-5
    want_gx = xd.grad * ((x > 0).double() if mask_input else 1.0)
    assert O.normalised_max_err(xg.grad, want_gx) <= 1e-5
    assert O.normalised_max_err(wg.grad, wd.grad) <= 1e-5
    assert O.normalised_max_err(bg.grad, bd.grad) <= 1e-5


@pytest.mark.gpu
def test_embedding_and_history_kernels(dev):
    from predict_pv_yield_b200 import ops

    g = torch.Generator().manual_seed(9)
    table = torch.randn((940, 16), generator=g)
    ids = torch.tensor([5, 939, 5, 0, 17], dtype=torch.int32)
    gy = torch.randn((5, 16), generator=g)
    td = table.double().requires_grad_(True)
    F.embedding(ids.long(), td).backward(gy.double())
    tg = table.to(dev).requires_grad_(True)
    y = ops.EmbeddingFn.apply(tg, ids.to(dev))
    y.backward(gy.to(dev))
    assert torch.equal(y.detach().cpu(), table[ids.long()])
    assert O.normalised_max_err(tg.grad, td.grad) <= 1e-6
    h = torch.rand((3, 7, 40), generator=g)
    h[0, 0, 3] = float("nan")
    h[2, 1, 0] = float("nan")
    got = ops.history_flatten(h.to(dev)[:, :, :33], 2, 33)  # a strided view: [:, :2, :33] of a wider tensor
    assert torch.equal(got.cpu(), h[:, :2, :33].nan_to_num(nan=0.0).reshape(3, -1))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(SAT_NWP_CASES))
def test_model_matches_reference_golden_and_oracle(dev, golden_dir, name):
    from predict_pv_yield_b200.models.conv3d.model_sat_nwp import Model

    g = dict(np.load(os.path.join(golden_dir, f"{name}.npz")))
    case = SAT_NWP_CASES[name]
    m = Model(**case["model"]).to(dev)
    m.batch_size = case["batch"]
    sd = golden_state_dict(m)
    m.load_state_dict(sd)
    batch = O.batch_to(sat_nwp_batch(name), dev)
    loss = m.training_step(batch, 0)
    loss.backward()
    with torch.no_grad():
        y_hat = m(batch)
    assert _nerr_np(y_hat.cpu().numpy(), g["y_hat"]) <= 1e-5
    assert abs(float(loss.detach()) - float(g["nmae"])) <= 1e-5 * abs(float(g["nmae"]))
    # gradients against the fp64 oracle (identical inputs); conv gradients carry the ReLU-flip noise of any two fp32
    # implementations (tests/test_gpu_model.py), the dense layers do not
    om = OracleSatNwpModel(**case["model"]).double()
    om.batch_size = case["batch"]
    om.load_state_dict({k: v.double() for k, v in sd.items()})
    om.step_losses(O.batch_to(sat_nwp_batch(name), float_dtype=torch.float64))["nmae"].backward()
    for (k, p), (_, q) in zip(m.named_parameters(), om.named_parameters()):
        assert p.grad is not None, k
        assert O.normalised_max_err(p.grad, q.grad) <= (2e-2 if "conv" in k else 2e-3), k
    # one optimiser step through FusedAdam keeps every parameter finite and moves it
    opt = m.configure_optimizers()
    before = {k: p.detach().clone() for k, p in m.named_parameters()}
    opt.step()
    for k, p in m.named_parameters():
        assert torch.isfinite(p).all() and not torch.equal(p, before[k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(SAT_NWP_CASES))
def test_bf16_model_within_tolerance_of_oracle(dev, name):
    """precision="bf16": both towers and their fc1 on the tensor cores; loss / forecast within 2e-2 of the torch reference."""
    from predict_pv_yield_b200.models.conv3d.model_sat_nwp import Model

    case = SAT_NWP_CASES[name]
    m = Model(**case["model"], precision="bf16").to(dev)
    m.batch_size = case["batch"]
    sd = golden_state_dict(m)
    m.load_state_dict(sd)
    om = OracleSatNwpModel(**case["model"])
    om.batch_size = case["batch"]
    om.load_state_dict(sd)
    batch = sat_nwp_batch(name)
    r = om.step_losses(batch)
    r["nmae"].backward()
    dbatch = O.batch_to(batch, dev)
    opt = m.configure_optimizers()
    loss = m.training_step(dbatch, 0)
    loss.backward()
    with torch.no_grad():
        y_hat = m(dbatch)
    assert O.normalised_max_err(y_hat, r["y_hat"].detach()) <= 2e-2
    assert abs(float(loss.detach()) - float(r["nmae"].detach())) <= 2e-2 * abs(float(r["nmae"].detach()))
    for (k, p), (_, q) in zip(m.named_parameters(), om.named_parameters()):
        assert p.grad is not None and torch.isfinite(p.grad).all(), k
        # bf16 activations / gradients through up to three layers on a 3-sample batch: dense layers to 2e-2, conv
        # gradients (ReLU flips + bf16 rounding of tiny sums) to 2e-1 -- a sanity gate; the kernels themselves are gated
        # at bf16 rounding level in tests/test_gpu_bf16.py::test_conv3d_bf16_time_padded
        assert O.normalised_max_err(p.grad, q.grad) <= (2e-1 if "conv" in k else 2e-2), k
    # two optimiser steps through the fused Adam + shadow path keep the forward consistent with a fresh model
    for step in range(2):
        opt.zero_grad()
        m.training_step(dbatch, step).backward()
        opt.step()
    ref = Model(**case["model"], precision="bf16").to(dev)
    ref.load_state_dict(m.state_dict())
    with torch.no_grad():
        assert torch.equal(m(dbatch), ref(dbatch))


def _production_yaml():
    import yaml

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = yaml.safe_load(open(os.path.join(root, "configs", "model", "conv3d_sat_nwp.yaml")))
    target = cfg.pop("_target_")
    return target, cfg


def test_production_yaml_instantiates_the_mirror():
    """configs/model/conv3d_sat_nwp.yaml: the reference's production keys (configs/model/conv3d_sat_nwp.yaml:3-16) with
    only `_target_` changed; the mirror and the oracle agree on every derived size."""
    import importlib

    target, cfg = _production_yaml()
    mod, cls = target.rsplit(".", 1)
    m = getattr(importlib.import_module(mod), cls)(**cfg)
    o = OracleSatNwpModel(**cfg)
    assert (m.cnn_output_size, m.nwp_cnn_output_size, m.forecast_len) == (o.cnn_output_size, o.nwp_cnn_output_size, o.forecast_len)
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, tuple(v.shape)) for k, v in o.state_dict().items()]
    assert m.cnn_output_size == 32 * 12 * 12 * 31 and m.forecast_len == 4


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("bf16", 2e-2)])
def test_production_yaml_step_matches_oracle(dev, precision, tol):
    """The production config (6 layers, 24x24 satellite, 64x64 NWP, gsp_yield, 31 time steps) on a 2-sample batch."""
    from predict_pv_yield_b200.models.conv3d.model_sat_nwp import Model

    _, cfg = _production_yaml()
    B = 2
    torch.manual_seed(4)
    m = Model(**cfg, precision=precision).to(dev)
    m.batch_size = B
    om = OracleSatNwpModel(**cfg)
    om.batch_size = B
    om.load_state_dict({k: v.cpu() for k, v in m.state_dict().items()})
    rs = np.random.RandomState(8)
    batch = {
        "satellite": {"data": torch.from_numpy(rs.randint(0, 1024, size=(B, 11, 31, 24, 24)).astype(np.int16))},
        "nwp": {"data": torch.from_numpy(rs.randn(B, 10, 4, 64, 64).astype(np.float32))},
        "pv": {"pv_yield": torch.from_numpy(rs.rand(B, 31, 128).astype(np.float32))},
        "gsp": {"gsp_yield": torch.from_numpy(rs.rand(B, 6, 32).astype(np.float32)),
                "gsp_id": torch.from_numpy(rs.randint(0, 338, size=(B, 32)).astype(np.int64))},
    }
    r = om.step_losses(batch)
    loss = m.training_step(O.batch_to(batch, dev), 0)
    loss.backward()
    assert O.normalised_max_err(m(O.batch_to(batch, dev)).detach(), r["y_hat"].detach()) <= tol
    assert abs(float(loss.detach()) - float(r["nmae"].detach())) <= tol * abs(float(r["nmae"].detach()))
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
