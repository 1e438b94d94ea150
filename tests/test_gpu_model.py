"""GPU parity tests of the whole step (Model.forward / training_step / backward / FusedAdam) through the
C ABI, against (a) the committed golden vectors produced by the UNMODIFIED reference
(tests/golden/*.npz, oracle/make_golden.py) and (b) the fp64 oracle on the same seeded inputs.

fp32-mode tolerances (normalised max error): forecast / loss <= 1e-5 (north star).
END-TO-END gradients are gated looser than the isolated kernels (1e-4, test_gpu_kernels.py): any two fp32
implementations differ in a few ReLU decisions near zero, and on these tiny batches one flipped mask moves a
conv weight-gradient by up to ~6e-3 of max|g| -- torch-fp32 itself sits that far from torch-fp64 on these
very cases (tools/parity_report.py prints the table; profiles/parity_r01.txt).  Gates: conv grads 2e-2,
fc grads 2e-3, of max|g| per tensor.
"""
import os

import numpy as np
import pytest
import torch

from oracle import conv3d_oracle as O
from oracle.golden_cases import CASES, golden_batch, golden_state_dict, thin

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _model(kw, dev):
    from predict_pv_yield_b200.models.conv3d.model import Model

    return Model(**kw).to(dev)


def _grad_tol(name):
    return 2e-2 if "conv" in name else 2e-3


def _nerr_np(a, b):
    scale = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max()) / scale


@pytest.mark.parametrize("name", list(CASES))
def test_model_matches_reference_golden(dev, golden_dir, name):
    g = dict(np.load(os.path.join(golden_dir, f"{name}.npz")))
    case = CASES[name]
    m = _model(case["model"], dev)
    m.batch_size = case["batch"]
    m.load_state_dict(golden_state_dict(m))
    batch = O.batch_to(golden_batch(name), dev)
    opt = m.configure_optimizers()
    for step in range(2):
        opt.zero_grad()
        loss = m.training_step(batch, step)
        loss.backward()
        if step == 0:
            with torch.no_grad():
                y_hat = m(batch)
            assert _nerr_np(y_hat.cpu().numpy(), g["y_hat"]) <= 1e-5
            assert abs(float(loss.detach()) - float(g["nmae"])) <= 1e-5 * abs(float(g["nmae"]))
            logged = m.logged_metrics
            for key, gk in (("MSE/Train", "mse"), ("NMAE/Train", "nmae"), ("MSE_EXP/Train", "mse_exp"), ("MAE_EXP/Train", "mae_exp")):
                assert abs(float(logged[key]) - float(g[gk])) <= 1e-5 * abs(float(g[gk])), key
            for k, p in m.named_parameters():
                e = _nerr_np(thin(p.grad), g["grad." + k])
                assert e <= _grad_tol(k), (k, e)
        opt.step()
    for k, p in m.named_parameters():
        ref = g["adam2." + k]
        got = thin(p)
        # Adam divides by sqrt(v): the update of an element is ~lr * sign(g) after step 1 and depends on g1/g2
        # ratios after step 2, so the end-to-end gradient noise above (relative error O(1) on the SMALL entries of
        # a conv gradient) becomes parameter differences of up to a few lr = 5e-4 on some entries.  Gate: the median
        # entry within 2e-5 (4 % of one lr step), every entry within 2 steps * 2 * lr.  (Adam itself is gated at 5e-7 against
        # torch.optim.Adam on identical gradients in test_gpu_kernels.py::test_fused_adam_matches_torch.)
        diff = np.abs(got.astype(np.float64) - ref.astype(np.float64))
        assert float(np.median(diff)) <= 2e-5, k
        assert float(diff.max()) <= 2.1e-3, k


@pytest.mark.parametrize("name", list(CASES))
def test_model_matches_fp64_oracle(dev, name):
    case = CASES[name]
    m = _model(case["model"], dev)
    m.batch_size = case["batch"]
    sd = golden_state_dict(m)
    m.load_state_dict(sd)
    om = O.OracleModel(**case["model"]).double()
    om.batch_size = case["batch"]
    om.load_state_dict({k: v.double() for k, v in sd.items()})
    batch = golden_batch(name)
    r = om.step_losses(O.batch_to(batch, float_dtype=torch.float64))
    r["nmae"].backward()
    loss = m.training_step(O.batch_to(batch, dev), 0)
    loss.backward()
    assert abs(float(loss.detach()) - float(r["nmae"])) <= 1e-5 * abs(float(r["nmae"]))
    # the same step by torch in fp32 (the reference's own arithmetic): its distance from fp64 is the noise floor of an
    # end-to-end fp32 gradient.  Where no ReLU decision flips, both sit at ~1e-6; where one does (tiny batches: one
    # position is a visible fraction of a conv gradient) torch-fp32 moves by 1e-3 and so may we.  Gate: 1e-5, or three
    # times torch's own distance on that tensor -- a wrong tap / dropped term is orders of magnitude outside either.
    o32 = O.OracleModel(**case["model"])
    o32.batch_size = case["batch"]
    o32.load_state_dict(sd)
    r32 = o32.step_losses(batch)
    r32["nmae"].backward()
    for (k, p), (_, q), (_, q32) in zip(m.named_parameters(), om.named_parameters(), o32.named_parameters()):
        e, floor = O.normalised_max_err(p.grad, q.grad), O.normalised_max_err(q32.grad, q.grad)
        print(f"{name} {k}: cuda-vs-fp64 {e:.2e}  torch32-vs-fp64 {floor:.2e}")
        assert e <= max(1e-5, 3.0 * floor), (k, e, floor)


def test_model_float_input_equals_int16_input(dev):
    """Already-normalised fp32 cubes (the reference's model input) and raw int16 cubes give the same bits."""
    case = CASES["nwp_pv_small"]
    m = _model(case["model"], dev)
    m.load_state_dict(golden_state_dict(m))
    batch = golden_batch("nwp_pv_small")
    sat = batch["satellite"]["data"]
    mean, std = O.sat_constants(sat.shape[1])
    fb = dict(batch)
    fb["satellite"] = {"data": O.sat_normalise(sat, torch.from_numpy(mean), torch.from_numpy(std))}
    with torch.no_grad():
        a = m(O.batch_to(batch, dev))
        b = m(O.batch_to(fb, dev))
    assert torch.equal(a, b)


def test_batch_size_attribute_truncates_target_like_reference(dev):
    """base_model.py:30,95: y[0:batch_size] -- a batch larger than the class default 32 must fail loudly."""
    case = CASES["test_yaml_pv"]
    m = _model(case["model"], dev)
    m.batch_size = 1
    with pytest.raises(RuntimeError, match="batch_size"):
        m.training_step(O.batch_to(golden_batch("test_yaml_pv"), dev), 0)


@pytest.mark.parametrize("B", [4, 32])
def test_full_size_config2_vs_oracle(dev, B):
    """BASELINE config 2 shape (12x19x64x64 int16, sat-only, fp32) at B = 4 and at the benchmarked B = 32 (the oracle's
    forward + backward takes ~1 s on the host cores): forward + backward, size-independent checks on top."""
    kw = dict(include_pv_yield=False, include_nwp=False, forecast_minutes=60, history_minutes=30)
    torch.manual_seed(518)
    om = O.OracleModel(**kw)
    om.batch_size = B
    m = _model(kw, dev)
    m.batch_size = B
    m.load_state_dict(om.state_dict())
    batch = O.make_synthetic_batch(B, seed=518)
    r = om.step_losses(batch)
    r["nmae"].backward()
    loss = m.training_step(O.batch_to(batch, dev), 0)
    loss.backward()
    with torch.no_grad():
        y_hat = m(O.batch_to(batch, dev))
    assert O.normalised_max_err(y_hat, r["y_hat"]) <= 1e-5
    assert abs(float(loss.detach()) - float(r["nmae"])) <= 1e-5 * abs(float(r["nmae"]))
    for (k, p), (_, q) in zip(m.named_parameters(), om.named_parameters()):
        # fp32-vs-fp32 end-to-end: ReLU-mask flips make torch itself differ from fp64 by ~1e-3 of max|g| on
        # the conv weights; isolated kernels are gated at 1e-4 in test_gpu_kernels.py
        assert O.normalised_max_err(p.grad, q.grad) <= _grad_tol(k), k
    # size-independent property: samples are independent -> permuting the batch permutes the forecast
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(B))
    pb = {"satellite": {"data": batch["satellite"]["data"][perm]}, "pv": {"pv_yield": batch["pv"]["pv_yield"][perm]}}
    with torch.no_grad():
        y_perm = m(O.batch_to(pb, dev))
    assert torch.equal(y_perm, y_hat[perm.to(dev)])


def test_inference_sweep_micro_batching(dev):
    """BASELINE config 4 (inference sweep): a no_grad forward over a batch larger than the micro-batch equals the
    concatenation of per-chunk forwards bit for bit (samples are independent), with the NWP / PV-history branches on."""
    case = CASES["nwp_pv_small"]
    m = _model(case["model"], dev)
    m.load_state_dict(golden_state_dict(m))
    m.inference_micro_batch = 5
    B = 13
    b = O.batch_to(O.make_synthetic_batch(B, 12, 19, 16, seed=7), dev)
    with torch.no_grad():
        y = m(b)
        assert y.shape == (B, m.forecast_len)
        m.inference_micro_batch = 256
        y_one = m(b)
    assert torch.equal(y, y_one)


def test_inference_full_size_batch_512(dev):
    """Config 4 at its smallest sweep point, full-size cubes: B = 512 streams through 256-sample micro-batches."""
    kw = dict(include_pv_yield=False, include_nwp=False, forecast_minutes=60, history_minutes=30)
    m = _model(kw, dev)
    g = torch.Generator().manual_seed(1)
    sat = torch.randint(0, 1024, (512, 12, 19, 64, 64), generator=g, dtype=torch.int32).to(torch.int16).to(dev)
    with torch.no_grad():
        y = m({"satellite": {"data": sat}})
        y_head = m({"satellite": {"data": sat[:8]}})
    assert y.shape == (512, 12) and bool(torch.isfinite(y).all())
    assert torch.equal(y[:8], y_head)


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("bf16", 2e-2)])
def test_deep_variant_config5_vs_oracle(dev, precision, tol):
    """BASELINE config 5 (deep variant): 8 conv layers x 32 channels on 128x128 crops, forward + loss vs the oracle
    (B = 1: the oracle needs ~56 GFLOP per sample on the CPU)."""
    from predict_pv_yield_b200.models.conv3d.model import Model

    kw = dict(include_pv_yield=False, include_nwp=False, forecast_minutes=60, history_minutes=30,
              number_of_conv3d_layers=8, image_size_pixels=128)
    torch.manual_seed(5)
    om = O.OracleModel(**kw)
    om.batch_size = 1
    m = Model(**kw, precision=precision).to(dev)
    m.batch_size = 1
    m.load_state_dict(om.state_dict())
    assert m.cnn_output_size == 32 * 112 * 112 * 3
    batch = O.make_synthetic_batch(1, image_size_pixels=128, seed=3)
    r = om.step_losses(batch)
    r["nmae"].backward()
    loss = m.training_step(O.batch_to(batch, dev), 0)
    loss.backward()
    with torch.no_grad():
        y_hat = m(O.batch_to(batch, dev))
    assert O.normalised_max_err(y_hat, r["y_hat"]) <= tol
    assert abs(float(loss.detach()) - float(r["nmae"])) <= tol * abs(float(r["nmae"]))
    assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in m.parameters())
    if precision == "bf16":
        # eight layers of bf16 roundings (and the ReLU decisions they flip) put a bf16 gradient 0.06-0.3 of max|g| from the
        # fp32 oracle's on this case -- the fp64 model that rounds where the bf16 path rounds (oracle.Bf16EmulatedOracle)
        # sits exactly there too (sat_conv0.weight: 0.201 against the device's 0.202).  Gate: against THAT model, 2^-4 or
        # 0.6 of its own distance from the fp32 oracle, per tensor (measured: convolutions 2.8e-2 .. 7.7e-2, fully
        # connected layers <= 4.7e-3, forecast 1.3e-5; profiles/parity_r02_tests.txt).
        oe = O.Bf16EmulatedOracle(**kw).double()
        oe.batch_size = 1
        oe.load_state_dict({k: v.double() for k, v in om.state_dict().items()})
        re = oe.step_losses(O.batch_to(batch, float_dtype=torch.float64))
        re["nmae"].backward()
        worst = []
        for (k, p), (_, q), (_, qe) in zip(m.named_parameters(), om.named_parameters(), oe.named_parameters()):
            e, e32, floor = O.normalised_max_err(p.grad, qe.grad), O.normalised_max_err(p.grad, q.grad), O.normalised_max_err(qe.grad, q.grad)
            gate = max(2.0 ** -4, 0.6 * floor)
            print(f"config5 bf16 {k}: vs bf16-emulating fp64 {e:.2e} (gate {gate:.2e}), vs fp32 oracle {e32:.2e}, emulation vs fp32 oracle {floor:.2e}")
            if e > gate:
                worst.append((k, e, gate))
        ey = O.normalised_max_err(y_hat, re["y_hat"])
        print(f"config5 bf16 forecast vs bf16-emulating fp64 {ey:.2e}")
        assert ey <= 2.0 ** -7
        assert not worst, worst
        return
    # fp32 mode: gradients against the fp64 oracle.  torch's own fp32 step sits 2-4e-3 of max|g| from fp64 on the
    # convolution weights of this case (ReLU decisions that flip under rounding; 1e-6 where none does), and the CUDA
    # path flips OTHER decisions at the same rate: the convolution tensors are gated at three times the largest of
    # torch's distances, the fully connected ones (no flip reaches them) at 1e-5 or three times torch's distance.
    o64 = O.OracleModel(**kw).double()
    o64.batch_size = 1
    o64.load_state_dict({k: v.double() for k, v in om.state_dict().items()})
    r64 = o64.step_losses(O.batch_to(batch, float_dtype=torch.float64))
    r64["nmae"].backward()
    floors = {k: O.normalised_max_err(q.grad, q64.grad) for (k, q), (_, q64) in zip(om.named_parameters(), o64.named_parameters())}
    conv_gate = 3.0 * max(v for k, v in floors.items() if "conv" in k)
    for (k, p), (_, q64) in zip(m.named_parameters(), o64.named_parameters()):
        e = O.normalised_max_err(p.grad, q64.grad)
        gate = conv_gate if "conv" in k else max(1e-5, 3.0 * floors[k])
        print(f"config5 fp32 {k}: cuda-vs-fp64 {e:.2e}  torch32-vs-fp64 {floors[k]:.2e}  gate {gate:.2e}")
        assert e <= gate, (k, e, gate)


def test_device_prefetcher_preserves_batches(dev):
    """Input pipeline: int16 cubes arrive intact and in order, one batch ahead, and feed the model."""
    from predict_pv_yield_b200.data import DevicePrefetcher

    case = CASES["test_yaml_pv"]
    m = _model(case["model"], dev)
    m.batch_size = case["batch"]
    host = [O.make_synthetic_batch(2, 11, 25, 16, seed=s, include_legacy_keys=False) for s in range(5)]
    got = []
    for i, b in enumerate(DevicePrefetcher(iter(host), dev)):
        assert b["satellite"]["data"].is_cuda and b["satellite"]["data"].dtype == torch.int16
        assert torch.equal(b["satellite"]["data"].cpu(), host[i]["satellite"]["data"])
        assert torch.equal(b["pv"]["pv_yield"].cpu(), host[i]["pv"]["pv_yield"])
        with torch.no_grad():
            got.append(m(b).cpu())
    assert len(got) == 5
    with torch.no_grad():
        want = m(O.batch_to(host[3], dev)).cpu()
    assert torch.equal(got[3], want)


def test_overlapped_adam_equals_serial_adam(dev):
    """FusedAdam(overlap_large=True) runs the update of fc1.weight on a side stream under the next forward pass: after three
    steps the weights, the moments and the losses equal the serial optimiser's bit for bit, and state_dict() waits for it."""
    kw = dict(include_pv_yield=False, include_nwp=False, forecast_minutes=60, history_minutes=30, image_size_pixels=24)
    batch = O.batch_to(O.make_synthetic_batch(4, image_size_pixels=24, seed=9), dev)
    out = {}
    for overlap in (False, True):
        torch.manual_seed(3)
        m = _model(kw, dev)
        m.batch_size = 4
        m.overlap_optimizer = overlap
        opt = m.configure_optimizers()
        assert opt.overlap_large is overlap
        losses = []
        for step in range(3):
            opt.zero_grad()
            loss = m.training_step(batch, step)
            loss.backward()
            opt.step()
            losses.append(float(loss.detach()))
        if overlap:
            assert m.fc1.weight.numel() >= opt.large_numel and getattr(m.fc1.weight, "_pvb_ready", None) is not None
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}  # the pre-hook waits for the side stream
        assert getattr(m.fc1.weight, "_pvb_ready", None) is None
        out[overlap] = (losses, sd, opt.state[m.fc1.weight]["exp_avg"].clone())
    assert out[True][0] == out[False][0]
    for k in out[True][1]:
        assert torch.equal(out[True][1][k], out[False][1][k]), k
    assert torch.equal(out[True][2], out[False][2])
