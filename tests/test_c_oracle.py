"""CPU tests of the plain-C oracle's backward / padded / pooling / optimiser statements (oracle/c/pv_oracle.c) against the
torch operators the reference delegates to (the arithmetic of the path lives in torch).  The C
code is scalar loops with double accumulation, so it pins operator SEMANTICS (padding, tap orientation of the data
gradient, arg-max tie rule, Adam's order of operations) independently of any library; the CUDA kernels are tested
against torch on the GPU, torch against this on the CPU."""
import ctypes
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g

    g.build()  # also compiles oracle/c (building the checker is not using it)
    return ctypes.CDLL(os.path.join(ROOT, "oracle", "libpv_oracle.so"))


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _close(got, want, tol=2e-6):
    want = np.asarray(want, np.float64)
    assert np.abs(got - want).max() <= tol * max(np.abs(want).max(), 1e-30), np.abs(got - want).max()


@pytest.mark.parametrize("pt,ph", [(0, 0), (1, 0), (1, 1)])
def test_conv3d_forward_and_gradients(lib, pt, ph):
    rs = np.random.RandomState(10 * pt + ph)
    B, Ci, T, H, W, Co = 2, 3, 4, 6, 5, 4
    x = rs.randn(B, Ci, T, H, W).astype(np.float32)
    w = (rs.randn(Co, Ci, 3, 3, 3) / 5).astype(np.float32)
    b = rs.randn(Co).astype(np.float32)
    xd = torch.from_numpy(x).double().requires_grad_(True)
    wd = torch.from_numpy(w).double().requires_grad_(True)
    bd = torch.from_numpy(b).double().requires_grad_(True)
    pre = F.conv3d(xd, wd, bd, padding=(pt, ph, ph))
    To, Ho, Wo = T + 2 * pt - 2, H + 2 * ph - 2, W + 2 * ph - 2
    assert tuple(pre.shape) == (B, Co, To, Ho, Wo)
    y = np.empty((B, Co, To, Ho, Wo), np.float32)
    lib.ora_conv3d_pad(_p(x), _p(w), _p(b), _p(y), B, Ci, T, H, W, Co, pt, ph, 1)
    _close(y, torch.relu(pre).detach().numpy())
    gz = rs.randn(B, Co, To, Ho, Wo).astype(np.float32)
    pre.backward(torch.from_numpy(gz).double())
    gx = np.empty_like(x)
    lib.ora_conv3d_dgrad(_p(gz), _p(w), _p(gx), B, Ci, T, H, W, Co, pt, ph)
    _close(gx, xd.grad.numpy())
    dw, db = np.empty_like(w), np.empty_like(b)
    lib.ora_conv3d_wgrad(_p(x), _p(gz), _p(dw), _p(db), B, Ci, T, H, W, Co, pt, ph)
    _close(dw, wd.grad.numpy())
    _close(db, bd.grad.numpy())


def test_unpadded_entry_equals_the_padded_one_with_zero_padding(lib):
    rs = np.random.RandomState(3)
    B, Ci, T, H, W, Co = 1, 2, 3, 5, 4, 3
    x = rs.randn(B, Ci, T, H, W).astype(np.float32)
    w = rs.randn(Co, Ci, 3, 3, 3).astype(np.float32)
    b = rs.randn(Co).astype(np.float32)
    y0 = np.empty((B, Co, T - 2, H - 2, W - 2), np.float32)
    y1 = np.empty_like(y0)
    lib.ora_conv3d_relu(_p(x), _p(w), _p(b), _p(y0), B, Ci, T, H, W, Co, 1)
    lib.ora_conv3d_pad(_p(x), _p(w), _p(b), _p(y1), B, Ci, T, H, W, Co, 0, 0, 1)
    assert np.array_equal(y0, y1)


def test_linear_backward(lib):
    rs = np.random.RandomState(4)
    B, I, O = 3, 17, 5
    x = rs.randn(B, I).astype(np.float32)
    w = rs.randn(O, I).astype(np.float32)
    gy = rs.randn(B, O).astype(np.float32)
    xd = torch.from_numpy(x).double().requires_grad_(True)
    wd = torch.from_numpy(w).double().requires_grad_(True)
    bd = torch.zeros(O, dtype=torch.float64, requires_grad=True)
    F.linear(xd, wd, bd).backward(torch.from_numpy(gy).double())
    gx, dw, db = np.empty_like(x), np.empty_like(w), np.empty(O, np.float32)
    lib.ora_linear_bwd(_p(gy), _p(x), _p(w), _p(gx), _p(dw), _p(db), B, ctypes.c_long(I), O)
    _close(gx, xd.grad.numpy())
    _close(dw, wd.grad.numpy())
    _close(db, bd.grad.numpy())


@pytest.mark.parametrize("ties", [False, True])
def test_maxpool3d_matches_aten_including_ties(lib, ties):
    rs = np.random.RandomState(5)
    P, T, H, W = 3, 4, 7, 6
    x = rs.randn(P, T, H, W).astype(np.float32)
    if ties:  # few distinct values: most windows hold several maxima, the arg-max rule decides where the gradient goes
        x = rs.randint(0, 3, size=(P, T, H, W)).astype(np.float32)
    xt = torch.from_numpy(x).reshape(1, P, T, H, W).requires_grad_(True)
    yt, it = F.max_pool3d(xt, 3, stride=(1, 2, 2), padding=1, return_indices=True)
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    assert tuple(yt.shape) == (1, P, T, Ho, Wo)
    y = np.empty((P, T, Ho, Wo), np.float32)
    idx = np.empty((P, T, Ho, Wo), np.int64)
    lib.ora_maxpool3d(_p(x), _p(y), _p(idx), ctypes.c_long(P), T, H, W)
    assert np.array_equal(y, yt.detach().numpy()[0])
    assert np.array_equal(idx, it.numpy()[0])
    gy = rs.randn(P, T, Ho, Wo).astype(np.float32)
    yt.backward(torch.from_numpy(gy).reshape(1, P, T, Ho, Wo))
    gx = np.empty_like(x)
    lib.ora_maxpool3d_bwd(_p(gy), _p(idx), _p(gx), ctypes.c_long(P), T, H, W)
    _close(gx, xt.grad.numpy()[0], tol=1e-6)


def test_adam_two_steps_match_torch(lib):
    """torch.optim.Adam(lr=5e-4) as configure_optimizers builds it (base_model.py:255-257)."""
    rs = np.random.RandomState(6)
    n = 257
    p0 = rs.randn(n).astype(np.float32)
    grads = [rs.randn(n).astype(np.float32) * s for s in (1.0, 1e-3)]
    pt_ = torch.nn.Parameter(torch.from_numpy(p0.copy()))
    opt = torch.optim.Adam([pt_], lr=0.0005)
    p, m, v = p0.copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
    for step, g in enumerate(grads, start=1):
        pt_.grad = torch.from_numpy(g.copy())
        opt.step()
        lib.ora_adam_step(_p(p), _p(g), _p(m), _p(v), ctypes.c_long(n), ctypes.c_float(0.0005), ctypes.c_float(0.9),
                          ctypes.c_float(0.999), ctypes.c_float(1e-8), step)
        want = pt_.detach().numpy()
        assert np.abs(p - want).max() <= 2e-7 * np.abs(want).max() + 1e-9, (step, np.abs(p - want).max())


@pytest.mark.parametrize("name", ["test_yaml_pv", "test_yaml_gsp", "nwp_pv_small", "pv_only_odd", "nwp_only_one_layer"])
def test_torch_free_forward_reproduces_the_reference_golden(lib, name):
    """The whole forward of the path (model.py:107-156) and the returned loss (base_model.py:90-98), composed from the plain-C
    operators (scalar loops, double accumulation) with numpy for the glue (nan_to_num, concat, slices) -- no torch in the
    arithmetic -- against ``y_hat`` / ``nmae`` recorded from the UNMODIFIED reference (tests/golden/, oracle/make_golden.py).
    A second, library-independent pin of the oracle: flatten order, Linear layout, branch order of the concat, target slice."""
    from oracle import conv3d_oracle as O
    from oracle.golden_cases import CASES, golden_batch, golden_state_dict

    case = CASES[name]
    kw, B = case["model"], case["batch"]
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz")))
    om = O.OracleModel(**kw)  # only for the state_dict key order / shapes of golden_state_dict and the derived sizes
    sd = {k: v.numpy() for k, v in golden_state_dict(om).items()}
    batch = golden_batch(name)

    # a1: int16 normalisation (netcdf_dataset.py:96-101)
    sat = np.ascontiguousarray(batch["satellite"]["data"].numpy())
    _, C, T, H, W = sat.shape
    mean, std = (a.copy() for a in O.sat_constants(C))
    x = np.empty(sat.shape, np.float32)
    lib.ora_sat_normalise(_p(sat), _p(x), _p(mean), _p(std), B, C, ctypes.c_long(T * H * W))

    # a3, a4: Conv3d + ReLU stack (model.py:117-120)
    L = kw["number_of_conv3d_layers"]
    for layer in range(L):
        key = "sat_conv0" if layer == 0 else f"conv3d_{layer}"
        w, b = np.ascontiguousarray(sd[key + ".weight"]), np.ascontiguousarray(sd[key + ".bias"])
        Co, Ci = w.shape[0], w.shape[1]
        y = np.empty((B, Co, T - 2, H - 2, W - 2), np.float32)
        lib.ora_conv3d_relu(_p(x), _p(w), _p(b), _p(y), B, Ci, T, H, W, Co, 1)
        x, T, H, W = y, T - 2, H - 2, W - 2

    def linear(inp, key, relu):
        w, b = np.ascontiguousarray(sd[key + ".weight"]), np.ascontiguousarray(sd[key + ".bias"])
        inp = np.ascontiguousarray(inp, np.float32)
        out = np.empty((inp.shape[0], w.shape[0]), np.float32)
        lib.ora_linear(_p(inp), _p(w), _p(b), _p(out), inp.shape[0], ctypes.c_long(w.shape[1]), w.shape[0], int(relu))
        return out

    # a5, a6: NCDHW flatten, fc1, fc2 (model.py:122-126)
    out = linear(linear(x.reshape(B, -1), "fc1", True), "fc2", True)
    var = kw.get("output_variable", "pv_yield")
    if kw["include_pv_yield"]:  # a7 (model.py:130-136)
        hist = np.nan_to_num(batch[var].numpy()[:, : om.history_len_30 + 1], nan=0.0).astype(np.float32)
        out = np.concatenate([out, hist.reshape(B, -1)], axis=1)
    if kw["include_nwp"]:  # a8 (model.py:139-148)
        out = np.concatenate([out, linear(batch["nwp"].numpy().reshape(B, -1), "fc_nwp", True)], axis=1)
    y_hat = linear(linear(out, "fc3", True), "fc4", False).reshape(B, om.forecast_len)  # a9 (model.py:151-154)
    _close(y_hat, g["y_hat"], tol=1e-5)

    # a10: target slice and the returned L1 loss (base_model.py:90-98)
    target = batch["pv" if var == "pv_yield" else "gsp"][var].numpy()[0:B, -om.forecast_len:, 0].astype(np.float32)
    target = np.ascontiguousarray(target)
    lib.ora_l1_loss.restype = ctypes.c_float
    nmae = lib.ora_l1_loss(_p(np.ascontiguousarray(y_hat)), _p(target), ctypes.c_long(y_hat.size))
    assert abs(nmae - float(g["nmae"])) <= 1e-5 * abs(float(g["nmae"]))
