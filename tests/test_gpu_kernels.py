"""GPU parity tests, kernel by kernel, through the C ABI (predict_pv_yield_b200.ops -> libpvb200.so).

Checker = torch CPU operators in fp64 (the library the reference's arithmetic lives in), on the same
seeded inputs.  Tolerances (normalised max error = max|a-b| / max|b|):
  * normalisation: bit-exact (fp32), RNE-exact (bf16)
  * fp32 forward results: <= 1e-5  (north star);  fp32 gradients: <= 1e-4 (torch's own fp32 wgrad
    sits at 5e-6..2.5e-5 of max|g| against fp64)
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import conv3d_oracle as O

pytestmark = pytest.mark.gpu

FWD_TOL = 1e-5
GRAD_TOL = 1e-4


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def ops(dev):
    from predict_pv_yield_b200 import lib, ops as _ops

    lib.load()
    return _ops


def nerr(a, b):
    return O.normalised_max_err(a, b)


# ---------------------------------------------------------------------------------------- normalise
def test_normalise_exhaustive_bit_exact(ops, dev):
    x = torch.arange(-32768, 32768, dtype=torch.int32).to(torch.int16)
    cube = x.view(1, 1, 1, 256, 256).expand(2, 12, 1, 256, 256).contiguous()
    mean, std = O.sat_constants(12)
    want = O.sat_normalise_numpy(cube.numpy(), mean, std)
    got = ops.sat_normalise(cube.to(dev), torch.from_numpy(mean).to(dev), torch.from_numpy(std).to(dev))
    assert np.array_equal(got.cpu().numpy().view(np.uint32), want.view(np.uint32))
    got16 = ops.sat_normalise(cube.to(dev), torch.from_numpy(mean).to(dev), torch.from_numpy(std).to(dev),
                              out_dtype=torch.bfloat16)
    want16 = torch.from_numpy(want).to(torch.bfloat16)  # torch's cast is RNE
    assert torch.equal(got16.cpu().view(torch.int16), want16.view(torch.int16))


@pytest.mark.parametrize("shape", [(1, 3, 1, 5, 7), (2, 11, 3, 9, 9), (1, 12, 19, 64, 64)])
def test_normalise_ragged_shapes(ops, dev, shape):
    g = torch.Generator().manual_seed(1)
    x = torch.randint(-5, 1024, shape, generator=g, dtype=torch.int32).to(torch.int16)
    mean, std = O.sat_constants(shape[1])
    want = O.sat_normalise_numpy(x.numpy(), mean, std)
    got = ops.sat_normalise(x.to(dev), torch.from_numpy(mean).to(dev), torch.from_numpy(std).to(dev))
    assert np.array_equal(got.cpu().numpy().view(np.uint32), want.view(np.uint32))


# ---------------------------------------------------------------------------------------- conv3d
CONV_SHAPES = [
    # B, Ci, T, H, W, Co
    (2, 12, 5, 16, 16, 32),
    (1, 11, 4, 9, 13, 32),
    (2, 32, 5, 14, 14, 32),
    (1, 32, 3, 62, 62, 32),   # full-width layer-1 plane: several tiles per plane, Wps = 64
    (1, 32, 4, 30, 27, 32),   # odd width: pitch padding
    (1, 5, 3, 8, 8, 20),      # channel counts that are not multiples of the tiles
    (1, 40, 3, 8, 8, 48),
]


def _conv_case(shape, seed=0):
    B, Ci, T, H, W, Co = shape
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((B, Ci, T, H, W), generator=g)
    w = torch.randn((Co, Ci, 3, 3, 3), generator=g) / np.sqrt(Ci * 27)
    b = torch.randn((Co,), generator=g) * 0.1
    return x, w, b


@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv3d_fwd(ops, dev, shape):
    x, w, b = _conv_case(shape)
    want = F.relu(F.conv3d(x.double(), w.double(), b.double()))
    got = ops.conv3d_fwd(x.to(dev), w.to(dev), b.to(dev), relu=True)
    assert got.shape == want.shape
    e = nerr(got, want)
    assert e <= FWD_TOL, e
    want_lin = F.conv3d(x.double(), w.double(), None)
    got_lin = ops.conv3d_fwd(x.to(dev), w.to(dev), None, relu=False)
    assert nerr(got_lin, want_lin) <= FWD_TOL


@pytest.mark.parametrize("shape", [(2, 12, 5, 16, 16, 32), (1, 11, 4, 10, 12, 32)])
def test_conv3d_fwd_fused_int16_normalise(ops, dev, shape):
    B, Ci, T, H, W, Co = shape
    g = torch.Generator().manual_seed(3)
    xi = torch.randint(-1, 1024, (B, Ci, T, H, W), generator=g, dtype=torch.int32).to(torch.int16)
    _, w, b = _conv_case(shape)
    mean, std = O.sat_constants(Ci)
    xn = O.sat_normalise(xi, torch.from_numpy(mean), torch.from_numpy(std))
    want = F.relu(F.conv3d(xn.double(), w.double(), b.double()))
    got = ops.conv3d_fwd(xi.to(dev), w.to(dev), b.to(dev), relu=True, mean=torch.from_numpy(mean).to(dev),
                         std=torch.from_numpy(std).to(dev))
    assert nerr(got, want) <= FWD_TOL
    # and identical to running the standalone normalise kernel first (same fp32 values enter the FMAs)
    xn_dev = ops.sat_normalise(xi.to(dev), torch.from_numpy(mean).to(dev), torch.from_numpy(std).to(dev))
    got2 = ops.conv3d_fwd(xn_dev, w.to(dev), b.to(dev), relu=True)
    assert torch.equal(got, got2)


@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv3d_dgrad(ops, dev, shape):
    B, Ci, T, H, W, Co = shape
    x, w, _ = _conv_case(shape, seed=1)
    g = torch.Generator().manual_seed(2)
    gz = torch.randn((B, Co, T - 2, H - 2, W - 2), generator=g)
    mask_src = torch.randn((B, Ci, T, H, W), generator=g)
    xd = x.double().requires_grad_(True)
    F.conv3d(xd, w.double(), None).backward(gz.double())
    want = xd.grad
    got = ops.conv3d_dgrad(gz.to(dev), w.to(dev), None, x.shape)
    assert nerr(got, want) <= GRAD_TOL
    got_m = ops.conv3d_dgrad(gz.to(dev), w.to(dev), mask_src.to(dev), x.shape)
    want_m = want * (mask_src > 0).double()
    assert nerr(got_m, want_m) <= GRAD_TOL


@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv3d_wgrad(ops, dev, shape):
    B, Ci, T, H, W, Co = shape
    x, w, b = _conv_case(shape, seed=4)
    g = torch.Generator().manual_seed(5)
    gz = torch.randn((B, Co, T - 2, H - 2, W - 2), generator=g)
    wd = w.double().requires_grad_(True)
    bd = b.double().requires_grad_(True)
    F.conv3d(x.double(), wd, bd).backward(gz.double())
    dw, db = ops.conv3d_wgrad(x.to(dev), gz.to(dev))
    assert nerr(dw, wd.grad) <= GRAD_TOL
    assert nerr(db, bd.grad) <= GRAD_TOL
    # deterministic: same bits on a second run
    dw2, db2 = ops.conv3d_wgrad(x.to(dev), gz.to(dev))
    assert torch.equal(dw, dw2) and torch.equal(db, db2)


def test_conv3d_wgrad_fused_int16(ops, dev):
    shape = (2, 12, 5, 16, 16, 32)
    B, Ci, T, H, W, Co = shape
    g = torch.Generator().manual_seed(7)
    xi = torch.randint(-1, 1024, (B, Ci, T, H, W), generator=g, dtype=torch.int32).to(torch.int16)
    gz = torch.randn((B, Co, T - 2, H - 2, W - 2), generator=g)
    mean, std = O.sat_constants(Ci)
    xn = O.sat_normalise(xi, torch.from_numpy(mean), torch.from_numpy(std))
    wd = torch.zeros((Co, Ci, 3, 3, 3), dtype=torch.float64, requires_grad=True)
    F.conv3d(xn.double(), wd, None).backward(gz.double())
    dw, _ = ops.conv3d_wgrad(xi.to(dev), gz.to(dev), torch.from_numpy(mean).to(dev), torch.from_numpy(std).to(dev))
    assert nerr(dw, wd.grad) <= GRAD_TOL


def test_conv3d_rejects_cpu_tensors(ops):
    x, w, b = _conv_case((1, 4, 3, 8, 8, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.conv3d_fwd(x, w, b)


# ---------------------------------------------------------------------------------------- head
HEAD_CASES = [
    # B, K1, F1, F2, F3, FO, pv(nt, ns) or None, nwp
    (3, 22528, 128, 128, 64, 12, (2, 128), True),
    (2, 34816, 16, 16, 16, 12, None, False),
    (5, 1001, 24, 40, 8, 2, (1, 32), False),   # K1 not a multiple of 4: scalar path; odd sizes
    (37, 4096, 128, 128, 64, 12, None, True),  # more than one batch tile
    (2, 300, 130, 16, 16, 4, None, False),     # F1 > one feature tile
    (48, 8200, 100, 64, 32, 12, None, False),  # tensor-core fc1: K1 not a multiple of the 64 / 128 tiles, F1 < 128
    (64, 12548, 128, 128, 64, 12, (2, 128), False),  # two batch chunks (B > 48) in all three tensor-core kernels; the weight gradient's second chunk adds through the TMA engine
    (17, 64 * 1300, 128, 128, 64, 12, None, False),  # several k tiles per CTA (accumulator folds, both TMEM windows)
]


def _linear(g, o, i):
    return torch.randn((o, i), generator=g) / np.sqrt(i), torch.randn((o,), generator=g) * 0.1


@pytest.mark.parametrize("case", HEAD_CASES)
def test_head_fwd_bwd(ops, dev, case):
    B, K1, F1, F2, F3, FO, pv, use_nwp = case
    g = torch.Generator().manual_seed(11)
    feats = F.relu(torch.randn((B, K1), generator=g))
    npv = pv[0] * pv[1] if pv else 0
    w1, b1 = _linear(g, F1, K1)
    w2, b2 = _linear(g, F2, F1)
    ncat = F2 + npv + (128 if use_nwp else 0)
    w3, b3 = _linear(g, F3, ncat)
    w4, b4 = _linear(g, FO, F3)
    wn, bn = _linear(g, 128, 760) if use_nwp else (None, None)
    nwp = torch.randn((B, 760), generator=g) if use_nwp else None
    pv_full = None
    if pv:
        pv_full = torch.rand((B, pv[0] + 3, pv[1]), generator=g)
        pv_full[:, 0, ::7] = float("nan")
    gout = torch.randn((B, FO), generator=g)

    # fp64 truth with torch
    P = [t.double().requires_grad_(True) if t is not None else None for t in (w1, b1, w2, b2, wn, bn, w3, b3, w4, b4)]
    fd = feats.double().requires_grad_(True)
    h = F.relu(F.linear(fd, P[0], P[1]))
    h = F.relu(F.linear(h, P[2], P[3]))
    if pv:
        hist = pv_full[:, : pv[0]].nan_to_num(nan=0.0).double().reshape(B, -1)
        h = torch.cat((h, hist), dim=1)
    if use_nwp:
        h = torch.cat((h, F.relu(F.linear(nwp.double(), P[4], P[5]))), dim=1)
    h = F.relu(F.linear(h, P[6], P[7]))
    out = F.linear(h, P[8], P[9])
    out.backward(gout.double())

    c = lambda t: None if t is None else t.to(dev)  # noqa: E731
    params = [c(t).requires_grad_(True) if t is not None else None for t in (w1, b1, w2, b2, wn, bn, w3, b3, w4, b4)]
    fdev = c(feats).requires_grad_(True)
    pv_view = c(pv_full)[:, : pv[0]] if pv else None
    got = ops.HeadFn.apply(fdev, pv_view, c(nwp), *params)
    assert nerr(got, out) <= FWD_TOL
    got.backward(c(gout))
    # the head returns the gradient w.r.t. the features with the ReLU mask (feats > 0) fused in
    want_gx = fd.grad * (feats > 0).double()
    assert nerr(fdev.grad, want_gx) <= GRAD_TOL
    names = ["w1", "b1", "w2", "b2", "wn", "bn", "w3", "b3", "w4", "b4"]
    for nm, pg, pw in zip(names, params, P):
        if pg is None:
            continue
        assert nerr(pg.grad, pw.grad) <= GRAD_TOL, nm


def test_head_fc1_tensor_core_equals_fma_full_size(ops, dev):
    """fc1 of the BASELINE model (32 x 1 103 872 -> 128) on the tensor cores (three-way bf16 split, fc1_bf16x3.cu) against
    the FMA-pipe kernels on the same inputs: forward, data gradient (ReLU mask fused) and weight gradient."""
    import ctypes

    from predict_pv_yield_b200 import lib
    L = lib.load()
    L.pvb200_debug_set_fc1x3.argtypes = [ctypes.c_int]
    L.pvb200_debug_set_fc1x3.restype = None
    B, K1, F1 = 32, 32 * 11 * 56 * 56, 128
    g = torch.Generator(device=dev).manual_seed(3)
    feats = F.relu(torch.randn((B, K1), device=dev, generator=g))
    gc = torch.Generator().manual_seed(4)
    small = [t.to(dev) for pair in (_linear(gc, 128, F1), _linear(gc, 64, 128), _linear(gc, 12, 64)) for t in pair]
    w1 = torch.randn((F1, K1), device=dev, generator=g) / np.sqrt(K1)
    b1 = torch.randn((F1,), device=dev, generator=g) * 0.1
    gout = torch.randn((B, 12), device=dev, generator=g)
    res = {}
    try:
        for mode in (0, 1):
            L.pvb200_debug_set_fc1x3(mode)
            params = [t.clone().requires_grad_(True) for t in (w1, b1, small[0], small[1])] + [None, None] + \
                     [t.clone().requires_grad_(True) for t in small[2:]]
            fdev = feats.clone().requires_grad_(True)
            out = ops.HeadFn.apply(fdev, None, None, *params)
            out.backward(gout)
            res[mode] = (out.detach(), fdev.grad, params[0].grad, params[1].grad)
    finally:
        L.pvb200_debug_set_fc1x3(1)
    for nm, a, b in zip(("out", "g_feats", "dw1", "db1"), res[1], res[0]):
        assert nerr(a, b) <= 2e-6, nm


# ---------------------------------------------------------------------------------------- loss / adam
def test_step_loss(ops, dev):
    g = torch.Generator().manual_seed(5)
    B, FO = 32, 12
    y_hat = torch.randn((B, FO), generator=g)
    yfull = torch.rand((B, 19, 128), generator=g)
    y = yfull[:, -FO:, 0]
    y_hat[0, 0] = y[0, 0]  # exercise sign(0) = 0
    w = O.weighted_loss_weights(FO)
    yh = y_hat.double().requires_grad_(True)
    d = yh - y.double()
    want = torch.stack([d.abs().mean(), (d ** 2).mean(), (w.double() * d ** 2).mean(), (w.double() * d.abs()).mean()])
    (3.0 * want[0]).backward()
    yd = y_hat.to(dev).requires_grad_(True)
    got = ops.StepLossFn.apply(yd, yfull.to(dev)[:, -FO:, 0], w.to(dev))
    assert nerr(got, want) <= 1e-6
    (3.0 * got[0]).backward()
    assert nerr(yd.grad, yh.grad) <= 1e-6


def test_fused_adam_matches_torch(ops, dev):
    from predict_pv_yield_b200.optim import FusedAdam

    g = torch.Generator().manual_seed(9)
    shapes = [(32, 12, 3, 3, 3), (32,), (16, 34816), (7,), (4099,)]
    ref_p = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in shapes]
    our_p = [torch.nn.Parameter(p.detach().clone().to(dev)) for p in ref_p]
    ref_opt = torch.optim.Adam(ref_p, lr=5e-4)
    our_opt = FusedAdam(our_p, lr=5e-4)
    for _ in range(3):
        for rp, op in zip(ref_p, our_p):
            gr = torch.randn(rp.shape, generator=g)
            rp.grad = gr.clone()
            op.grad = gr.to(dev)
        ref_opt.step()
        our_opt.step()
    for rp, op in zip(ref_p, our_p):
        assert float((op.detach().cpu() - rp.detach()).abs().max()) <= 5e-7
    assert our_opt.state[our_p[0]]["step"] == 3
