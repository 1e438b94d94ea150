"""GPU parity tests of the bf16 tensor-core (tcgen05) kernels through the C ABI.

Checker: torch CPU fp64 on the SAME bf16-rounded inputs and weights, so the only differences are the fp32
accumulation order in TMEM and the final bf16 rounding of the stored result: |err| <= 2^-8 * max|y| is the
gate (north star: <= 2e-2 for bf16 loss / forecast).
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import conv3d_oracle as O

pytestmark = pytest.mark.gpu
BF16_TOL = 2.0 ** -8


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


def r16(t):
    return t.to(torch.bfloat16).to(torch.float32)


SHAPES = [
    # B, Ci, T, H, W, Co
    (2, 32, 5, 14, 14, 32),
    (1, 12, 4, 16, 16, 32),
    (1, 32, 3, 62, 62, 32),
    (2, 32, 7, 30, 27, 32),
    (1, 11, 3, 9, 13, 32),
    (3, 32, 19, 20, 20, 32),   # many t tiles: exercises the plane ring across segments and units
    (1, 16, 3, 10, 10, 16),
]


@pytest.mark.parametrize("shape", [(1, 12, 3, 5, 7), (2, 32, 4, 6, 6), (1, 5, 2, 3, 3)])
def test_blocked_layout_roundtrip(ops, dev, shape):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(shape, generator=g)
    for pad in (0, 2):
        xb = ops.to_blocked_bf16(x.to(dev), pad=pad)
        B, C, T, H, W = shape
        Cg = ops.blocked_groups(C)
        assert xb.shape == (B, Cg, T + 2 * pad, H + 2 * pad, W + 2 * pad, 8)
        inner = xb[:, :, pad:pad + T, pad:pad + H, pad:pad + W].contiguous()
        back = ops.from_blocked_bf16(inner, C)
        assert torch.equal(back.cpu(), r16(x))
        # torch view of the same layout
        ref = torch.zeros((B, Cg * 8, T, H, W))
        ref[:, :C] = r16(x)
        ref = ref.view(B, Cg, 8, T, H, W).permute(0, 1, 3, 4, 5, 2)
        assert torch.equal(inner.float().cpu(), ref)
        if pad:
            assert float(xb.float().abs().sum()) == pytest.approx(float(inner.float().abs().sum()))


@pytest.mark.parametrize("shape", SHAPES)
def test_conv3d_fwd_bf16(ops, dev, shape):
    B, Ci, T, H, W, Co = shape
    g = torch.Generator().manual_seed(1)
    x = r16(torch.randn((B, Ci, T, H, W), generator=g))
    w = torch.randn((Co, Ci, 3, 3, 3), generator=g) / np.sqrt(Ci * 27)
    b = torch.randn((Co,), generator=g) * 0.1
    want = F.relu(F.conv3d(x.double(), r16(w).double(), b.double()))
    xb = ops.to_blocked_bf16(x.to(dev))
    yb = ops.conv3d_fwd_bf16(xb, w.to(dev), b.to(dev), relu=True)
    got = ops.from_blocked_bf16(yb, Co)
    assert got.shape == want.shape
    e = O.normalised_max_err(got, want)
    assert e <= BF16_TOL, e
    # written into a padded tensor: same interior, zero border
    ybp = ops.conv3d_fwd_bf16(xb, w.to(dev), b.to(dev), relu=True, out_pad=2)
    inner = ybp[:, :, 2:-2, 2:-2, 2:-2].contiguous()
    assert torch.equal(inner, yb)
    assert float(ybp.float().abs().sum()) == pytest.approx(float(yb.float().abs().sum()), rel=1e-6)


@pytest.mark.parametrize("shape", [s for s in SHAPES if s[1] % 16 == 0 and s[5] <= 32])
def test_conv3d_dgrad_bf16(ops, dev, shape):
    B, Ci, T, H, W, Co = shape
    g = torch.Generator().manual_seed(2)
    w = torch.randn((Co, Ci, 3, 3, 3), generator=g) / np.sqrt(Ci * 27)
    gz = r16(torch.randn((B, Co, T - 2, H - 2, W - 2), generator=g))
    mask_src = r16(torch.randn((B, Ci, T, H, W), generator=g))
    xd = torch.zeros((B, Ci, T, H, W), dtype=torch.float64, requires_grad=True)
    F.conv3d(xd, r16(w).double(), None).backward(gz.double())
    want = xd.grad
    gzp = ops.to_blocked_bf16(gz.to(dev), pad=2)
    got = ops.from_blocked_bf16(ops.conv3d_dgrad_bf16(gzp, w.to(dev), None), Ci)
    assert O.normalised_max_err(got, want) <= BF16_TOL
    mb = ops.to_blocked_bf16(mask_src.to(dev))
    got_m = ops.from_blocked_bf16(ops.conv3d_dgrad_bf16(gzp, w.to(dev), mb), Ci)
    assert O.normalised_max_err(got_m, want * (mask_src > 0).double()) <= BF16_TOL
