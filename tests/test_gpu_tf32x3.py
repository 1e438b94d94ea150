"""GPU parity tests of the fp32-MODE tensor-core convolutions (3xTF32, `csrc/conv3d_igemm_tf32x3.cu`) through the C ABI.

Checker = torch CPU operators in fp64 on the same seeded inputs (fp32 values, no rounding of the operands: the point of
the 3xTF32 split is fp32-class accuracy).  Tolerances (normalised max error = max|a-b| / max|b|): forward <= 1e-5 (north
star), data gradient <= 1e-5 (tighter than the 1e-4 of the FMA kernels' tests: the split sits at ~1e-6); a single TF32
MMA would be at ~8e-4 (tools/tf32x3_study.py), so these gates also prove that all three terms are there.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import conv3d_oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-5


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


def _case(shape, seed=0):
    B, Ci, T, H, W, Co = shape
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((B, Ci, T, H, W), generator=g)
    w = (torch.rand((Co, Ci, 3, 3, 3), generator=g) * 2 - 1) / np.sqrt(Ci * 27)
    b = torch.randn((Co,), generator=g) * 0.1
    return x, w, b


@pytest.mark.parametrize("shape", [(2, 12, 3, 7, 9), (1, 32, 2, 5, 5), (3, 5, 1, 4, 6)])
def test_blocked_f32_layout_round_trip(ops, dev, shape):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(shape, generator=g).to(dev)
    B, Cc, T, H, W = shape
    xb = ops.to_blocked_f32(x)
    G = ops.blocked4_groups(Cc)
    assert tuple(xb.shape) == (B, G, T, H, W, 4) and G % 2 == 0
    want = torch.zeros((B, G * 4, T, H, W), device=dev)
    want[:, :Cc] = x
    assert torch.equal(xb, want.view(B, G, 4, T, H, W).permute(0, 1, 3, 4, 5, 2).contiguous())
    assert torch.equal(ops.from_blocked_f32(xb, Cc), x)
    xp = ops.to_blocked_f32(x, pad=2)
    assert torch.equal(xp[:, :, 2:-2, 2:-2, 2:-2], xb)
    assert float(xp.abs().sum()) == pytest.approx(float(xb.abs().sum()), rel=1e-6)


@pytest.mark.parametrize("shape", [(2, 12, 3, 9, 9), (1, 11, 2, 8, 10)])
def test_normalise_blocked_f32_bit_exact(ops, dev, shape):
    g = torch.Generator().manual_seed(2)
    x = torch.randint(-5, 1024, shape, generator=g, dtype=torch.int32).to(torch.int16)
    mean, std = O.sat_constants(shape[1])
    want = O.sat_normalise_numpy(x.numpy(), mean, std)
    got = ops.sat_normalise_blocked_f32(x.to(dev), torch.from_numpy(mean).to(dev), torch.from_numpy(std).to(dev))
    back = ops.from_blocked_f32(got, shape[1])
    assert np.array_equal(back.cpu().numpy().view(np.uint32), want.view(np.uint32))
    G = ops.blocked4_groups(shape[1])
    flat = got.permute(0, 1, 5, 2, 3, 4).reshape(shape[0], G * 4, *shape[2:])
    assert float(flat[:, shape[1]:].abs().sum()) == 0.0  # pad channels are zero


FWD_SHAPES = [
    (2, 32, 5, 10, 10, 32),   # the 32 -> 32 layers in miniature
    (1, 12, 4, 16, 16, 32),   # layer 0: 12 channels (one zero group)
    (2, 11, 3, 9, 12, 32),    # the production yaml's 11 channels, rectangular plane
    (1, 32, 3, 11, 13, 16),   # Cout = 16: a single half
    (1, 8, 3, 8, 8, 24),      # Cout = 24: second half partly empty
    (1, 32, 4, 64, 64, 32),   # full-size plane (30 position tiles, halo of 130 positions)
    (1, 32, 21, 6, 6, 32),    # runs longer than the accumulator-block ring
    (40, 4, 3, 5, 5, 8),      # more columns than CTAs per half with tiny planes
]


@pytest.mark.parametrize("shape", FWD_SHAPES)
def test_conv3d_fwd_tf32x3(ops, dev, shape):
    B, Ci, T, H, W, Co = shape
    x, w, b = _case(shape)
    want_pre = F.conv3d(x.double(), w.double(), b.double())
    xb = ops.to_blocked_f32(x.to(dev))
    y_blk, y_nc = ops.conv3d_fwd_tf32x3(xb, w.to(dev), b.to(dev), relu=True, want_blk=True, want_nc=True)
    assert tuple(y_nc.shape) == (B, Co, T - 2, H - 2, W - 2)
    err = nerr(y_nc, F.relu(want_pre))
    print(f"tf32x3 fwd {shape}: normalised error {err:.2e}")
    assert err <= TOL
    assert torch.equal(ops.from_blocked_f32(y_blk, Co), y_nc)  # the two copies hold the same values
    # no ReLU, no bias, padded blocked output only
    y_pad, none = ops.conv3d_fwd_tf32x3(xb, w.to(dev), None, relu=False, out_pad=2, want_blk=True, want_nc=False)
    assert none is None
    want_nb = F.conv3d(x.double(), w.double(), None)
    got = ops.from_blocked_f32(y_pad[:, :, 2:-2, 2:-2, 2:-2].contiguous(), Co)
    assert nerr(got, want_nb) <= TOL
    assert float(y_pad.abs().sum()) == pytest.approx(float(got.abs().sum()), rel=1e-5)  # border stayed zero


@pytest.mark.parametrize("shape", [(2, 32, 5, 10, 10, 32), (1, 32, 4, 20, 22, 32), (1, 16, 3, 9, 9, 32), (2, 32, 3, 8, 8, 12),
                                   (1, 32, 19, 7, 7, 32)])
def test_conv3d_dgrad_tf32x3(ops, dev, shape):
    B, Ci, T, H, W, Co = shape
    x, w, _ = _case(shape, seed=1)
    g = torch.Generator().manual_seed(2)
    gz = torch.randn((B, Co, T - 2, H - 2, W - 2), generator=g)
    mask_src = torch.randn((B, Ci, T, H, W), generator=g)
    xd = x.double().requires_grad_(True)
    F.conv3d(xd, w.double(), None).backward(gz.double())
    want = xd.grad
    gzp = ops.to_blocked_f32(gz.to(dev), pad=2)
    gx_blk, gx_nc = ops.conv3d_dgrad_tf32x3(gzp, w.to(dev), None, want_blk=True, want_nc=True)
    err = nerr(gx_nc, want)
    print(f"tf32x3 dgrad {shape}: normalised error {err:.2e}")
    assert err <= TOL
    assert torch.equal(ops.from_blocked_f32(gx_blk, Ci), gx_nc)
    mb = ops.to_blocked_f32(mask_src.to(dev))
    gx_pad, gx_nc_m = ops.conv3d_dgrad_tf32x3(gzp, w.to(dev), mb, out_pad=2, want_blk=True, want_nc=True)
    want_m = want * (mask_src > 0).double()
    assert nerr(gx_nc_m, want_m) <= TOL
    assert torch.equal(ops.from_blocked_f32(gx_pad[:, :, 2:-2, 2:-2, 2:-2].contiguous(), Ci), gx_nc_m)
    assert float(gx_pad.abs().sum()) == pytest.approx(float(gx_nc_m.abs().sum()), rel=1e-5)


@pytest.mark.parametrize("shape", [(2, 32, 5, 10, 10, 32), (1, 12, 1, 9, 9, 32), (2, 32, 19, 8, 8, 32)])
def test_conv3d_tf32x3_time_padded(ops, dev, shape):
    """padding (1, 0, 0) (the towers of conv3d_sat_nwp): the planes of the padding are skipped in the kernel."""
    B, Ci, T, H, W, Co = shape
    x, w, b = _case(shape, seed=21)
    g = torch.Generator().manual_seed(22)
    gz = torch.randn((B, Co, T, H - 2, W - 2), generator=g)
    xd = x.double().requires_grad_(True)
    pre = F.conv3d(xd, w.double(), b.double(), padding=(1, 0, 0))
    pre.backward(gz.double())
    xb = ops.to_blocked_f32(x.to(dev))
    _, y = ops.conv3d_fwd_tf32x3(xb, w.to(dev), b.to(dev), relu=True, pad_t=1, want_blk=False, want_nc=True)
    assert nerr(y, F.relu(pre.detach())) <= TOL
    gzp = ops.to_blocked_f32(gz.to(dev), pad=2)
    _, gx = ops.conv3d_dgrad_tf32x3(gzp, w.to(dev), None, pad_t=1, want_blk=False, want_nc=True)
    assert tuple(gx.shape) == (B, Ci, T, H, W)
    assert nerr(gx, xd.grad) <= TOL


WGRAD_SHAPES = [
    (2, 32, 5, 10, 10, 32),
    (1, 12, 4, 16, 16, 32),   # layer 0: 12 channels
    (2, 11, 3, 9, 12, 32),    # odd channel count, rectangular plane (Wo = 10: two K steps, the second half padding)
    (1, 32, 3, 11, 13, 16),   # Cout = 16: N = 48
    (1, 8, 3, 8, 8, 24),      # Cout = 24 padded to 32 columns per tap
    (1, 32, 3, 64, 64, 32),   # full-size rows (the over-read of the M = 128 instruction stays inside shared memory)
    (3, 32, 21, 6, 6, 32),    # many short steps: several flush windows per CTA
    (37, 4, 3, 5, 5, 8),      # more (sample, plane, row) steps than CTAs
]


@pytest.mark.parametrize("shape", WGRAD_SHAPES)
@pytest.mark.parametrize("gz_pad", [0, 2])
def test_conv3d_wgrad_bf16x3(ops, dev, shape, gz_pad):
    B, Ci, T, H, W, Co = shape
    assert ops.wgrad_bf16x3_supported(Ci, Co, H, W)
    x, w, b = _case(shape, seed=4)
    g = torch.Generator().manual_seed(5)
    gz = torch.randn((B, Co, T - 2, H - 2, W - 2), generator=g)
    wd = w.double().requires_grad_(True)
    bd = b.double().requires_grad_(True)
    F.conv3d(x.double(), wd, bd).backward(gz.double())
    xb = ops.to_blocked_f32(x.to(dev))
    gzb = ops.to_blocked_f32(gz.to(dev), pad=gz_pad)
    dw, db = ops.conv3d_wgrad_bf16x3(xb, gzb, Ci, Co, gz_pad=gz_pad)
    e_w, e_b = nerr(dw, wd.grad), nerr(db, bd.grad)
    print(f"tf32x3 wgrad {shape} pad {gz_pad}: dw {e_w:.2e} db {e_b:.2e}")
    assert e_w <= TOL and e_b <= TOL
    dw2, db2 = ops.conv3d_wgrad_bf16x3(xb, gzb, Ci, Co, gz_pad=gz_pad)  # deterministic: same bits on a second run
    assert torch.equal(dw, dw2) and torch.equal(db, db2)


def _amax(ops, dev, *tensors):
    out = torch.zeros((len(tensors),), device=dev)
    for i, t in enumerate(tensors):
        ops.absmax_f32(t, out[i:i + 1])
    return tuple(out[i:i + 1] for i in range(len(tensors)))


@pytest.mark.parametrize("shape", WGRAD_SHAPES)
@pytest.mark.parametrize("scales", [(1.0, 1.0), (3.0e4, 2.0e-9), (1.0e-12, 7.0e6)])
def test_conv3d_wgrad_f16x2(ops, dev, shape, scales):
    """Two-way fp16 split of the scaled operands (three products): the fp32 parity bound whatever the magnitudes of the
    activations and of the gradient (the kernel scales by the tensors' largest magnitudes into fp16's range)."""
    B, Ci, T, H, W, Co = shape
    x, w, b = _case(shape, seed=4)
    g = torch.Generator().manual_seed(5)
    gz = torch.randn((B, Co, T - 2, H - 2, W - 2), generator=g) * scales[1]
    x = x * scales[0]
    wd = w.double().requires_grad_(True)
    bd = b.double().requires_grad_(True)
    F.conv3d(x.double(), wd, bd).backward(gz.double())
    xb = ops.to_blocked_f32(x.to(dev))
    gzb = ops.to_blocked_f32(gz.to(dev), pad=2)
    amax = _amax(ops, dev, xb, gzb)
    assert float(amax[0]) == float(x.abs().max()) and float(amax[1]) == float(gz.abs().max())
    dw, db = ops.conv3d_wgrad_bf16x3(xb, gzb, Ci, Co, gz_pad=2, amax=amax)
    e_w, e_b = nerr(dw, wd.grad), nerr(db, bd.grad)
    print(f"f16x2 wgrad {shape} scales {scales}: dw {e_w:.2e} db {e_b:.2e}")
    assert e_w <= TOL and e_b <= TOL
    dw2, db2 = ops.conv3d_wgrad_bf16x3(xb, gzb, Ci, Co, gz_pad=2, amax=amax)
    assert torch.equal(dw, dw2) and torch.equal(db, db2)


@pytest.mark.parametrize("shape", [(2, 32, 5, 10, 10, 32), (3, 32, 21, 6, 6, 32), (37, 4, 3, 5, 5, 8), (2, 32, 7, 62, 62, 32)])
@pytest.mark.parametrize("two_way", [True, False])
def test_conv3d_wgrad_dynamic_chunks(ops, dev, shape, two_way):
    """The weight gradient with its steps claimed in chunks from an atomic counter (what data parallelism switches on):
    same result as the static split up to the order of the fp32 sums, more chunks than CTAs, and CTAs that get none."""
    from predict_pv_yield_b200 import lib

    B, Ci, T, H, W, Co = shape
    x, w, b = _case(shape, seed=31)
    g = torch.Generator().manual_seed(32)
    gz = torch.randn((B, Co, T - 2, H - 2, W - 2), generator=g)
    wd = w.double().requires_grad_(True)
    bd = b.double().requires_grad_(True)
    F.conv3d(x.double(), wd, bd).backward(gz.double())
    xb, gzb = ops.to_blocked_f32(x.to(dev)), ops.to_blocked_f32(gz.to(dev), pad=2)
    amax = _amax(ops, dev, xb, gzb) if two_way else None
    L = lib.load()
    old = L.pvb200_set_dynamic_tiles(1)
    try:
        for _ in range(3):  # the claim order differs from run to run
            dw, db = ops.conv3d_wgrad_bf16x3(xb, gzb, Ci, Co, gz_pad=2, amax=amax)
            assert nerr(dw, wd.grad) <= TOL and nerr(db, bd.grad) <= TOL
    finally:
        L.pvb200_set_dynamic_tiles(old)


def test_conv3d_wgrad_full_size_accuracy(ops, dev):
    """The weight gradient at the FULL conv1 shape of the BASELINE model (32 -> 32 channels, 17x62x62 input, batch 32): a CTA
    then runs ~220 steps = 14 flush windows of the toward-zero accumulators, which is where their length shows (16 steps:
    3.4e-6 -- torch's own fp32 convolution is 3.4e-6 from fp64 here; 32 steps: 7e-6; 8 steps: 1.8e-6).  Checker: torch fp64
    on the GPU, in slices of four samples."""
    torch.backends.cudnn.allow_tf32 = False
    B, Ci, T, S, Co = 32, 32, 17, 62, 32
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.relu(torch.randn(B, Ci, T, S, S, device=dev, generator=g))
    gz = torch.randn(B, Co, T - 2, S - 2, S - 2, device=dev, generator=g)
    wd = torch.zeros(Co, Ci, 3, 3, 3, device=dev, dtype=torch.float64, requires_grad=True)
    bd = torch.zeros(Co, device=dev, dtype=torch.float64, requires_grad=True)
    for b0 in range(0, B, 4):
        F.conv3d(x[b0:b0 + 4].double(), wd, bd).backward(gz[b0:b0 + 4].double())
    am = torch.zeros(2, device=dev)
    xb, gzb = ops.to_blocked_f32(x, amax=am[0:1]), ops.to_blocked_f32(gz, pad=2, amax=am[1:2])
    for name, kw in (("f16x2", dict(amax=(am[0:1], am[1:2]))), ("bf16x3", {})):
        dw, db = ops.conv3d_wgrad_bf16x3(xb, gzb, Ci, Co, gz_pad=2, **kw)
        e_w, e_b = nerr(dw, wd.grad), nerr(db, bd.grad)
        print(f"full-size wgrad {name}: dw {e_w:.2e} db {e_b:.2e}")
        assert e_w <= TOL and e_b <= TOL, name


def test_conv3d_wgrad_f16x2_wide_dynamic_range(ops, dev):
    """Operands whose magnitudes span 2^40 inside one tensor: the small values lose RELATIVE precision in the fp16 split,
    the gradient (dominated by the large ones) stays within the fp32 bound; an all-zero gradient gives exact zeros."""
    shape = (2, 32, 5, 12, 12, 32)
    B, Ci, T, H, W, Co = shape
    g = torch.Generator().manual_seed(9)
    x = torch.randn((B, Ci, T, H, W), generator=g) * torch.exp2(torch.randint(-30, 10, (B, Ci, T, H, W), generator=g).float())
    gz = torch.randn((B, Co, T - 2, H - 2, W - 2), generator=g) * torch.exp2(torch.randint(-40, 0, (B, Co, T - 2, H - 2, W - 2), generator=g).float())
    wd = torch.zeros((Co, Ci, 3, 3, 3), dtype=torch.float64, requires_grad=True)
    F.conv3d(x.double(), wd).backward(gz.double())
    xb, gzb = ops.to_blocked_f32(x.to(dev)), ops.to_blocked_f32(gz.to(dev), pad=2)
    dw, db = ops.conv3d_wgrad_bf16x3(xb, gzb, Ci, Co, gz_pad=2, amax=_amax(ops, dev, xb, gzb))
    assert nerr(dw, wd.grad) <= TOL and nerr(db, gz.double().sum((0, 2, 3, 4))) <= TOL
    gzb.zero_()
    dw, db = ops.conv3d_wgrad_bf16x3(xb, gzb, Ci, Co, gz_pad=2, amax=_amax(ops, dev, xb, gzb))
    assert not bool(dw.any()) and not bool(db.any())


F16_CONV_SHAPES = [
    (2, 32, 5, 10, 10, 32),   # two 16-channel steps
    (1, 12, 4, 16, 16, 32),   # Cin = 12 padded to one 16-channel step
    (3, 32, 4, 64, 64, 32),   # full-size planes, an odd number of (sample, q-tile) columns
    (1, 28, 6, 9, 12, 24),    # Cout = 24 (the pair kernel's 32 columns per tap, 8 of them zero)
]


@pytest.mark.parametrize("shape", F16_CONV_SHAPES)
@pytest.mark.parametrize("scales", [(1.0, 1.0), (2.0e4, 3.0e-7), (1.0e-10, 5.0e3)])
def test_conv3d_fwd_dgrad_f16x2(ops, dev, shape, scales):
    """Forward and data gradient through the two-way fp16 split (CTA-pair kernel): the fp32 parity bound whatever the
    magnitudes of the input and of the weights."""
    B, Ci, T, H, W, Co = shape
    x, w, b = _case(shape, seed=21)
    x, w = x * scales[0], w * scales[1]
    b = b * scales[0] * scales[1]
    ref = F.conv3d(x.double(), w.double(), b.double())
    xb = ops.to_blocked_f32(x.to(dev))
    (ax,) = _amax(ops, dev, xb)
    am_out = torch.zeros((1,), device=dev)
    y_blk, y_nc = ops.conv3d_fwd_tf32x3(xb, w.to(dev), b.to(dev), relu=False, want_blk=True, want_nc=True, amax_in=ax, amax=am_out)
    e_f = nerr(y_nc, ref)
    assert float(am_out) == float(y_nc.abs().max())
    g = torch.Generator().manual_seed(22)
    gz = torch.randn(ref.shape, generator=g) * scales[0]
    xd = x.double().requires_grad_(True)
    F.conv3d(xd, w.double()).backward(gz.double())
    mask = torch.rand(x.shape, generator=g) - 0.3  # ReLU mask source
    gzb = ops.to_blocked_f32(gz.to(dev), pad=2)
    (ag,) = _amax(ops, dev, gzb)
    _, gx = ops.conv3d_dgrad_tf32x3(gzb, w.to(dev), ops.to_blocked_f32(mask.to(dev)), want_blk=False, want_nc=True, amax_in=ag)
    e_d = nerr(gx, xd.grad * (mask > 0).double())
    print(f"f16x2 conv {shape} scales {scales}: fwd {e_f:.2e} dgrad {e_d:.2e}")
    assert e_f <= TOL and e_d <= TOL


def test_amax_out_of_the_producing_kernels(ops, dev):
    """Every kernel that writes a blocked fp32 tensor of the encoder can report the largest magnitude it wrote (what the
    two-way fp16 split scales by): normalise, layout change, convolution forward and data gradient."""
    g = torch.Generator().manual_seed(12)
    am = torch.zeros((4,), device=dev)
    sat = torch.randint(0, 1024, (2, 12, 5, 10, 10), generator=g, dtype=torch.int32).to(torch.int16).to(dev)
    mean, std = torch.rand(12, generator=g).to(dev) * 500, (torch.rand(12, generator=g) * 100 + 50).to(dev)
    xb = ops.sat_normalise_blocked_f32(sat, mean, std, amax=am[0:1])
    assert float(am[0]) == float(xb.abs().max())
    x, w, b = _case((2, 12, 5, 10, 10, 32), seed=13)
    xb = ops.to_blocked_f32(x.to(dev), amax=am[1:2])
    assert float(am[1]) == float(x.abs().max())
    y_blk, _ = ops.conv3d_fwd_tf32x3(xb, w.to(dev), b.to(dev), relu=True, want_blk=True, amax=am[2:3])
    assert float(am[2]) == float(y_blk.abs().max()) > 0
    gz = torch.randn((2, 32, 3, 8, 8), generator=g).to(dev)
    gx_blk, _ = ops.conv3d_dgrad_tf32x3(ops.to_blocked_f32(gz, pad=2), w.to(dev), xb, out_pad=2, want_blk=True, amax=am[3:4])
    assert float(am[3]) == float(gx_blk.abs().max()) > 0


def test_conv3d_wgrad_bf16x3_time_padded(ops, dev):
    shape = (2, 32, 5, 10, 10, 32)
    B, Ci, T, H, W, Co = shape
    x, w, b = _case(shape, seed=6)
    g = torch.Generator().manual_seed(7)
    gz = torch.randn((B, Co, T, H - 2, W - 2), generator=g)
    wd = w.double().requires_grad_(True)
    bd = b.double().requires_grad_(True)
    F.conv3d(x.double(), wd, bd, padding=(1, 0, 0)).backward(gz.double())
    dw, db = ops.conv3d_wgrad_bf16x3(ops.to_blocked_f32(x.to(dev)), ops.to_blocked_f32(gz.to(dev), pad=2), Ci, Co, gz_pad=2, pad_t=1)
    assert nerr(dw, wd.grad) <= TOL and nerr(db, bd.grad) <= TOL


def test_conv3d_wgrad_bf16x3_rejects_wide_planes(ops, dev):
    """128-wide rows (the deep variant's first layers) do not fit the staging buffers: the host says so and the encoder
    keeps those layers on the fp32 FMA weight gradient."""
    from predict_pv_yield_b200 import lib

    assert not ops.wgrad_bf16x3_supported(32, 32, 126, 126)
    assert ops.wgrad_bf16x3_supported(32, 32, 64, 64)
    L = lib.load()
    rc = L.pvb200_conv3d_wgrad_bf16x3(1, 1, 0, 1, 1, 1, 0, 1, 32, 3, 126, 126, 32, 0, 0)
    assert rc != 0 and b"not supported" in L.pvb200_last_error()


def test_conv3d_tf32x3_four_layer_chain(ops, dev):
    """Four chained 32 -> 32 layers with ReLU (the depth of the BASELINE model): the toward-zero accumulator of the
    tensor core must not let the error grow past the 1e-5 bound (per-plane accumulator blocks + corrections first)."""
    g = torch.Generator().manual_seed(5)
    x = F.relu(torch.randn((1, 32, 11, 24, 24), generator=g))
    ws = [(torch.rand((32, 32, 3, 3, 3), generator=g) * 2 - 1) / np.sqrt(32 * 27) * 2.0 for _ in range(4)]
    bs = [torch.randn((32,), generator=g) * 0.05 for _ in range(4)]
    ref = x.double()
    xb = ops.to_blocked_f32(x.to(dev))
    y = None
    for l in range(4):
        ref = F.relu(F.conv3d(ref, ws[l].double(), bs[l].double()))
        xb, y = ops.conv3d_fwd_tf32x3(xb, ws[l].to(dev), bs[l].to(dev), relu=True, want_blk=True, want_nc=True)
    err = nerr(y, ref)
    print(f"tf32x3 four-layer chain: normalised error {err:.2e}")
    assert err <= TOL


def test_fp32_model_tensor_core_and_fma_paths_agree(dev):
    """The fp32 model through the 3xTF32 encoder and through the direct FMA kernels: same loss / forecast / gradients
    within the fp32 parity bound, and the tensor-core path really launched its kernels."""
    from oracle.golden_cases import CASES, golden_batch, golden_state_dict
    from predict_pv_yield_b200.models.conv3d.model import Model

    case = CASES["nwp_pv_small"]
    batch = O.batch_to(golden_batch("nwp_pv_small"), dev)
    outs = {}
    for tc in (True, False):
        m = Model(**case["model"]).to(dev)
        m.fp32_tensor_cores = tc
        m.batch_size = case["batch"]
        m.load_state_dict(golden_state_dict(m))
        loss = m.training_step(batch, 0)
        loss.backward()
        outs[tc] = (float(loss.detach()), {k: p.grad.clone() for k, p in m.named_parameters()})
    assert abs(outs[True][0] - outs[False][0]) <= 1e-5 * abs(outs[False][0])
    for k in outs[True][1]:
        assert nerr(outs[True][1][k], outs[False][1][k]) <= (2e-2 if "conv" in k else 2e-3), k
