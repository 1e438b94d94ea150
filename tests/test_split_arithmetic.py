"""CPU statement of the split-precision arithmetic the fp32 mode runs on the tensor cores, in numpy:
the identities the kernels rely on hold for every input class the GPU tests feed them, and the error of a K = 864
contraction (one output of a 32 -> 32 channel 3x3x3 convolution) stays two orders below the 1e-5 parity bound.

* three-way bf16 split (``conv3d_wgrad_bf16x3``, ``fc1_bf16x3``): an fp32 value IS the sum of three bf16 values;
* 3xTF32 (``conv3d_igemm_tf32x3``): ``x_lo = x - trunc(x)`` is exact in fp32 and fits TF32 after truncation to 2^-22;
* two-way fp16 split (``*_f16x2``): operands scaled by ``2^(14 - floor(log2 max|v|))`` so that the largest magnitude lands
  in [2^14, 2^15); two fp16 pieces keep 22 significand bits of everything within 2^18 of the tensor's maximum.

The contraction tests round the fp32 accumulator to nearest once per MMA; the hardware rounds it TOWARD ZERO (a half-ulp bias
per accumulating MMA), which the kernels bound by keeping chains short (36 main-term MMAs per accumulator block, corrections
first; tools/tf32x3_study.py emulates that variant).

These are properties of the number formats, not of the device: the kernels themselves are gated against torch fp64 in
tests/test_gpu_tf32x3.py (forward, data and weight gradients, operands scaled by 1e-12 .. 3e4, magnitudes spanning 2^40).
"""
import numpy as np
import pytest


def _bf16_rne(v: np.ndarray) -> np.ndarray:
    """fp32 -> nearest bf16 (ties to even), returned as fp32 (what ``cvt.rn.bf16x2.f32`` produces)."""
    u = v.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def _tf32_trunc(v: np.ndarray) -> np.ndarray:
    """What ``tcgen05.mma.kind::tf32`` reads of a 32-bit operand: the low 13 mantissa bits are dropped (probe,
    profiles/tf32_probe_r02.txt)."""
    return (v.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def _operands(kind: str, rs: np.random.RandomState, n: int) -> np.ndarray:
    if kind == "normal":
        return rs.randn(n).astype(np.float32)
    if kind == "relu":  # activations: half of them exact zeros
        return np.maximum(rs.randn(n), 0).astype(np.float32)
    if kind == "tiny":  # the scale of a gradient late in training
        return (rs.randn(n) * 1e-12).astype(np.float32)
    if kind == "large":
        return (rs.randn(n) * 3e4).astype(np.float32)
    if kind == "wide":  # magnitudes spanning 2^40 inside one tensor
        return (rs.randn(n) * np.exp2(rs.uniform(-40, 0, n))).astype(np.float32)
    raise ValueError(kind)


KINDS = ["normal", "relu", "tiny", "large", "wide"]


@pytest.mark.parametrize("kind", KINDS)
def test_an_fp32_value_is_the_sum_of_three_bf16_values(kind):
    v = _operands(kind, np.random.RandomState(1), 1 << 16)
    b0 = _bf16_rne(v)
    r1 = (v - b0).astype(np.float32)  # exact: the residual of a round-to-nearest fits fp32
    assert np.array_equal(r1.astype(np.float64), v.astype(np.float64) - b0.astype(np.float64))
    b1 = _bf16_rne(r1)
    r2 = (r1 - b1).astype(np.float32)
    assert np.array_equal(r2.astype(np.float64), r1.astype(np.float64) - b1.astype(np.float64))
    b2 = _bf16_rne(r2)
    # 8 + 8 + 8 significand bits cover fp32's 24: nothing is left, unless the third piece falls below bf16's (= fp32's)
    # subnormal range -- only for |v| < 2^-110, which "tiny" (1e-12 = 2^-40) does not reach
    assert np.array_equal(b0.astype(np.float64) + b1.astype(np.float64) + b2.astype(np.float64), v.astype(np.float64))


@pytest.mark.parametrize("kind", KINDS)
def test_tf32_residual_is_exact_and_small(kind):
    v = _operands(kind, np.random.RandomState(2), 1 << 16)
    hi = _tf32_trunc(v)
    lo = (v - hi).astype(np.float32)
    assert np.array_equal(lo.astype(np.float64), v.astype(np.float64) - hi.astype(np.float64))  # exact in fp32
    nz = v != 0
    assert np.all(np.abs(lo[nz]) < np.abs(v[nz]) * 2.0 ** -10)
    # the hardware truncates lo too: what is dropped is 2^-21 of the value at most (2^-10 * 2^-11 ... the kernels' bound 2^-22
    # is for the dropped lo*lo PRODUCT)
    lolo = (lo - _tf32_trunc(lo)).astype(np.float64)
    assert np.all(np.abs(lolo[nz]) <= np.abs(v[nz].astype(np.float64)) * 2.0 ** -20)


def _scale_exponent(v: np.ndarray) -> float:
    """``s = 2^(14 - floor(log2 max|v|))``: the largest magnitude lands in [2^14, 2^15) (fp16's largest finite value is 65504)."""
    amax = float(np.abs(v).max())
    return float(np.exp2(14 - np.floor(np.log2(amax))))


def _f16_pieces(v: np.ndarray):
    s = np.float32(_scale_exponent(v))
    sv = (v * s).astype(np.float32)  # a power of two: exact (no fp32 under/overflow for the tensors of this model)
    h0 = sv.astype(np.float16)
    h1 = (sv - h0.astype(np.float32)).astype(np.float32).astype(np.float16)
    return s, h0, h1


@pytest.mark.parametrize("kind", KINDS)
def test_fp16_pieces_keep_22_bits_relative_to_the_tensor_scale(kind):
    v = _operands(kind, np.random.RandomState(3), 1 << 16)
    s, h0, h1 = _f16_pieces(v)
    assert np.all(np.isfinite(h0.astype(np.float32))) and float(np.abs(h0.astype(np.float32)).max()) < 2.0 ** 15
    back = (h0.astype(np.float64) + h1.astype(np.float64)) / float(s)
    err = np.abs(back - v.astype(np.float64))
    amax = float(np.abs(v).max())
    # absolute error <= 2^-22 of the value, or 2^-39 of the tensor's maximum where the second piece is an fp16 subnormal
    bound = np.maximum(np.abs(v.astype(np.float64)) * 2.0 ** -22, amax * 2.0 ** -39)
    assert np.all(err <= bound)


def _contract_f16x2(x: np.ndarray, w: np.ndarray) -> np.ndarray:
    """[N, K] x [K] -> [N]: x.w ~ (x0.w1 + x1.w0 + x0.w0) / (sx * sw), corrections first, fp32 accumulation per MMA (K = 16)."""
    sx, x0, x1 = _f16_pieces(x)
    sw, w0, w1 = _f16_pieces(w)
    acc = np.zeros(x.shape[0], np.float32)
    x0f, x1f, w0f, w1f = (a.astype(np.float32) for a in (x0, x1, w0, w1))
    for a, b in ((x0f, w1f), (x1f, w0f), (x0f, w0f)):
        for k in range(0, x.shape[1], 16):
            # products of two fp16 values are exact in fp32; the 16 products of one MMA are summed before the accumulator
            # is rounded (float64 partial sum here, one fp32 rounding per MMA)
            acc = (acc.astype(np.float64) + (a[:, k:k + 16].astype(np.float64) * b[k:k + 16].astype(np.float64)).sum(1)).astype(np.float32)
    return (acc.astype(np.float64) / (float(sx) * float(sw))).astype(np.float32)


def _contract_tf32x3(x: np.ndarray, w: np.ndarray) -> np.ndarray:
    xh, wh = _tf32_trunc(x), _tf32_trunc(w)
    xl, wl = _tf32_trunc((x - xh).astype(np.float32)), _tf32_trunc((w - wh).astype(np.float32))
    acc = np.zeros(x.shape[0], np.float32)
    for a, b in ((xh, wl), (xl, wh), (xh, wh)):
        for k in range(0, x.shape[1], 8):
            acc = (acc.astype(np.float64) + (a[:, k:k + 8].astype(np.float64) * b[k:k + 8].astype(np.float64)).sum(1)).astype(np.float32)
    return acc


@pytest.mark.parametrize("xkind,wkind", [("relu", "normal"), ("normal", "normal"), ("tiny", "normal"), ("large", "tiny"), ("wide", "wide")])
@pytest.mark.parametrize("scheme", ["f16x2", "tf32x3"])
def test_contraction_error_of_the_split_schemes(scheme, xkind, wkind):
    """One output of a 32 -> 32 channel 3x3x3 layer is a K = 864 dot product.  Normalised max error (max|a - b| / max|b|, the
    definition of every parity gate) of the emulated scheme against fp64: two orders below 1e-5, for every operand class."""
    rs = np.random.RandomState(4)
    N, K = 4096, 864
    x = _operands(xkind, rs, N * K).reshape(N, K)
    w = _operands(wkind, rs, K)
    want = x.astype(np.float64) @ w.astype(np.float64)
    got = (_contract_f16x2 if scheme == "f16x2" else _contract_tf32x3)(x, w).astype(np.float64)
    err = float(np.abs(got - want).max() / np.abs(want).max())
    plain = float(np.abs((x @ w).astype(np.float64) - want).max() / np.abs(want).max())  # numpy's own fp32 dot product
    print(f"{scheme} {xkind} x {wkind}: {err:.2e} (fp32 dot product: {plain:.2e})")
    assert err <= 2e-6


def test_a_single_reduced_precision_product_is_not_enough():
    """Why the split exists: one TF32 (or bf16) MMA per product misses the 1e-5 bound by two orders (8e-4)."""
    rs = np.random.RandomState(5)
    x, w = _operands("relu", rs, 2048 * 864).reshape(2048, 864), _operands("normal", rs, 864)
    want = x.astype(np.float64) @ w.astype(np.float64)
    one = _tf32_trunc(x).astype(np.float64) @ _tf32_trunc(w).astype(np.float64)
    assert float(np.abs(one - want).max() / np.abs(want).max()) > 1e-4
