"""SURVEY.md section 8f rank 4 -- Conv3dMaxPool (predict_pv_yield/models/perceiver/perceiver_conv3d_nwp_sat.py:42-57).

CPU: a torch restatement of the block is pinned against outputs of the UNMODIFIED reference class
(tests/golden/conv3d_maxpool_*.npz, oracle/make_golden.py).  GPU: the CUDA path (fp32 direct convolution with padding
(1,1,1), max-pool forward / gather backward through the C ABI) against those goldens, 1e-5 (forward) / 1e-4 (gradients).
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import conv3d_oracle as O
from oracle.golden_cases import MAXPOOL_CASES, golden_state_dict, maxpool_inputs


def _oracle_block(Co, Ci):
    """perceiver_conv3d_nwp_sat.py:44-57 restated (test infrastructure)."""
    m = torch.nn.Module()
    m.sat_conv3d = torch.nn.Conv3d(Ci, Co, kernel_size=(3, 3, 3), padding=(1, 1, 1))
    m.forward = lambda x: F.max_pool3d(m.sat_conv3d(x), 3, stride=(1, 2, 2), padding=(1, 1, 1))
    return m


@pytest.mark.parametrize("name", list(MAXPOOL_CASES))
def test_oracle_matches_reference(golden_dir, name):
    g = dict(np.load(os.path.join(golden_dir, f"{name}.npz")))
    B, Ci, T, H, W, Co = MAXPOOL_CASES[name]
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        m = _oracle_block(Co, Ci)
        m.load_state_dict(golden_state_dict(m))
        x, gy = maxpool_inputs(name)
        x.requires_grad_(True)
        y = m.forward(x)
        y.backward(gy)
    finally:
        torch.set_num_threads(n)
    assert np.array_equal(y.detach().numpy(), g["y"])
    assert np.array_equal(x.grad.numpy(), g["gx"])
    assert O.normalised_max_err(m.sat_conv3d.weight.grad, torch.from_numpy(g["dw"])) <= 1e-6


def test_mirror_state_dict_and_cpu_refusal():
    from predict_pv_yield_b200.models.perceiver.conv3d_maxpool import Conv3dMaxPool

    m = Conv3dMaxPool(out_channels=16, in_channels=11)
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [("sat_conv3d.weight", (16, 11, 3, 3, 3)), ("sat_conv3d.bias", (16,))]
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 11, 3, 8, 8))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(MAXPOOL_CASES))
def test_cuda_block_matches_reference_golden(golden_dir, name):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from predict_pv_yield_b200.models.perceiver.conv3d_maxpool import Conv3dMaxPool

    dev = torch.device("cuda:0")
    g = dict(np.load(os.path.join(golden_dir, f"{name}.npz")))
    B, Ci, T, H, W, Co = MAXPOOL_CASES[name]
    m = Conv3dMaxPool(out_channels=Co, in_channels=Ci).to(dev)
    m.load_state_dict(golden_state_dict(m))
    x, gy = maxpool_inputs(name)
    x = x.to(dev).requires_grad_(True)
    y = m(x)
    y.backward(gy.to(dev))
    assert tuple(y.shape) == g["y"].shape
    assert O.normalised_max_err(y.detach(), torch.from_numpy(g["y"])) <= 1e-5
    assert O.normalised_max_err(x.grad, torch.from_numpy(g["gx"])) <= 1e-4
    assert O.normalised_max_err(m.sat_conv3d.weight.grad, torch.from_numpy(g["dw"])) <= 1e-4
    assert O.normalised_max_err(m.sat_conv3d.bias.grad, torch.from_numpy(g["db"])) <= 1e-4


@pytest.mark.gpu
def test_maxpool_kernels_ties_and_full_size():
    """Ties resolve to the first maximum like torch (identical arg-max routing), and the pool runs at the Perceiver
    hybrid's full size (32 x 32 channels x 31 x 64 x 64) with the gradient summing to the upstream gradient."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from predict_pv_yield_b200 import lib

    L = lib.load()
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(5)
    x = torch.randint(0, 3, (2, 3, 4, 9, 10), generator=g).float()  # many ties
    gy = torch.randn((2, 3, 4, 5, 5), generator=g)
    xd = x.clone().requires_grad_(True)
    F.max_pool3d(xd, 3, stride=(1, 2, 2), padding=(1, 1, 1)).backward(gy)
    y = torch.empty((2, 3, 4, 5, 5), device=dev)
    arg = torch.empty((2, 3, 4, 5, 5), dtype=torch.int32, device=dev)
    gx = torch.empty_like(x, device=dev)
    lib.check(L.pvb200_maxpool3d_fwd_f32(x.to(dev).data_ptr(), y.data_ptr(), arg.data_ptr(), 6, 4, 9, 10, st))
    lib.check(L.pvb200_maxpool3d_bwd_f32(gy.to(dev).data_ptr(), arg.data_ptr(), gx.data_ptr(), 6, 4, 9, 10, st))
    assert torch.equal(y.cpu(), F.max_pool3d(x, 3, stride=(1, 2, 2), padding=(1, 1, 1)))
    assert torch.allclose(gx.cpu(), xd.grad, atol=1e-6)
    xb = torch.randn((32 * 32, 31, 64, 64), device=dev)
    yb = torch.empty((32 * 32, 31, 32, 32), device=dev)
    ab = torch.empty((32 * 32, 31, 32, 32), dtype=torch.int32, device=dev)
    gb = torch.empty_like(xb)
    lib.check(L.pvb200_maxpool3d_fwd_f32(xb.data_ptr(), yb.data_ptr(), ab.data_ptr(), 32 * 32, 31, 64, 64, st))
    ones = torch.ones_like(yb)
    lib.check(L.pvb200_maxpool3d_bwd_f32(ones.data_ptr(), ab.data_ptr(), gb.data_ptr(), 32 * 32, 31, 64, 64, st))
    assert float(gb.sum()) == float(ones.sum())
    assert torch.equal(yb[:4], F.max_pool3d(xb[:4].unsqueeze(0), 3, stride=(1, 2, 2), padding=(1, 1, 1))[0])
