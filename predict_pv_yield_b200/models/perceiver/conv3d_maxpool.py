"""``Conv3dMaxPool`` -- drop-in for the block of the same name in
``predict_pv_yield/models/perceiver/perceiver_conv3d_nwp_sat.py:42-57``: a Conv3d with
"same" padding followed by ``MaxPool3d(3, stride=(1, 2, 2), padding=(1, 1, 1))``, the front-end the Perceiver hybrid puts
in front of its satellite and NWP inputs (``:94-95,151,163``).  Same constructor, same sub-module names (``sat_conv3d``
holds the parameters, ``sat_maxpool`` has none), so ``state_dict`` interoperates; the arithmetic runs in
``libpvb200.so`` (fp32 direct convolution with padding (1,1,1), max-pool forward / deterministic backward).  The
Perceiver itself (external ``perceiver_pytorch``) is out of scope.
"""
from __future__ import annotations

import torch
from torch import nn

from ... import ops


class Conv3dMaxPool(nn.Module):
    def __init__(self, out_channels: int, in_channels: int):
        super().__init__()
        # parameter container: convolution, padded so the output is the same size
        self.sat_conv3d = nn.Conv3d(in_channels=in_channels, out_channels=out_channels, kernel_size=(3, 3, 3), padding=(1, 1, 1))
        # max pool that keeps the time sequence the same length (no parameters; kept for the module tree)
        self.sat_maxpool = nn.MaxPool3d(3, stride=(1, 2, 2), padding=(1, 1, 1))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("predict_pv_yield_b200 Conv3dMaxPool is CUDA (sm_100a) only: there is no CPU fallback")
        return ops.Conv3dMaxPoolFn.apply(x.float(), self.sat_conv3d.weight, self.sat_conv3d.bias)
