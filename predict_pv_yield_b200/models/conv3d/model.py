"""Conv3d PV-yield model -- drop-in for ``predict_pv_yield.models.conv3d.model.Model``.

Mirrors ``predict_pv_yield/models/conv3d/model.py:14-156``: same constructor arguments and defaults
(``:18-32``), same attributes (``cnn_output_size`` ``:74-78``), same sub-module names and shapes
(``sat_conv0``, ``conv3d_{i}``, ``fc1``..``fc4``, ``fc_nwp``; ``:80-103``) so ``state_dict`` /
``load_state_dict`` interoperate with reference checkpoints, same ``forward(x)`` contract (dict or
BatchML in, fp32 ``[B, forecast_len]`` out; ``:107-156``).

What differs is where the arithmetic runs.  The ``nn.Conv3d`` / ``nn.Linear`` members only HOLD the
parameters (torch default init, reference key names); their ``forward`` is never called.  The step is
executed by hand-written sm_100a kernels in ``libpvb200.so``:

  int16 satellite cube --(fused normalise, netcdf_dataset.py:96-101)--> conv0+ReLU -> conv_l+ReLU ...
     -> flatten (NCDHW order, model.py:122) -> fc1 (split-K weight-streaming GEMM) -> fused tail
     [fc2, PV-history nan_to_num+concat, fc_nwp, concat, fc3, fc4]

There is no CPU fallback: tensors must live on a CUDA device and ``libpvb200.so`` must be built.
"""
from __future__ import annotations

import logging

import numpy as np
import torch
from torch import nn

from ..base_model import BaseModel
from ...batch import as_batch
from ... import ops

logging.basicConfig()
_LOG = logging.getLogger("predict_pv_yield_b200")

# predict_pv_yield/netcdf_dataset.py:16-32 (channel order HRV, IR_016, IR_039, IR_087, IR_097, IR_108,
# IR_120, IR_134, VIS006, VIS008, WV_062, WV_073)
SAT_MEAN = np.array(
    [93.23458, 131.71373, 843.7779, 736.6148, 771.1189, 589.66034,
     862.29816, 927.69586, 90.70885, 107.58985, 618.4583, 532.47394], dtype=np.float32)
SAT_STD = np.array(
    [115.34247, 139.92636, 36.99538, 57.366386, 30.346825,
     149.68007, 51.70631, 35.872967, 115.77212, 120.997154,
     98.57828, 99.76469], dtype=np.float32)


class Model(BaseModel):

    name = "conv3d"

    # inference (no_grad) batches larger than this are streamed through the kernels in micro-batches, so the
    # activation workspace stays fixed when forecasting all GB PV systems at once (BASELINE config 4: B = 512..8192)
    inference_micro_batch = 256
    # fp32 mode: run Conv3d forward / data gradient as 3xTF32 implicit GEMMs on the tensor cores (fp32-class accuracy);
    # False keeps every convolution on the direct fp32 FMA kernels
    fp32_tensor_cores = True

    def __init__(
        self,
        include_pv_yield: bool = True,
        include_nwp: bool = True,
        forecast_minutes: int = 30,
        history_minutes: int = 60,
        number_of_conv3d_layers: int = 4,
        conv3d_channels: int = 32,
        image_size_pixels: int = 64,
        number_sat_channels: int = 12,
        fc1_output_features: int = 128,
        fc2_output_features: int = 128,
        fc3_output_features: int = 64,
        output_variable: str = "pv_yield",
        precision: str = "fp32",
    ):
        """
        3d conv model, that takes in different data streams (same arguments as the reference model, plus
        ``precision``: "fp32" = fp32-accurate kernels, 3xTF32 tensor-core / FMA-pipe (parity <= 1e-5), "bf16" = bf16 tensor-core
        (tcgen05) convolutions with fp32 accumulation and fp32 master weights (parity <= 2e-2)).

        include_pv_yield: include pv yield history
        include_nwp: include nwp data
        forecast_minutes / history_minutes: forecast horizon / history length in minutes
        number_of_conv3d_layers, conv3d_channels: depth and width of the Conv3d encoder
        image_size_pixels: the input satellite image size
        number_sat_channels: number of satellite channels
        fc{1,2,3}_output_features: widths of the fully connected layers
        output_variable: 'pv_yield' or 'gsp_yield'
        """
        # stored BEFORE BaseModel.__init__, which reads them (model.py:57-68)
        self.include_pv_yield = include_pv_yield
        self.include_nwp = include_nwp
        self.number_of_conv3d_layers = number_of_conv3d_layers
        self.number_of_nwp_features = 10 * 19 * 2 * 2
        self.fc1_output_features = fc1_output_features
        self.fc2_output_features = fc2_output_features
        self.fc3_output_features = fc3_output_features
        self.forecast_minutes = forecast_minutes
        self.history_minutes = history_minutes
        self.output_variable = output_variable
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        self.precision = precision

        super().__init__()

        self.number_sat_channels = number_sat_channels
        self.cnn_output_size = (
            conv3d_channels
            * ((image_size_pixels - 2 * self.number_of_conv3d_layers) ** 2)
            * (self.forecast_len_5 + self.history_len_5 + 1 - 2 * self.number_of_conv3d_layers)
        )
        if self.cnn_output_size <= 0:
            raise ValueError("image / sequence too small for this many 3x3x3 layers")

        # parameter containers, constructed in the reference's order (same init stream under a seed)
        self.sat_conv0 = nn.Conv3d(
            in_channels=number_sat_channels, out_channels=conv3d_channels, kernel_size=(3, 3, 3), padding=0
        )
        for i in range(0, self.number_of_conv3d_layers - 1):
            layer = nn.Conv3d(
                in_channels=conv3d_channels, out_channels=conv3d_channels, kernel_size=(3, 3, 3), padding=0
            )
            setattr(self, f"conv3d_{i + 1}", layer)

        self.fc1 = nn.Linear(in_features=self.cnn_output_size, out_features=self.fc1_output_features)
        self.fc2 = nn.Linear(in_features=self.fc1_output_features, out_features=self.fc2_output_features)

        fc3_in_features = self.fc2_output_features
        if include_pv_yield:
            fc3_in_features += self.number_of_samples_per_batch * (self.history_len_30 + 1)
        if include_nwp:
            self.fc_nwp = nn.Linear(in_features=self.number_of_nwp_features, out_features=128)
            fc3_in_features += 128

        self.fc3 = nn.Linear(in_features=fc3_in_features, out_features=self.fc3_output_features)
        self.fc4 = nn.Linear(in_features=self.fc3_output_features, out_features=self.forecast_len)

        # normalisation constants for int16 input; non-persistent => state_dict keys equal the reference's.
        # Fewer than 12 channels: the LAST n constants (HRV, index 0, is what the 11-channel production
        # config drops); override the buffers for any other channel selection.
        n = number_sat_channels
        if n <= 12:
            mean, std = SAT_MEAN[12 - n:], SAT_STD[12 - n:]
        else:
            mean, std = np.zeros(n, np.float32), np.ones(n, np.float32)
        # bf16 mode: tensor-core shadow of fc1.weight; FusedAdam finds it through the parameter and keeps it fresh
        self._fc1_shadow = ops.Fc1Shadow()
        if precision == "bf16":
            self.fc1.weight._pvb_shadow = self._fc1_shadow
        else:
            # fp32 mode: every reader of fc1.weight inside this module goes through wait_ready (forward) -- FusedAdam may run
            # its update on a side stream under the next step's convolutions; state_dict() waits through the hook below
            self.fc1.weight._pvb_overlap_ok = True
        self.register_state_dict_pre_hook(lambda *a, **k: self._wait_params())
        self.register_buffer("sat_mean", torch.from_numpy(mean.copy()), persistent=False)
        self.register_buffer("sat_std", torch.from_numpy(std.copy()), persistent=False)

    def _wait_params(self):
        from ...dp import wait_ready

        wait_ready(self.fc1.weight)

    def invalidate_shadow(self) -> None:
        """bf16 mode: force the tensor-core shadow of ``fc1.weight`` to be rebuilt by the next forward.  The shadow follows
        ``load_state_dict``, in-place edits of the parameter and optimiser steps by itself (``ops.Fc1Shadow``); writes
        through ``fc1.weight.data`` (EMA / SWA weight swaps, custom initialisers) do not bump the parameter's version
        counter, so call this after them.  Under a row-sharded optimiser (``dp.GradientExchange(shard_large=True)``) gather the master rows
        first (``gather_master_weights()``): the rebuild reads the fp32 master."""
        self._fc1_shadow.key = None

    def _conv_params(self):
        wb = [self.sat_conv0.weight, self.sat_conv0.bias]
        for i in range(0, self.number_of_conv3d_layers - 1):
            layer = getattr(self, f"conv3d_{i + 1}")
            wb += [layer.weight, layer.bias]
        return wb

    def forward(self, x):
        x = as_batch(x)
        n = x.satellite.data.shape[0]
        if not torch.is_grad_enabled() and n > self.inference_micro_batch:
            outs = []
            for i in range(0, n, self.inference_micro_batch):
                outs.append(self._forward(x, slice(i, min(i + self.inference_micro_batch, n))))
            return torch.cat(outs, dim=0)
        return self._forward(x, None)

    def _forward(self, x, rows):
        """One pass of the step's forward on the batch rows ``rows`` (None = all)."""
        sel = (lambda t: t) if rows is None else (lambda t: t[rows])

        # ******************* Satellite imagery *************************
        # Shape: batch_size, channel, seq_length, height, width
        sat_data = sel(x.satellite.data)
        if not sat_data.is_cuda:
            raise RuntimeError(
                "predict_pv_yield_b200.Model is CUDA (sm_100a) only: move the batch to the GPU "
                "(there is no CPU fallback)"
            )
        if sat_data.dtype == torch.int16:
            # raw SEVIRI counts: normalisation (netcdf_dataset.py:96-101) is fused into conv0's loads
            mean, std = self.sat_mean, self.sat_std
        else:
            sat_data = sat_data.float()  # model.py:113 (already-normalised input)
            mean = std = None
        sat_data = sat_data.contiguous()
        batch_size = sat_data.shape[0]

        # Conv3d + ReLU stack, flattened in NCDHW order (model.py:117-122)
        # tensor-core fc1 needs fc1_output_features <= 128, channels in whole (even) groups of 8 and a batch that fits
        # its shared-memory tiles: 256 for the forward alone, 128 when the weight gradient will run too
        max_b = 128 if torch.is_grad_enabled() else 256
        bf16_head = (self.precision == "bf16" and batch_size <= max_b and self.fc1_output_features <= 128
                     and self.sat_conv0.out_channels % 16 == 0)
        link = {"shadow": self._fc1_shadow} if bf16_head else None
        if self.precision == "bf16":
            out = ops.EncoderBf16Fn.apply(link, sat_data, mean, std, *self._conv_params())
            n_feat = out[0].numel() if bf16_head else out.shape[1]
        else:
            wb = self._conv_params()
            on_tc = self.fp32_tensor_cores and all(ops.tf32x3_supported(w.shape[1], w.shape[0]) for w in wb[0::2])
            # fp32 mode: forward / data gradient as 3xTF32 implicit GEMMs on the tensor cores when the channel counts
            # fit (<= 32), else the direct fp32 kernels on the FMA pipe -- both hold the 1e-5 parity bound
            out = (ops.EncoderTf32Fn if on_tc else ops.EncoderFn).apply(sat_data, mean, std, *wb)
            n_feat = out.shape[1]
        if n_feat != self.cnn_output_size:
            raise RuntimeError(
                f"satellite cube {tuple(sat_data.shape)} gives {n_feat} conv features, "
                f"model expects cnn_output_size={self.cnn_output_size}"
            )

        # add pv yield history (model.py:130-136); nan_to_num + concat are fused into the head kernel
        pv_yield_history = None
        if self.include_pv_yield:
            pv_yield_history = sel(x[self.output_variable])[:, : self.history_len_30 + 1]
            if pv_yield_history.dtype != torch.float32:
                pv_yield_history = pv_yield_history.float()
            if pv_yield_history.stride(-1) != 1:
                pv_yield_history = pv_yield_history.contiguous()

        # NWP data (model.py:139-148)
        nwp_data = None
        wn = bn = None
        if self.include_nwp:
            nwp_data = sel(x["nwp"]).float().flatten(start_dim=1).contiguous()
            wn, bn = self.fc_nwp.weight, self.fc_nwp.bias

        head_args = (
            out, pv_yield_history, nwp_data,
            self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, wn, bn,
            self.fc3.weight, self.fc3.bias, self.fc4.weight, self.fc4.bias,
        )
        if getattr(self.fc1.weight, "_pvb_ready", None) is not None:
            # fc1.weight is still being written on another stream -- its Adam step on the optimiser's side stream
            # (FusedAdam.overlap_large) or, data parallel in fp32 mode, the all-gather of the rows other ranks updated
            # (dp.ShardSpec.all_gather_rows): the head is the first reader
            from ...dp import wait_ready

            wait_ready(self.fc1.weight)
        out = ops.HeadBf16Fn.apply(link, *head_args) if bf16_head else ops.HeadFn.apply(*head_args)
        out = out.reshape(batch_size, self.forecast_len)
        return out
