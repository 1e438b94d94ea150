"""Two-tower Conv3d model (satellite + NWP) -- drop-in for ``predict_pv_yield.models.conv3d.model_sat_nwp.Model``
(the production ``experiment=conv3d_sat_nwp`` config and the Optuna target).

Mirrors ``predict_pv_yield/models/conv3d/model_sat_nwp.py:14-268``: same constructor arguments and defaults
(``:18-37``), same derived sizes (``cnn_output_size`` ``:85-93``, ``nwp_cnn_output_size`` ``:95-99``), same sub-module
names / shapes / construction order (``sat_conv{i}``, ``fc1``, ``fc2``, ``nwp_conv{i}``, ``nwp_fc1``, ``nwp_fc2``,
``pv_system_id_embedding``, ``pv_fc1``, ``fc3``, ``fc4``; ``:101-172``) so ``state_dict`` interoperates with reference
checkpoints, same ``forward(x)`` contract (``:174-268``).  The ``nn.Conv3d`` / ``nn.Linear`` / ``nn.Embedding`` members only
HOLD the parameters; the arithmetic runs in ``libpvb200.so`` (fp32 mode):

  satellite cube -> [Conv3d(pad (1,0,0)) + ReLU] x L (``ops.TowerFn``) -> fc1 (weight-streaming GEMM) -> fc2
  NWP cube       -> [Conv3d(pad (1,0,0)) + ReLU] x L                   -> nwp_fc1                     -> nwp_fc2
  PV / GSP history -> nan_to_num + flatten;  PV history (5 min) -> pv_fc1;  system id -> embedding
  torch.cat (tiny [B, <= 600] tensors, no FLOPs) -> fc3 -> fc4

``precision="bf16"`` runs both towers and their fc1 / nwp_fc1 on the tensor cores (tcgen05 implicit-GEMM convolutions
with the time padding expressed as skipped zero planes, weight-streaming fc1 over a bf16 shadow of the fp32 master
weights, parity <= 2e-2); the small layers stay fp32.  There is no CPU fallback.
"""
from __future__ import annotations

import logging

import numpy as np
import torch
from torch import nn

from ..base_model import BaseModel
from ...batch import as_batch
from ... import ops
from .model import SAT_MEAN, SAT_STD

logging.basicConfig()
_LOG = logging.getLogger("predict_pv_yield_b200")


class Model(BaseModel):

    name = "conv3d_sat_nwp"

    def __init__(
        self,
        include_pv_or_gsp_yield_history: bool = True,
        include_nwp: bool = True,
        forecast_minutes: int = 30,
        history_minutes: int = 60,
        number_of_conv3d_layers: int = 4,
        conv3d_channels: int = 32,
        image_size_pixels: int = 64,
        nwp_image_size_pixels: int = 64,
        number_sat_channels: int = 12,
        number_nwp_channels: int = 10,
        fc1_output_features: int = 128,
        fc2_output_features: int = 128,
        fc3_output_features: int = 64,
        output_variable: str = "pv_yield",
        embedding_dem: int = 16,
        include_pv_yield_history: int = True,
        include_future_satellite: int = True,
        precision: str = "fp32",
    ):
        """Same arguments as the reference model (model_sat_nwp.py:18-72), plus ``precision`` ("fp32" | "bf16")."""
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        self.precision = precision
        self.include_pv_or_gsp_yield_history = include_pv_or_gsp_yield_history
        self.include_nwp = include_nwp
        self.number_of_conv3d_layers = number_of_conv3d_layers
        self.number_of_nwp_features = 128
        self.fc1_output_features = fc1_output_features
        self.fc2_output_features = fc2_output_features
        self.fc3_output_features = fc3_output_features
        self.forecast_minutes = forecast_minutes
        self.history_minutes = history_minutes
        self.output_variable = output_variable
        self.number_nwp_channels = number_nwp_channels
        self.embedding_dem = embedding_dem
        self.include_pv_yield_history = include_pv_yield_history
        self.include_future_satellite = include_future_satellite

        super().__init__()

        # model_sat_nwp.py:85-99 (the time-padded convolutions keep the sequence length)
        if include_future_satellite:
            cnn_output_size_time = self.forecast_len_5 + self.history_len_5 + 1
        else:
            cnn_output_size_time = self.history_len_5 + 1
        self.cnn_output_size = (
            conv3d_channels * ((image_size_pixels - 2 * self.number_of_conv3d_layers) ** 2) * cnn_output_size_time
        )
        self.nwp_cnn_output_size = (
            conv3d_channels
            * ((nwp_image_size_pixels - 2 * self.number_of_conv3d_layers) ** 2)
            * (self.forecast_len_60 + self.history_len_60 + 1)
        )
        if self.cnn_output_size <= 0 or (include_nwp and self.nwp_cnn_output_size <= 0):
            raise ValueError("image too small for this many 3x3x3 layers")

        # parameter containers, constructed in the reference's order (same init stream under a seed)
        self.sat_conv0 = nn.Conv3d(
            in_channels=number_sat_channels, out_channels=conv3d_channels, kernel_size=(3, 3, 3), padding=(1, 0, 0)
        )
        for i in range(0, self.number_of_conv3d_layers - 1):
            layer = nn.Conv3d(
                in_channels=conv3d_channels, out_channels=conv3d_channels, kernel_size=(3, 3, 3), padding=(1, 0, 0)
            )
            setattr(self, f"sat_conv{i + 1}", layer)

        self.fc1 = nn.Linear(in_features=self.cnn_output_size, out_features=self.fc1_output_features)
        self.fc2 = nn.Linear(in_features=self.fc1_output_features, out_features=self.fc2_output_features)

        if include_nwp:
            self.nwp_conv0 = nn.Conv3d(
                in_channels=number_nwp_channels, out_channels=conv3d_channels, kernel_size=(3, 3, 3), padding=(1, 0, 0)
            )
            for i in range(0, self.number_of_conv3d_layers - 1):
                layer = nn.Conv3d(
                    in_channels=conv3d_channels, out_channels=conv3d_channels, kernel_size=(3, 3, 3), padding=(1, 0, 0)
                )
                setattr(self, f"nwp_conv{i + 1}", layer)
            self.nwp_fc1 = nn.Linear(in_features=self.nwp_cnn_output_size, out_features=self.fc1_output_features)
            self.nwp_fc2 = nn.Linear(in_features=self.fc1_output_features, out_features=self.number_of_nwp_features)

        if self.embedding_dem:
            self.pv_system_id_embedding = nn.Embedding(num_embeddings=940, embedding_dim=self.embedding_dem)

        if self.include_pv_yield_history:
            self.pv_fc1 = nn.Linear(
                in_features=self.number_of_pv_samples_per_batch * (self.history_len_5 + 1), out_features=128
            )

        fc3_in_features = self.fc2_output_features
        if include_pv_or_gsp_yield_history:
            fc3_in_features += self.number_of_samples_per_batch * (self.history_len_30 + 1)
        if include_nwp:
            fc3_in_features += 128
        if self.embedding_dem:
            fc3_in_features += self.embedding_dem
        if self.include_pv_yield_history:
            fc3_in_features += 128

        self.fc3 = nn.Linear(in_features=fc3_in_features, out_features=self.fc3_output_features)
        self.fc4 = nn.Linear(in_features=self.fc3_output_features, out_features=self.forecast_len)

        # normalisation constants for int16 satellite input (same convention as the single-tower model)
        n = number_sat_channels
        if n <= 12:
            mean, std = SAT_MEAN[12 - n:], SAT_STD[12 - n:]
        else:
            mean, std = np.zeros(n, np.float32), np.ones(n, np.float32)
        self.register_buffer("sat_mean", torch.from_numpy(mean.copy()), persistent=False)
        self.register_buffer("sat_std", torch.from_numpy(std.copy()), persistent=False)
        # bf16 mode: tensor-core shadows of the two big fc1 weights; FusedAdam finds them through the parameters
        self._fc1_shadow = ops.Fc1Shadow()
        self._nwp_fc1_shadow = ops.Fc1Shadow()
        if precision == "bf16":
            self.fc1.weight._pvb_shadow = self._fc1_shadow
            if include_nwp:
                self.nwp_fc1.weight._pvb_shadow = self._nwp_fc1_shadow

    def _tower_params(self, prefix: str):
        wb = []
        for i in range(self.number_of_conv3d_layers):
            layer = getattr(self, f"{prefix}{i}")
            wb += [layer.weight, layer.bias]
        return wb

    def _tower_and_fc1(self, cube, mean, std, prefix, fc1, shadow, n_features, what):
        """[Conv3d(pad (1,0,0)) + ReLU] x L -> flatten -> relu(fc1): fp32 kernels, or tensor cores in bf16 mode (batches up to
        128 when training / 256 forward-only, 32-channel towers; otherwise the fp32 kernels)."""
        wb = self._tower_params(prefix)
        B = cube.shape[0]
        use_tc = (self.precision == "bf16" and B <= (128 if torch.is_grad_enabled() else 256) and fc1.out_features <= 128
                  and wb[0].shape[0] % 16 == 0)
        if use_tc:
            link = {"shadow": shadow, "pad_t": 1, "tag": prefix}
            act = ops.EncoderBf16Fn.apply(link, cube, mean, std, *wb)
            if act[0].numel() != n_features:
                raise RuntimeError(f"{what} cube {tuple(cube.shape)} gives {act[0].numel()} conv features, model expects {n_features}")
            return ops.Fc1Bf16Fn.apply(link, act, fc1.weight, fc1.bias)
        out = ops.TowerFn.apply(1, cube, mean, std, *wb)
        if out.shape[1] != n_features:
            raise RuntimeError(f"{what} cube {tuple(cube.shape)} gives {out.shape[1]} conv features, model expects {n_features}")
        # fc1 hands the tower the gradient of its last PRE-activation (model_sat_nwp.py:195, 243)
        return ops.LinearFn.apply(out, fc1.weight, fc1.bias, True, True)

    @staticmethod
    def _need_cuda(t: torch.Tensor, what: str) -> None:
        if not t.is_cuda:
            raise RuntimeError(
                f"predict_pv_yield_b200 conv3d_sat_nwp is CUDA (sm_100a) only: {what} is on {t.device} (there is no CPU fallback)"
            )

    def invalidate_shadow(self) -> None:
        """bf16 mode: force the tensor-core shadows of ``fc1.weight`` / ``nwp_fc1.weight`` to be rebuilt by the next forward
        (needed only after writes through ``.data``, which torch's version counter does not see; ``ops.Fc1Shadow``)."""
        self._fc1_shadow.key = None
        self._nwp_fc1_shadow.key = None

    def forward(self, x):
        x = as_batch(x)

        # ******************* Satellite imagery *************************
        # Shape: batch_size, channel, seq_length, height, width
        sat_data = x.satellite.data
        self._need_cuda(sat_data, "satellite.data")
        batch_size = sat_data.shape[0]
        if not self.include_future_satellite:
            sat_data = sat_data[:, :, : self.history_len_5 + 1]
        if sat_data.dtype == torch.int16:
            mean, std = self.sat_mean, self.sat_std  # raw SEVIRI counts: normalised on the device (netcdf_dataset.py:96-101)
        else:
            sat_data, mean, std = sat_data.float(), None, None  # model_sat_nwp.py:180
        out = self._tower_and_fc1(sat_data.contiguous(), mean, std, "sat_conv", self.fc1, self._fc1_shadow,
                                  self.cnn_output_size, "satellite")
        out = ops.LinearFn.apply(out, self.fc2.weight, self.fc2.bias, True, False)
        parts = [out]

        # add pv / gsp yield history (model_sat_nwp.py:200-216)
        if self.include_pv_or_gsp_yield_history:
            hist = x.gsp.gsp_yield if self.output_variable == "gsp_yield" else x.pv.pv_yield
            hist = hist if hist.dtype == torch.float32 else hist.float()
            parts.append(ops.history_flatten(hist, self.history_len_30 + 1, hist.shape[2]))

        # the 5-minute PV history through its own layer (model_sat_nwp.py:219-232): first 128 systems
        if self.include_pv_yield_history:
            pv = x.pv.pv_yield
            pv = pv if pv.dtype == torch.float32 else pv.float()
            pv_hist = ops.history_flatten(pv, self.history_len_5 + 1, min(128, pv.shape[2]))
            parts.append(ops.LinearFn.apply(pv_hist, self.pv_fc1.weight, self.pv_fc1.bias, True, False))

        # *********************** NWP Data ************************************
        if self.include_nwp:
            nwp_data = x.nwp.data.float().contiguous()  # shape: batch_size, n_chans, seq_len, height, width
            self._need_cuda(nwp_data, "nwp.data")
            out_nwp = self._tower_and_fc1(nwp_data, None, None, "nwp_conv", self.nwp_fc1, self._nwp_fc1_shadow,
                                          self.nwp_cnn_output_size, "NWP")
            out_nwp = ops.LinearFn.apply(out_nwp, self.nwp_fc2.weight, self.nwp_fc2.bias, True, False)
            parts.append(out_nwp)

        # ********************** Embedding of PV system ID ********************
        if self.embedding_dem:
            if self.output_variable == "pv_yield":
                ids = x.pv.pv_system_row_number[0: self.batch_size, 0]
            else:
                ids = x.gsp.gsp_id[0: self.batch_size, 0]
            # the reference round-trips the ids through the CPU (.type(torch.IntTensor), model_sat_nwp.py:257-258);
            # here they stay on the device
            ids = ids.to(device=out.device, dtype=torch.int32).contiguous()
            parts.append(ops.EmbeddingFn.apply(self.pv_system_id_embedding.weight, ids))

        # join up (torch.cat of tiny [B, n] tensors), then the last two layers (model_sat_nwp.py:262-266)
        out = torch.cat(parts, dim=1) if len(parts) > 1 else parts[0]
        out = ops.LinearFn.apply(out, self.fc3.weight, self.fc3.bias, True, False)
        out = ops.LinearFn.apply(out, self.fc4.weight, self.fc4.bias, False, False)
        out = out.reshape(batch_size, self.forecast_len)
        return out
