"""TEST INFRASTRUCTURE ONLY -- CPU restatement (the "oracle") of the reference two-tower model
``predict_pv_yield/models/conv3d/model_sat_nwp.py``.

Same rules as ``oracle/conv3d_oracle.py``: only ``tests/`` (and the golden generator) may import it; it is pinned
against outputs of the UNMODIFIED reference (``oracle/make_golden.py`` -> ``tests/golden/sat_nwp_*.npz``), bit-for-bit
on the forward pass and the losses.  It calls the same torch CPU operators as the reference, in the same order.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from .conv3d_oracle import _NS, derived_sizes, sat_constants, sat_normalise, weighted_loss_weights


class OracleSatNwpModel(nn.Module):
    """model_sat_nwp.py:14-268 + base_model.py:27-153,255-257; same sub-module names and construction order."""

    name = "conv3d_sat_nwp"
    batch_size = 32  # base_model.py:30

    def __init__(self, include_pv_or_gsp_yield_history=True, include_nwp=True, forecast_minutes=30, history_minutes=60,
                 number_of_conv3d_layers=4, conv3d_channels=32, image_size_pixels=64, nwp_image_size_pixels=64,
                 number_sat_channels=12, number_nwp_channels=10, fc1_output_features=128, fc2_output_features=128,
                 fc3_output_features=64, output_variable="pv_yield", embedding_dem=16, include_pv_yield_history=True,
                 include_future_satellite=True):
        super().__init__()
        self.include_pv_or_gsp_yield_history = include_pv_or_gsp_yield_history
        self.include_nwp = include_nwp
        self.number_of_conv3d_layers = L = number_of_conv3d_layers
        self.output_variable = output_variable
        self.embedding_dem = embedding_dem
        self.include_pv_yield_history = include_pv_yield_history
        self.include_future_satellite = include_future_satellite
        for k, v in derived_sizes(history_minutes, forecast_minutes, output_variable).items():
            setattr(self, k, v)
        self.number_of_pv_samples_per_batch = 128  # base_model.py:74
        self.weights_exp = weighted_loss_weights(self.forecast_len)
        # model_sat_nwp.py:85-99
        t_sat = (self.forecast_len_5 + self.history_len_5 + 1) if include_future_satellite else (self.history_len_5 + 1)
        self.cnn_output_size = conv3d_channels * ((image_size_pixels - 2 * L) ** 2) * t_sat
        self.nwp_cnn_output_size = (conv3d_channels * ((nwp_image_size_pixels - 2 * L) ** 2)
                                    * (self.forecast_len_60 + self.history_len_60 + 1))
        # model_sat_nwp.py:101-172
        self.sat_conv0 = nn.Conv3d(number_sat_channels, conv3d_channels, kernel_size=(3, 3, 3), padding=(1, 0, 0))
        for i in range(L - 1):
            setattr(self, f"sat_conv{i + 1}", nn.Conv3d(conv3d_channels, conv3d_channels, kernel_size=(3, 3, 3), padding=(1, 0, 0)))
        self.fc1 = nn.Linear(self.cnn_output_size, fc1_output_features)
        self.fc2 = nn.Linear(fc1_output_features, fc2_output_features)
        if include_nwp:
            self.nwp_conv0 = nn.Conv3d(number_nwp_channels, conv3d_channels, kernel_size=(3, 3, 3), padding=(1, 0, 0))
            for i in range(L - 1):
                setattr(self, f"nwp_conv{i + 1}",
                        nn.Conv3d(conv3d_channels, conv3d_channels, kernel_size=(3, 3, 3), padding=(1, 0, 0)))
            self.nwp_fc1 = nn.Linear(self.nwp_cnn_output_size, fc1_output_features)
            self.nwp_fc2 = nn.Linear(fc1_output_features, 128)
        if embedding_dem:
            self.pv_system_id_embedding = nn.Embedding(num_embeddings=940, embedding_dim=embedding_dem)
        if include_pv_yield_history:
            self.pv_fc1 = nn.Linear(self.number_of_pv_samples_per_batch * (self.history_len_5 + 1), 128)
        fc3_in = fc2_output_features
        if include_pv_or_gsp_yield_history:
            fc3_in += self.number_of_samples_per_batch * (self.history_len_30 + 1)
        if include_nwp:
            fc3_in += 128
        if embedding_dem:
            fc3_in += embedding_dem
        if include_pv_yield_history:
            fc3_in += 128
        self.fc3 = nn.Linear(fc3_in, fc3_output_features)
        self.fc4 = nn.Linear(fc3_output_features, self.forecast_len)

    # -- model_sat_nwp.py:174-268 ----------------------------------------------------------------------------
    def forward(self, x):
        if isinstance(x, dict):
            x = _NS(**x)
        dt = self.sat_conv0.weight.dtype
        sat = x.satellite.data
        if sat.dtype == torch.int16:  # the step includes the int16 normalisation (netcdf_dataset.py:96-101)
            mean, std = sat_constants(sat.shape[1])
            sat = sat_normalise(sat, torch.from_numpy(mean), torch.from_numpy(std))
        sat = sat.to(dt)  # :180
        B = sat.shape[0]
        if not self.include_future_satellite:  # :183-184
            sat = sat[:, :, : self.history_len_5 + 1]
        out = F.relu(self.sat_conv0(sat))  # :187
        for i in range(self.number_of_conv3d_layers - 1):  # :188-190
            out = F.relu(getattr(self, f"sat_conv{i + 1}")(out))
        out = out.reshape(B, self.cnn_output_size)  # :192
        out = F.relu(self.fc1(out))  # :195
        out = F.relu(self.fc2(out))  # :196
        if self.include_pv_or_gsp_yield_history:  # :200-216
            h = x.gsp.gsp_yield if self.output_variable == "gsp_yield" else x.pv.pv_yield
            h = h[:, : self.history_len_30 + 1].nan_to_num(nan=0.0).to(dt)
            out = torch.cat((out, h.reshape(h.shape[0], h.shape[1] * h.shape[2])), dim=1)
        if self.include_pv_yield_history:  # :219-232
            h = x.pv.pv_yield[:, : self.history_len_5 + 1, :128].nan_to_num(nan=0.0).to(dt)
            h = F.relu(self.pv_fc1(h.reshape(h.shape[0], h.shape[1] * h.shape[2])))
            out = torch.cat((out, h), dim=1)
        if self.include_nwp:  # :235-249
            n = x.nwp.data.to(dt)
            n = F.relu(self.nwp_conv0(n))
            for i in range(self.number_of_conv3d_layers - 1):
                n = F.relu(getattr(self, f"nwp_conv{i + 1}")(n))
            n = n.reshape(B, self.nwp_cnn_output_size)
            n = F.relu(self.nwp_fc2(F.relu(self.nwp_fc1(n))))
            out = torch.cat((out, n), dim=1)
        if self.embedding_dem:  # :252-260
            ids = x.pv.pv_system_row_number[0: self.batch_size, 0] if self.output_variable == "pv_yield" \
                else x.gsp.gsp_id[0: self.batch_size, 0]
            out = torch.cat((out, self.pv_system_id_embedding(ids.type(torch.IntTensor))), dim=1)
        out = F.relu(self.fc3(out))  # :263
        out = self.fc4(out)  # :264
        return out.reshape(B, self.forecast_len)  # :266

    # -- base_model.py:78-146 --------------------------------------------------------------------------------
    def step_losses(self, batch):
        if isinstance(batch, dict):
            batch = _NS(**batch)
        y_hat = self(batch)
        y = batch.gsp.gsp_yield if self.output_variable == "gsp_yield" else batch.pv.pv_yield
        y = y[0: self.batch_size, -self.forecast_len:, 0].to(y_hat.dtype)
        w = self.weights_exp.to(y_hat.dtype)
        return dict(mse=F.mse_loss(y_hat, y), nmae=(y_hat - y).abs().mean(), mse_exp=torch.mean(w * (y_hat - y) ** 2),
                    mae_exp=torch.mean(w * torch.abs(y_hat - y)), y_hat=y_hat)

    def training_step(self, batch, batch_idx=0):
        return self.step_losses(batch)["nmae"]

    def configure_optimizers(self):
        return torch.optim.Adam(self.parameters(), lr=0.0005)
