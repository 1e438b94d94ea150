"""TEST INFRASTRUCTURE ONLY -- CPU restatement (the "oracle") of the reference Conv3d PV-yield step.

Who may import this file: ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs.  Never the product package
(``predict_pv_yield_b200/``): the product path has no CPU fallback.

Parity pinning: the reference's own tests hold NO numeric golden vectors for this path
(shapes only), so the oracle is pinned against outputs of the reference itself:
``oracle/make_golden.py`` imports the UNMODIFIED reference ``Model`` (through
``oracle/ref_shims.py``) in the build container, runs it on seeded inputs and commits the
results under ``tests/golden/``; ``tests/test_oracle.py`` checks this restatement against those
vectors bit-for-bit (forward, loss) / to 1e-6 (gradients, Adam) on CPU.

Every function cites the reference lines it follows (paths relative to ``/root/reference``).
The arithmetic itself lives in third-party ``torch`` (unpinned, ``requirements.txt:10``); this
file calls the same torch CPU operators the reference calls, in the same order.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

# ---------------------------------------------------------------------------------------------
# a1. int16 satellite normalisation -- predict_pv_yield/netcdf_dataset.py:16-32 (constants),
#     :96-101 (operation); same op in notebooks/15_int16.ipynb:13721-13727.
# ---------------------------------------------------------------------------------------------
SAT_VARIABLE_NAMES = (
    "HRV", "IR_016", "IR_039", "IR_087", "IR_097", "IR_108", "IR_120",
    "IR_134", "VIS006", "VIS008", "WV_062", "WV_073",
)
SAT_MEAN = np.array(
    [93.23458, 131.71373, 843.7779, 736.6148, 771.1189, 589.66034,
     862.29816, 927.69586, 90.70885, 107.58985, 618.4583, 532.47394], dtype=np.float32)
SAT_STD = np.array(
    [115.34247, 139.92636, 36.99538, 57.366386, 30.346825,
     149.68007, 51.70631, 35.872967, 115.77212, 120.997154,
     98.57828, 99.76469], dtype=np.float32)


def sat_constants(n_channels: int):
    """Mean/std for an ``n_channels`` cube.  12 -> all; fewer -> the LAST n (HRV, index 0, is the
    channel the 11-channel production config drops)."""
    assert 1 <= n_channels <= 12
    return SAT_MEAN[12 - n_channels:].copy(), SAT_STD[12 - n_channels:].copy()


def sat_normalise_numpy(x: np.ndarray, mean: np.ndarray, std: np.ndarray) -> np.ndarray:
    """netcdf_dataset.py:96-101 verbatim in numpy: astype(float32); - mean; /= std.  x: [B,C,T,H,W]."""
    assert x.dtype == np.int16
    y = x.astype(np.float32)
    y = y - mean.astype(np.float32).reshape(1, -1, 1, 1, 1)
    y /= std.astype(np.float32).reshape(1, -1, 1, 1, 1)
    return y


def sat_normalise(x: torch.Tensor, mean: torch.Tensor, std: torch.Tensor) -> torch.Tensor:
    """Same arithmetic in torch fp32 (two IEEE roundings: subtract, then true division)."""
    assert x.dtype == torch.int16
    y = x.to(torch.float32)
    y = y - mean.to(torch.float32).view(1, -1, 1, 1, 1)
    y = y / std.to(torch.float32).view(1, -1, 1, 1, 1)
    return y


# ---------------------------------------------------------------------------------------------
# a13. derived sizes -- predict_pv_yield/models/base_model.py:41-74
# ---------------------------------------------------------------------------------------------
def derived_sizes(history_minutes: int, forecast_minutes: int, output_variable: str = "pv_yield") -> Dict[str, int]:
    d = dict(
        history_len_5=history_minutes // 5,
        forecast_len_5=forecast_minutes // 5,
        history_len_30=history_minutes // 30,
        forecast_len_30=forecast_minutes // 30,
        history_len_60=int(np.ceil(history_minutes / 60)),
        forecast_len_60=forecast_minutes // 60,
    )
    if output_variable == "pv_yield":
        d.update(forecast_len=d["forecast_len_5"], history_len=d["history_len_5"], number_of_samples_per_batch=128)
    else:
        d.update(forecast_len=d["forecast_len_30"], history_len=d["history_len_30"], number_of_samples_per_batch=32)
    d["number_of_pv_samples_per_batch"] = 128
    return d


# ---------------------------------------------------------------------------------------------
# nowcasting_utils.models.loss.WeightedLosses (third-party, unpinned requirements.txt:2; call
# sites base_model.py:76,102-103) -- restated from the published package.
# ---------------------------------------------------------------------------------------------
def weighted_loss_weights(forecast_length: int, decay_rate: Optional[float] = None) -> torch.Tensor:
    decay_rate = math.log(2) if decay_rate is None else decay_rate
    w = torch.FloatTensor([math.exp(-decay_rate * i) for i in range(forecast_length)])
    return w / w.sum() * len(w)


class _NS:
    """Attribute+item view of a nested dict (what the reference's BatchML offers on this path)."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, _NS(**v) if isinstance(v, dict) else v)

    def __getitem__(self, k):
        return getattr(self, k)


# ---------------------------------------------------------------------------------------------
# a3-a12. the model, step, loss and optimiser
# ---------------------------------------------------------------------------------------------
class OracleModel(nn.Module):
    """Restatement of predict_pv_yield/models/conv3d/model.py:14-156 + base_model.py:27-153,255-257.

    Same constructor arguments and defaults (model.py:18-32), same sub-module names (so the
    ``state_dict`` keys match the reference: ``sat_conv0``, ``conv3d_{i}``, ``fc1``..``fc4``,
    ``fc_nwp``), same construction ORDER (so torch's default init draws the same random stream
    as the reference under the same seed).
    """

    name = "conv3d"
    batch_size = 32  # base_model.py:30

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
    ):
        super().__init__()
        self.include_pv_yield = include_pv_yield
        self.include_nwp = include_nwp
        self.number_of_conv3d_layers = number_of_conv3d_layers
        self.number_of_nwp_features = 10 * 19 * 2 * 2  # model.py:60,72
        self.fc1_output_features = fc1_output_features
        self.fc2_output_features = fc2_output_features
        self.fc3_output_features = fc3_output_features
        self.forecast_minutes = forecast_minutes
        self.history_minutes = history_minutes
        self.output_variable = output_variable
        self.number_sat_channels = number_sat_channels
        for k, v in derived_sizes(history_minutes, forecast_minutes, output_variable).items():
            setattr(self, k, v)
        self.weights_exp = weighted_loss_weights(self.forecast_len)

        L = number_of_conv3d_layers
        # model.py:74-78
        self.cnn_output_size = (
            conv3d_channels * ((image_size_pixels - 2 * L) ** 2) * (self.forecast_len_5 + self.history_len_5 + 1 - 2 * L)
        )
        # model.py:80-90
        self.sat_conv0 = nn.Conv3d(number_sat_channels, conv3d_channels, kernel_size=(3, 3, 3), padding=0)
        for i in range(L - 1):
            setattr(self, f"conv3d_{i + 1}", nn.Conv3d(conv3d_channels, conv3d_channels, kernel_size=(3, 3, 3), padding=0))
        # model.py:92-103
        self.fc1 = nn.Linear(self.cnn_output_size, fc1_output_features)
        self.fc2 = nn.Linear(fc1_output_features, fc2_output_features)
        fc3_in = fc2_output_features
        if include_pv_yield:
            fc3_in += self.number_of_samples_per_batch * (self.history_len_30 + 1)
        if include_nwp:
            self.fc_nwp = nn.Linear(self.number_of_nwp_features, 128)
            fc3_in += 128
        self.fc3 = nn.Linear(fc3_in, fc3_output_features)
        self.fc4 = nn.Linear(fc3_output_features, self.forecast_len)

    # -- model.py:107-156 ------------------------------------------------------------------
    def forward(self, x, return_activations: bool = False):
        if isinstance(x, dict):
            x = _NS(**x)
        sat = x.satellite.data
        if sat.dtype == torch.int16:  # a1: the step includes the int16 normalisation (north star)
            mean, std = sat_constants(sat.shape[1])
            sat = sat_normalise(sat, torch.from_numpy(mean), torch.from_numpy(std))
        sat = sat.to(self.sat_conv0.weight.dtype)  # model.py:113 (.float(); .double() for the fp64 truth)
        B = sat.shape[0]
        acts = []
        out = F.relu(self.sat_conv0(sat))  # model.py:117
        acts.append(out)
        for i in range(self.number_of_conv3d_layers - 1):  # model.py:118-120
            out = F.relu(getattr(self, f"conv3d_{i + 1}")(out))
            acts.append(out)
        out = out.reshape(B, self.cnn_output_size)  # model.py:122
        out = F.relu(self.fc1(out))  # model.py:125
        out = F.relu(self.fc2(out))  # model.py:126
        if self.include_pv_yield:  # model.py:130-136
            h = x[self.output_variable][:, : self.history_len_30 + 1].nan_to_num(nan=0.0).to(out.dtype)
            h = h.reshape(h.shape[0], h.shape[1] * h.shape[2])
            out = torch.cat((out, h), dim=1)
        if self.include_nwp:  # model.py:139-148
            nwp = x["nwp"].to(out.dtype).flatten(start_dim=1)
            out = torch.cat((out, F.relu(self.fc_nwp(nwp))), dim=1)
        out = F.relu(self.fc3(out))  # model.py:151
        out = self.fc4(out)  # model.py:152
        out = out.reshape(B, self.forecast_len)  # model.py:154
        if return_activations:
            return out, acts
        return out

    # -- base_model.py:78-146 --------------------------------------------------------------
    def step_losses(self, batch):
        """Returns dict(nmae, mse, mse_exp, mae_exp, y_hat).  ``nmae`` is the loss that is
        back-propagated (base_model.py:99,146)."""
        if isinstance(batch, dict):
            batch = _NS(**batch)
        y_hat = self(batch)
        y = batch.gsp.gsp_yield if self.output_variable == "gsp_yield" else batch.pv.pv_yield  # :90-94
        y = y[0: self.batch_size, -self.forecast_len:, 0].to(y_hat.dtype)  # :95
        w = self.weights_exp.to(y_hat.dtype)
        return dict(
            mse=F.mse_loss(y_hat, y),  # :98
            nmae=(y_hat - y).abs().mean(),  # :99
            mse_exp=torch.mean(w * (y_hat - y) ** 2),  # :102
            mae_exp=torch.mean(w * torch.abs(y_hat - y)),  # :103
            y_hat=y_hat,
        )

    def training_step(self, batch, batch_idx=0):  # base_model.py:148-153
        return self.step_losses(batch)["nmae"]

    def configure_optimizers(self):  # base_model.py:255-257
        return torch.optim.Adam(self.parameters(), lr=0.0005)


# ---------------------------------------------------------------------------------------------
# bf16-mode checker: the same model in fp64 with the ROUNDING POINTS of the bf16 tensor-core path
# ---------------------------------------------------------------------------------------------
class _RoundFwd(torch.autograd.Function):
    """value -> nearest bf16 in the forward pass, identity for the gradient (a tensor STORED as bf16)."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.float32).to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundBwd(torch.autograd.Function):
    """identity in the forward pass, gradient -> nearest bf16 (a GRADIENT tensor stored as bf16)."""

    @staticmethod
    def forward(ctx, x):
        return x

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.float32).to(torch.bfloat16).to(g.dtype)


class Bf16EmulatedOracle(OracleModel):
    """OracleModel evaluated in fp64 with every tensor the bf16 mode stores as bf16 rounded at the same place: the
    normalised input cube, every conv activation (after bias + ReLU), the conv / fc1 weights as the tensor cores read
    them, and on the way back the gradients w.r.t. the conv and fc1 pre-activations.  Accumulation is exact (fp64) where
    the device accumulates in fp32, biases and the small layers of the head stay unrounded (fp32 on the device).
    A bf16-mode result is expected within a few bf16 ulps (2^-8 = 3.9e-3 each) of THIS model -- far tighter than the
    2e-2 the north star allows against the fp32 reference, and tight enough to catch a wrong tap or a missing term.
    Use in .double()."""

    def forward(self, x, return_activations: bool = False):
        if isinstance(x, dict):
            x = _NS(**x)
        rf, rb = _RoundFwd.apply, _RoundBwd.apply
        sat = x.satellite.data
        if sat.dtype == torch.int16:
            mean, std = sat_constants(sat.shape[1])
            sat = sat_normalise(sat, torch.from_numpy(mean), torch.from_numpy(std))
        dt = self.sat_conv0.weight.dtype
        out = rf(sat.to(dt))
        B = sat.shape[0]
        convs = [self.sat_conv0] + [getattr(self, f"conv3d_{i + 1}") for i in range(self.number_of_conv3d_layers - 1)]
        acts = []
        for conv in convs:
            out = rf(F.relu(rb(F.conv3d(out, rf(conv.weight), conv.bias))))
            acts.append(out)
        out = out.reshape(B, self.cnn_output_size)
        out = F.relu(rb(F.linear(out, rf(self.fc1.weight), self.fc1.bias)))
        out = F.relu(self.fc2(out))
        if self.include_pv_yield:
            h = x[self.output_variable][:, : self.history_len_30 + 1].nan_to_num(nan=0.0).to(out.dtype)
            h = h.reshape(h.shape[0], h.shape[1] * h.shape[2])
            out = torch.cat((out, h), dim=1)
        if self.include_nwp:
            nwp = x["nwp"].to(out.dtype).flatten(start_dim=1)
            out = torch.cat((out, F.relu(self.fc_nwp(nwp))), dim=1)
        out = F.relu(self.fc3(out))
        out = self.fc4(out)
        out = out.reshape(B, self.forecast_len)
        if return_activations:
            return out, acts
        return out


# ---------------------------------------------------------------------------------------------
# a12. Adam, restated as scalar arithmetic (torch.optim.Adam single-tensor path, defaults
# betas=(0.9,0.999), eps=1e-8, weight_decay=0, amsgrad=False; base_model.py:256)
# ---------------------------------------------------------------------------------------------
def adam_step_numpy(p, g, m, v, step: int, lr=5e-4, beta1=0.9, beta2=0.999, eps=1e-8):
    """In-place on float32 numpy arrays.  ``step`` is the 1-based step count AFTER increment."""
    f = np.float32
    m += (g - m) * f(1 - beta1)  # exp_avg.lerp_(grad, 1-beta1)
    v *= f(beta2)
    v += f(1 - beta2) * g * g  # addcmul_
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    step_size = lr / bc1
    denom = np.sqrt(v) / f(math.sqrt(bc2)) + f(eps)
    p += f(-step_size) * (m / denom)
    return p, m, v


# ---------------------------------------------------------------------------------------------
# Synthetic inputs (seed 518 = configs/experiment/conv3d.yaml:16)
# ---------------------------------------------------------------------------------------------
def make_synthetic_batch(
    batch: int,
    number_sat_channels: int = 12,
    seq_len: int = 19,
    image_size_pixels: int = 64,
    output_variable: str = "pv_yield",
    n_yield_timesteps: Optional[int] = None,
    seed: int = 518,
    include_legacy_keys: bool = True,
    nan_in_history: bool = True,
    sat_dtype=torch.int16,
) -> dict:
    """Nested batch dict with the keys the path reads (model.py:113,131,141; base_model.py:90-95)."""
    g = torch.Generator().manual_seed(seed)
    sat = torch.randint(0, 1024, (batch, number_sat_channels, seq_len, image_size_pixels, image_size_pixels),
                        generator=g, dtype=torch.int32)
    missing = torch.rand(sat.shape, generator=g) < 1e-3
    sat = torch.where(missing, torch.full_like(sat, -1), sat).to(torch.int16)
    if sat_dtype != torch.int16:
        mean, std = sat_constants(number_sat_channels)
        sat = sat_normalise(sat, torch.from_numpy(mean), torch.from_numpy(std)).to(sat_dtype)
    n_sys = 128 if output_variable == "pv_yield" else 32
    n_t = n_yield_timesteps if n_yield_timesteps is not None else seq_len
    yld = torch.rand((batch, n_t, n_sys), generator=g, dtype=torch.float32)
    legacy = yld.clone()
    if nan_in_history:  # NaNs only in the first history row of the legacy (model-input) copy
        nanmask = torch.rand((batch, n_sys), generator=g) < 0.05
        legacy[:, 0, :] = torch.where(nanmask, torch.full_like(legacy[:, 0, :], float("nan")), legacy[:, 0, :])
    nwp = torch.randn((batch, 10, 19, 2, 2), generator=g, dtype=torch.float32)
    b = {"satellite": {"data": sat}}
    if output_variable == "pv_yield":
        b["pv"] = {"pv_yield": yld}
    else:
        b["gsp"] = {"gsp_yield": yld}
    if include_legacy_keys:
        b[output_variable] = legacy
        b["nwp"] = nwp
    return b


def batch_to(batch: dict, device=None, float_dtype=None) -> dict:
    out = {}
    for k, v in batch.items():
        if isinstance(v, dict):
            out[k] = batch_to(v, device, float_dtype)
        elif torch.is_tensor(v):
            t = v
            if float_dtype is not None and t.is_floating_point():
                t = t.to(float_dtype)
            out[k] = t.to(device) if device is not None else t
        else:
            out[k] = v
    return out


def normalised_max_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max|a-b| / max|b| -- the 'rel err' definition used by every parity gate."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    denom = float(b.abs().max())
    if denom == 0.0:
        return float((a - b).abs().max())
    return float((a - b).abs().max()) / denom
