"""BaseModel -- host-side mirror of ``predict_pv_yield/models/base_model.py:27-257`` for the Conv3d path.

Same class attributes (``batch_size = 32``, ``results_file_name``, ``results_dfs``), same derived
sizes (``base_model.py:41-74``), same step (``_training_or_validation_step``, ``:78-146``): forward,
target slice ``y[0:batch_size, -forecast_len:, 0]``, the L1 loss that is returned plus MSE and the two
exponentially weighted losses that are logged, the per-horizon metrics on validation/test (including
the reference's key collision at ``:126-136``, kept for log compatibility) and ``configure_optimizers``
(Adam, lr 5e-4).  The arithmetic runs in ``libpvb200.so`` (``ops.StepLossFn``, ``optim.FusedAdam``).

Works with or without ``pytorch_lightning`` installed (the reference targets Lightning 1.4-1.5, which
is not in this image): when it is importable the class derives from ``pl.LightningModule``.
Out of scope: the plotting / pandas CSV tail of ``validation_step`` (``:165-241``).
"""
from __future__ import annotations

import logging

import numpy as np
import torch

from ..batch import as_batch
from ..losses import WeightedLosses
from ..optim import FusedAdam
from .. import ops

logger = logging.getLogger(__name__)

try:  # pragma: no cover - depends on the environment
    import pytorch_lightning as pl

    _Base = pl.LightningModule
except Exception:  # pytorch_lightning absent: minimal stand-in with the hooks the step touches

    class _Base(torch.nn.Module):
        current_epoch = 0
        logger = None

        def log_dict(self, dictionary, *args, **kwargs):
            self.logged_metrics = {**getattr(self, "logged_metrics", {}), **dictionary}

        def log(self, name, value, *args, **kwargs):
            self.log_dict({name: value})


default_output_variable = "pv_yield"


class BaseModel(_Base):

    # default batch_size (base_model.py:30)
    batch_size = 32

    # results file name
    results_file_name = "results_epoch"

    # list of results dataframes. This is used to save validation results
    results_dfs = []

    def __init__(self):
        super().__init__()

        # base_model.py:41-59
        self.history_len_5 = self.history_minutes // 5
        self.forecast_len_5 = self.forecast_minutes // 5
        self.history_len_30 = self.history_minutes // 30
        self.forecast_len_30 = self.forecast_minutes // 30
        self.history_len_60 = int(np.ceil(self.history_minutes / 60))
        self.forecast_len_60 = self.forecast_minutes // 60

        if not hasattr(self, "output_variable"):
            self.output_variable = default_output_variable

        # base_model.py:66-74
        if self.output_variable == "pv_yield":
            self.forecast_len = self.forecast_len_5
            self.history_len = self.history_len_5
            self.number_of_samples_per_batch = 128
        else:
            self.forecast_len = self.forecast_len_30
            self.history_len = self.history_len_30
            self.number_of_samples_per_batch = 32
        self.number_of_pv_samples_per_batch = 128

        self.weighted_losses = WeightedLosses(forecast_length=self.forecast_len)
        self.register_buffer("_loss_weights", self.weighted_losses.weights.clone(), persistent=False)

    def _training_or_validation_step(self, batch, tag: str, return_model_outputs: bool = False):
        """
        batch: The batch data
        tag: either 'Train', 'Validation' , 'Test'
        """
        batch = as_batch(batch)

        # put the batch data through the model
        y_hat = self(batch)

        # the target: first system = the one at the centre of the image (base_model.py:90-95)
        if self.output_variable == "gsp_yield":
            y = batch.gsp.gsp_yield
        else:
            y = batch.pv.pv_yield
        y = y[0: self.batch_size, -self.forecast_len:, 0]
        if y.dtype != torch.float32:
            y = y.float()

        # fused: nmae (L1, returned), mse, mse_exp, mae_exp  (base_model.py:98-103)
        losses = ops.StepLossFn.apply(y_hat, y, self._loss_weights)
        nmae_loss, mse_loss, mse_exp, mae_exp = losses[0], losses[1], losses[2], losses[3]

        self.log_dict(
            {
                f"MSE/{tag}": mse_loss,
                f"NMAE/{tag}": nmae_loss,
                f"MSE_EXP/{tag}": mse_exp,
                f"MAE_EXP/{tag}": mae_exp,
            },
            on_step=True,
            on_epoch=True,
            sync_dist=True,  # Required for distributed training
        )

        if tag != "Train":
            # metrics for each forecast horizon (nowcasting_utils.models.metrics: mean over the batch axis);
            # logging-only, tiny [B, forecast_len] tensors
            with torch.no_grad():
                d = y_hat.detach() - y
                mse_h = (d * d).mean(dim=0)
                mae_h = d.abs().mean(dim=0)
            metrics_mse = {f"MSE_forecast_horizon_{i}/{tag}": mse_h[i] for i in range(self.forecast_len_30)}
            # NOTE: the reference logs the MAE under the MSE key as well (base_model.py:131-134), so the MAE
            # values overwrite the MSE ones; reproduced so dashboards keyed on these names see the same numbers
            metrics_mae = {f"MSE_forecast_horizon_{i}/{tag}": mae_h[i] for i in range(self.forecast_len_30)}
            self.log_dict({**metrics_mse, **metrics_mae}, on_step=True, on_epoch=True, sync_dist=True)

        if return_model_outputs:
            return nmae_loss, y_hat
        else:
            return nmae_loss

    def training_step(self, batch, batch_idx):
        return self._training_or_validation_step(batch, tag="Train")

    def validation_step(self, batch, batch_idx):
        nmae_loss, _ = self._training_or_validation_step(batch, tag="Validation", return_model_outputs=True)
        return nmae_loss

    def validation_epoch_end(self, outputs):
        logger.info("Validation epoch end")

    def test_step(self, batch, batch_idx):
        self._training_or_validation_step(batch, tag="Test")

    def predict_step(self, batch, batch_idx=0, dataloader_idx=0):
        with torch.no_grad():
            return self(batch)

    def configure_optimizers(self):
        # base_model.py:255-257: torch.optim.Adam(self.parameters(), lr=0.0005)
        return FusedAdam(self.parameters(), lr=0.0005)
