"""BaseModel -- host-side mirror of ``predict_pv_yield/models/base_model.py:27-257`` for the Conv3d path.

Same class attributes (``batch_size = 32``, ``results_file_name``, ``results_dfs``), same derived
sizes (``base_model.py:41-74``), same step (``_training_or_validation_step``, ``:78-146``): forward,
target slice ``y[0:batch_size, -forecast_len:, 0]``, the L1 loss that is returned plus MSE and the two
exponentially weighted losses that are logged, the per-horizon metrics on validation/test (including
the reference's key collision at ``:126-136``, kept for log compatibility) and ``configure_optimizers``
(Adam, lr 5e-4).  The arithmetic runs in ``libpvb200.so`` (``ops.StepLossFn``, ``optim.FusedAdam``).

Works with or without ``pytorch_lightning`` installed (the reference targets Lightning 1.4-1.5, which
is not in this image): when it is importable the class derives from ``pl.LightningModule``.
The capacity-scaled validation table of ``validation_step`` (``:222-250``) is produced on the device in one kernel
(``predict_pv_yield_b200/validation.py``); only the plotting (``:165-220``) is out of scope.
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

        def log_dict(self, dictionary, *args, sync_dist: bool = False, **kwargs):
            # Lightning's ``sync_dist=True`` (base_model.py:108-119) averages every logged scalar over the ranks, one
            # collective per scalar; without Lightning the same mean is taken here as ONE all-reduce of the packed
            # vector when a process group is initialised (dp.reduce_logged_scalars)
            if sync_dist and torch.distributed.is_available() and torch.distributed.is_initialized():
                from ..dp import reduce_logged_scalars

                dictionary = reduce_logged_scalars(dictionary)
            self.logged_metrics = {**getattr(self, "logged_metrics", {}), **dictionary}

        def log(self, name, value, *args, **kwargs):
            self.log_dict({name: value})


default_output_variable = "pv_yield"


class BaseModel(_Base):

    # default batch_size (base_model.py:30)
    batch_size = 32

    # not in the reference: see configure_optimizers
    overlap_optimizer = False

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
            # metrics for each forecast horizon (nowcasting_utils.models.metrics: mean over the batch axis) and the
            # capacity-scaled validation results, in one kernel
            cap = self._gsp_capacity_view(batch, y_hat)
            with torch.no_grad():
                out, horizon = ops.validation_results(y_hat.detach(), y, cap)
            self._last_validation = (out, cap is not None)
            mse_h, mae_h = horizon[0], horizon[1]
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

    def _gsp_capacity_view(self, batch, y_hat):
        """``batch.gsp.gsp_capacity[:, -forecast_len_30:, 0]`` (base_model.py:223) when the batch carries it and the
        model forecasts GSP yield at 30-minute steps (the only case in which the reference's scaling is shape-correct)."""
        gsp = getattr(batch, "gsp", None)
        cap = getattr(gsp, "gsp_capacity", None)
        if cap is None or self.output_variable != "gsp_yield":
            return None
        if not torch.is_tensor(cap) or cap.dim() != 3 or cap.shape[0] < y_hat.shape[0] or cap.shape[1] < y_hat.shape[1]:
            return None
        cap = cap[0: y_hat.shape[0], -y_hat.shape[1]:, 0]
        return cap if cap.dtype == torch.float32 else cap.float()

    def validation_step(self, batch, batch_idx):
        batch = as_batch(batch)
        nmae_loss, _ = self._training_or_validation_step(batch, tag="Validation", return_model_outputs=True)
        # save validation results (base_model.py:222-241): capacity-scaled MW table, one device -> host copy per batch
        out, has_capacity = self._last_validation
        gsp, meta = getattr(batch, "gsp", None), getattr(batch, "metadata", None)
        if has_capacity and getattr(gsp, "gsp_id", None) is not None and getattr(meta, "t0_datetime_utc", None) is not None:
            from ..validation import make_validation_results

            host = out.cpu().numpy()  # [3, B, forecast_len_30]
            t0 = meta.t0_datetime_utc
            t0 = t0.cpu().numpy() if torch.is_tensor(t0) else t0
            results = make_validation_results(truths_mw=host[1], predictions_mw=host[0], capacity_mwp=host[2],
                                              gsp_ids=gsp.gsp_id[0: host.shape[1], 0], batch_idx=batch_idx, t0_datetimes_utc=t0)
            if batch_idx == 0:
                self.results_dfs = []
            self.results_dfs.append(results)
        return nmae_loss

    def validation_epoch_end(self, outputs):
        logger.info("Validation epoch end")
        from ..validation import save_validation_results

        save_validation_results(self.results_dfs, self.results_file_name, self.current_epoch)

    def test_step(self, batch, batch_idx):
        self._training_or_validation_step(batch, tag="Test")

    def predict_step(self, batch, batch_idx=0, dataloader_idx=0):
        with torch.no_grad():
            return self(batch)

    def configure_optimizers(self):
        # base_model.py:255-257: torch.optim.Adam(self.parameters(), lr=0.0005)
        # overlap_optimizer (class attribute, default False): run the Adam step of fc1.weight on a side stream under the next
        # step's convolution forward (optim.FusedAdam.overlap_large)
        return FusedAdam(self.parameters(), lr=0.0005, overlap_large=bool(getattr(self, "overlap_optimizer", False)))
