"""TEST INFRASTRUCTURE ONLY -- dependency shims that let the UNMODIFIED reference model import here.

Only ``oracle/make_golden.py`` (run in the build container, where ``/root/reference`` exists)
uses this file.  Nothing under ``predict_pv_yield_b200/`` may import it, and nothing in the
``-m gpu`` tests / ``smoke()`` / ``bench.py`` can: ``/root/reference`` is absent on the GPU box.

The reference (``/root/reference/predict_pv_yield/models/base_model.py:1-11`` and
``models/conv3d/model.py:8``) imports ``pytorch_lightning``, ``nowcasting_dataloader``,
``nowcasting_utils`` and ``nowcasting_dataset``; none is installed in this image and there is
no network.  We register minimal stand-ins in ``sys.modules`` *before* importing the reference
file, so that the reference's own arithmetic (torch ``nn.Conv3d`` / ``nn.Linear`` / ``F.relu`` /
``torch.cat`` / ``F.mse_loss`` / ``torch.optim.Adam``) runs unmodified.

Stand-ins and what they restate:

* ``pytorch_lightning.LightningModule``  -> ``torch.nn.Module`` with a no-op ``log_dict`` that
  records the last logged dict, and ``current_epoch = 0``.
* ``nowcasting_dataloader.batch.BatchML`` -> nested namespace with attribute and item access
  (the reference uses both: ``x.satellite.data`` at ``model.py:113`` and ``x["nwp"]`` at ``:141``).
* ``nowcasting_utils.models.loss.WeightedLosses`` -> restated from the published package
  (``nowcasting_utils/models/loss.py``, unpinned in ``requirements.txt:2``): weights
  ``exp(-ln2 * i)`` normalised to mean 1; ``get_mse_exp = mean(w * (o - t)**2)``,
  ``get_mae_exp = mean(w * |o - t|)``.
* the remaining imports (plot helpers, validation-result writers, NWP names) are placeholders:
  they are only touched by ``validation_step`` plotting/CSV code, which is out of scope.
"""
from __future__ import annotations

import math
import sys
import types

import torch

REFERENCE_ROOT = "/root/reference"


class _Namespace:
    """Attribute- and item-accessible view over a (nested) dict."""

    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, _Namespace(**v) if isinstance(v, dict) else v)

    def __getitem__(self, key):
        return getattr(self, key)

    def __contains__(self, key):
        return hasattr(self, key)


class _WeightedLosses:
    def __init__(self, decay_rate=None, forecast_length: int = 6):
        self.decay_rate = math.log(2) if decay_rate is None else decay_rate
        self.forecast_length = forecast_length
        w = torch.FloatTensor([math.exp(-self.decay_rate * i) for i in range(forecast_length)])
        self.weights = w / w.sum() * len(w)

    def get_mse_exp(self, output, target):
        return torch.mean(self.weights.to(output.device) * (output - target) ** 2)

    def get_mae_exp(self, output, target):
        return torch.mean(self.weights.to(output.device) * torch.abs(output - target))


class _LightningModule(torch.nn.Module):
    current_epoch = 0
    logger = None

    def log_dict(self, d, *a, **k):
        self._last_logged = {kk: (vv.detach().clone() if torch.is_tensor(vv) else vv) for kk, vv in d.items()}

    def log(self, *a, **k):
        pass


def _mod(name: str, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _placeholder(*a, **k):
    raise NotImplementedError("out-of-scope reference dependency (validation plotting / CSV)")


def install():
    """Register the stand-in modules (idempotent)."""
    if "pytorch_lightning" not in sys.modules:
        _mod("pytorch_lightning", LightningModule=_LightningModule)
    for pkg in (
        "nowcasting_dataloader",
        "nowcasting_utils",
        "nowcasting_utils.visualization",
        "nowcasting_utils.models",
        "nowcasting_utils.metrics",
        "nowcasting_dataset",
        "nowcasting_dataset.data_sources",
        "nowcasting_dataset.data_sources.nwp",
    ):
        if pkg not in sys.modules:
            _mod(pkg)
    _mod("nowcasting_dataloader.batch", BatchML=_Namespace)
    _mod("nowcasting_utils.visualization.visualization", plot_example=_placeholder)
    _mod("nowcasting_utils.visualization.line", plot_batch_results=_placeholder)
    _mod("nowcasting_utils.models.loss", WeightedLosses=_WeightedLosses)
    _mod(
        "nowcasting_utils.models.metrics",
        mae_each_forecast_horizon=lambda output, target: (output - target).abs().mean(dim=0),
        mse_each_forecast_horizon=lambda output, target: ((output - target) ** 2).mean(dim=0),
    )
    _mod(
        "nowcasting_utils.metrics.validation",
        make_validation_results=_placeholder,
        save_validation_results_to_logger=_placeholder,
    )
    _mod("nowcasting_dataset.data_sources.nwp.nwp_data_source", NWP_VARIABLE_NAMES=tuple("abcdefghij"))


def import_reference_model():
    """Return the reference ``Model`` class (``predict_pv_yield/models/conv3d/model.py:14``), unmodified."""
    install()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from predict_pv_yield.models.conv3d.model import Model  # noqa: E402  (the real reference file)

    return Model


def import_reference_sat_nwp_model():
    """Return the reference two-tower ``Model`` class (``predict_pv_yield/models/conv3d/model_sat_nwp.py:14``), unmodified."""
    install()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from predict_pv_yield.models.conv3d.model_sat_nwp import Model  # noqa: E402  (the real reference file)

    return Model


def import_reference_conv3d_maxpool():
    """Return the reference ``Conv3dMaxPool`` class (``predict_pv_yield/models/perceiver/perceiver_conv3d_nwp_sat.py:42``),
    unmodified.  Its module also imports ``perceiver_pytorch`` and ``nowcasting_dataset.consts`` (absent here): stand-ins."""
    install()
    if "perceiver_pytorch" not in sys.modules:
        _mod("perceiver_pytorch", Perceiver=type("Perceiver", (torch.nn.Module,), {}))
    if "nowcasting_dataset.consts" not in sys.modules:
        _mod("nowcasting_dataset.consts", NWP_VARIABLE_NAMES=tuple("abcdefghijklmnop"), SAT_VARIABLE_NAMES=tuple("abcdefghijkl"))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from predict_pv_yield.models.perceiver.perceiver_conv3d_nwp_sat import Conv3dMaxPool  # noqa: E402

    return Conv3dMaxPool
