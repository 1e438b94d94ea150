"""``WeightedLosses`` -- exponentially weighted MSE / MAE logged next to the loss.

Restates ``nowcasting_utils.models.loss.WeightedLosses`` (external, unpinned in the reference's
``requirements.txt:2``; constructed at ``base_model.py:76`` and called at ``:102-103``): weights
``exp(-ln 2 * i)`` for forecast step ``i``, normalised to mean 1.  On the CUDA path the two weighted
losses are produced by the fused loss kernel (``ops.StepLossFn``); the methods below exist for API
compatibility and run on whatever device the inputs are on.
"""
from __future__ import annotations

import math
from typing import Optional

import torch


class WeightedLosses:
    def __init__(self, decay_rate: Optional[float] = None, forecast_length: int = 6):
        self.decay_rate = math.log(2) if decay_rate is None else decay_rate
        self.forecast_length = forecast_length
        w = torch.FloatTensor([math.exp(-self.decay_rate * i) for i in range(forecast_length)])
        self.weights = w / w.sum() * len(w)

    def get_mse_exp(self, output, target):
        return torch.mean(self.weights.to(output.device) * (output - target) ** 2)

    def get_mae_exp(self, output, target):
        return torch.mean(self.weights.to(output.device) * torch.abs(output - target))
