"""Validation results of the step: ``predict_pv_yield/models/base_model.py:222-250``.

The reference scales forecast and truth by the GSP capacity on the host (four ``.cpu().numpy()`` round trips per batch),
builds a long-format table with ``nowcasting_utils.metrics.validation.make_validation_results`` and writes one CSV per
epoch with ``save_validation_results_to_logger``.  Here the scaling (and the per-horizon error metrics of
``base_model.py:121-136``) run in one kernel (``pvb200_validation_results_f32``) and reach the host in ONE copy; the table
has the columns the reference's own test asserts (``tests/models/baseline/test_baseline_model_gsp.py:103-111``):
``t0_datetime_utc, target_datetime_utc, gsp_id, actual_gsp_pv_outturn_mw, forecast_gsp_pv_outturn_mw`` (+ ``capacity_mwp``,
``batch_index``, ``example_index``), one row per (example, 30-minute horizon), ``len == B * forecast_len_30``.
``nowcasting_utils`` is external and unpinned (``requirements.txt:2``): the schema beyond those assertions is a
restatement of its published layout, not pinned against it.  pandas is needed only here (never on the training path).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch


def make_validation_results(truths_mw: np.ndarray, predictions_mw: np.ndarray, capacity_mwp: np.ndarray, gsp_ids,
                            t0_datetimes_utc, batch_idx: Optional[int] = None, forecast_minutes_per_step: int = 30):
    """Long-format DataFrame: one row per (example, forecast horizon)."""
    import pandas as pd

    truths_mw, predictions_mw, capacity_mwp = (np.asarray(a, dtype=np.float32) for a in (truths_mw, predictions_mw, capacity_mwp))
    if truths_mw.shape != predictions_mw.shape or truths_mw.ndim != 2:
        raise ValueError(f"validation results: truths {truths_mw.shape} and predictions {predictions_mw.shape} must be [B, horizons]")
    B, F = truths_mw.shape
    gsp_ids = np.asarray(torch.as_tensor(gsp_ids).cpu() if torch.is_tensor(gsp_ids) else gsp_ids).reshape(-1)
    t0 = pd.to_datetime(pd.Series(np.asarray(t0_datetimes_utc).reshape(-1)))
    if len(gsp_ids) != B or len(t0) != B:
        raise ValueError("validation results: gsp_ids / t0_datetimes_utc must have one entry per example")
    frames = []
    for i in range(F):
        frames.append(pd.DataFrame({
            "t0_datetime_utc": t0.values,
            "target_datetime_utc": (t0 + pd.Timedelta(minutes=forecast_minutes_per_step * (i + 1))).values,
            "gsp_id": gsp_ids,
            "actual_gsp_pv_outturn_mw": truths_mw[:, i],
            "forecast_gsp_pv_outturn_mw": predictions_mw[:, i],
            "capacity_mwp": capacity_mwp[:, i],
            "batch_index": batch_idx,
            "example_index": np.arange(B),
        }))
    return pd.concat(frames, ignore_index=True)


def save_validation_results(results_dfs, results_file_name: str, current_epoch: int) -> Optional[str]:
    """One CSV per epoch: ``{results_file_name}_{current_epoch}.csv`` (what the reference's test reads back)."""
    import pandas as pd

    if not results_dfs:
        return None
    path = f"{results_file_name}_{current_epoch}.csv"
    pd.concat(results_dfs, ignore_index=True).to_csv(path, index=False)
    return path
