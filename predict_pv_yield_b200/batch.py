"""Batch container on the hot path.

The reference wraps incoming dicts in ``nowcasting_dataloader.batch.BatchML`` (a pydantic model,
``models/conv3d/model.py:109-110``, ``base_model.py:84-85``) and then uses BOTH attribute access
(``x.satellite.data``, ``batch.pv.pv_yield``) and item access (``x["nwp"]``, ``x[self.output_variable]``).
``nowcasting_dataloader`` is an external package; this light stand-in offers exactly those two access
styles over a nested dict so the same batch dicts drop in.  A real ``BatchML`` instance is also accepted
anywhere a batch is expected (duck typing).
"""
from __future__ import annotations

from typing import Any


class BatchML:
    def __init__(self, **kwargs: Any):
        for k, v in kwargs.items():
            setattr(self, k, BatchML(**v) if isinstance(v, dict) else v)

    def __getitem__(self, key: str) -> Any:
        try:
            return getattr(self, key)
        except AttributeError as e:  # mirror dict-style failure
            raise KeyError(key) from e

    def __contains__(self, key: str) -> bool:
        return hasattr(self, key)

    def keys(self):
        return self.__dict__.keys()


def as_batch(x: Any) -> Any:
    """dict -> BatchML (reference: ``if type(x) == dict: x = BatchML(**x)``); anything else passes through."""
    return BatchML(**x) if isinstance(x, dict) else x
