"""``load_config`` with the semantics of the reference's ``predict_pv_yield/utils.py:16-32``:
read a model yaml relative to the repository root and drop the Hydra ``_target_`` key, so that
``Model(**load_config("configs/model/conv3d.yaml"))`` works as in the reference's tests
(``tests/models/conv3d/test_conv3d_model.py:12-15``)."""
from __future__ import annotations

import os

import yaml

import predict_pv_yield_b200


def load_config(config_file: str) -> dict:
    path = os.path.dirname(predict_pv_yield_b200.__file__)
    full = config_file if os.path.isabs(config_file) else f"{path}/../{config_file}"
    with open(full, "r") as cfg:
        config = yaml.load(cfg, Loader=yaml.FullLoader)
    if "_target_" in config.keys():
        config.pop("_target_")  # this is only for Hydra
    return config
