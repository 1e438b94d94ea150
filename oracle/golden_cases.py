"""TEST INFRASTRUCTURE ONLY -- the seeded cases behind ``tests/golden/*.npz``.

Shared by ``oracle/make_golden.py`` (runs the real reference on them, build container only) and by
``tests/`` (runs the oracle and the CUDA path on the same inputs).  Inputs and weights come from
numpy's legacy ``RandomState`` (bit-stable across numpy versions), NOT from torch's RNG, so the
fixtures stay valid if the torch version on the GPU box differs.
"""
from __future__ import annotations

import numpy as np
import torch

# name -> (model kwargs, batch size B, n yield timesteps)
CASES = {
    # the reference's own test yaml (tests/configs/model/conv3d.yaml): 11 ch, 16x16, hist 60 / fcst 60 => T=25
    "test_yaml_pv": dict(
        model=dict(include_pv_yield=False, include_nwp=False, forecast_minutes=60, history_minutes=60,
                   number_of_conv3d_layers=4, conv3d_channels=32, image_size_pixels=16, number_sat_channels=11,
                   fc1_output_features=16, fc2_output_features=16, fc3_output_features=16),
        batch=2,
    ),
    # tests/configs/model/conv3d_gsp.yaml
    "test_yaml_gsp": dict(
        model=dict(include_pv_yield=False, include_nwp=False, forecast_minutes=60, history_minutes=60,
                   number_of_conv3d_layers=4, conv3d_channels=32, image_size_pixels=16, number_sat_channels=11,
                   fc1_output_features=16, fc2_output_features=16, fc3_output_features=16,
                   output_variable="gsp_yield"),
        batch=2,
    ),
    # BASELINE config 3 in miniature: NWP + PV-history branches on (model.py:129-148), 12 ch, T=19
    "nwp_pv_small": dict(
        model=dict(include_pv_yield=True, include_nwp=True, forecast_minutes=60, history_minutes=30,
                   number_of_conv3d_layers=4, conv3d_channels=32, image_size_pixels=16, number_sat_channels=12,
                   fc1_output_features=128, fc2_output_features=128, fc3_output_features=64),
        batch=3,
    ),
    # production yaml in miniature (configs/model/conv3d.yaml: 6 layers, gsp_yield, hist 30 / fcst 120 => T=31)
    "prod_yaml_small": dict(
        model=dict(include_pv_yield=True, include_nwp=True, forecast_minutes=120, history_minutes=30,
                   number_of_conv3d_layers=6, conv3d_channels=32, image_size_pixels=16, number_sat_channels=11,
                   fc1_output_features=128, fc2_output_features=128, fc3_output_features=64,
                   output_variable="gsp_yield"),
        batch=2,
    ),
    # edges none of the yaml-shaped cases reach: ONLY the PV-history branch (fc3 in = 384), 16 channels (the fp32 mode's
    # 3xTF32 kernels instead of the fp16 split), an odd crop (17 -> 13x13 planes) and an odd number of positions per
    # channel (9 * 13 * 13 = 1521: the scalar tail of the fused Adam + shadow kernel), two layers, hist 30 / fcst 30 => T=13
    "pv_only_odd": dict(
        model=dict(include_pv_yield=True, include_nwp=False, forecast_minutes=30, history_minutes=30,
                   number_of_conv3d_layers=2, conv3d_channels=16, image_size_pixels=17, number_sat_channels=12,
                   fc1_output_features=128, fc2_output_features=128, fc3_output_features=64),
        batch=3,
    ),
    # ONLY the NWP branch (fc3 in = 256) behind a ONE-layer encoder (model.py:86-90: the loop over conv3d_1.. is empty):
    # no data gradient at all, sat_conv0 feeds fc1 directly; 11 channels, 12x12 crops
    "nwp_only_one_layer": dict(
        model=dict(include_pv_yield=False, include_nwp=True, forecast_minutes=60, history_minutes=30,
                   number_of_conv3d_layers=1, conv3d_channels=32, image_size_pixels=12, number_sat_channels=11,
                   fc1_output_features=64, fc2_output_features=32, fc3_output_features=16),
        batch=2,
    ),
}

# ---- the two-tower model (predict_pv_yield/models/conv3d/model_sat_nwp.py) ---------------------
SAT_NWP_CASES = {
    # the reference's test yaml (tests/configs/model/conv3d_sat_nwp.yaml) shape family: small crops, both towers,
    # PV history, embedding
    "sat_nwp_pv": dict(
        model=dict(include_pv_or_gsp_yield_history=True, include_nwp=True, forecast_minutes=60, history_minutes=60,
                   number_of_conv3d_layers=3, conv3d_channels=16, image_size_pixels=12, nwp_image_size_pixels=10,
                   number_sat_channels=11, number_nwp_channels=10, fc1_output_features=32, fc2_output_features=32,
                   fc3_output_features=16, output_variable="pv_yield", embedding_dem=16, include_pv_yield_history=True,
                   include_future_satellite=True),
        batch=3,
    ),
    # gsp target, no future satellite, no PV-history layer (the production yaml's switches, configs/model/conv3d_sat_nwp.yaml)
    "sat_nwp_gsp": dict(
        model=dict(include_pv_or_gsp_yield_history=False, include_nwp=True, forecast_minutes=120, history_minutes=30,
                   number_of_conv3d_layers=2, conv3d_channels=32, image_size_pixels=10, nwp_image_size_pixels=8,
                   number_sat_channels=12, number_nwp_channels=10, fc1_output_features=128, fc2_output_features=128,
                   fc3_output_features=64, output_variable="gsp_yield", embedding_dem=16, include_pv_yield_history=False,
                   include_future_satellite=False),
        batch=2,
    ),
}


def sat_nwp_batch(name: str, seed: int = 519) -> dict:
    """Deterministic nested batch dict for SAT_NWP_CASES[name] (int16 satellite, NaNs in the first history row)."""
    case = SAT_NWP_CASES[name]
    kw, B = case["model"], case["batch"]
    rs = np.random.RandomState(seed)
    C, T, S = kw["number_sat_channels"], seq_len_of(kw), kw["image_size_pixels"]
    sat = rs.randint(0, 1024, size=(B, C, T, S, S)).astype(np.int16)
    sat[rs.rand(*sat.shape) < 2e-3] = -1
    t_nwp = kw["forecast_minutes"] // 60 + int(np.ceil(kw["history_minutes"] / 60)) + 1
    nwp = rs.randn(B, kw["number_nwp_channels"], t_nwp, kw["nwp_image_size_pixels"], kw["nwp_image_size_pixels"]).astype(np.float32)
    pv = rs.rand(B, T, 128).astype(np.float32)
    pv[:, 0, :][rs.rand(B, 128) < 0.05] = np.nan
    pv[:, -kw["forecast_minutes"] // 5:, 0] = rs.rand(B, kw["forecast_minutes"] // 5).astype(np.float32)  # finite targets
    n30 = kw["history_minutes"] // 30 + 1 + kw["forecast_minutes"] // 30
    gsp = rs.rand(B, n30, 32).astype(np.float32)
    b = {
        "satellite": {"data": torch.from_numpy(sat)},
        "nwp": {"data": torch.from_numpy(nwp)},
        "pv": {"pv_yield": torch.from_numpy(pv), "pv_system_row_number": torch.from_numpy(rs.randint(0, 940, size=(B, 128)).astype(np.int64))},
        "gsp": {"gsp_yield": torch.from_numpy(gsp), "gsp_id": torch.from_numpy(rs.randint(0, 338, size=(B, 32)).astype(np.int64))},
    }
    return b


# ---- Conv3dMaxPool (perceiver_conv3d_nwp_sat.py:42-57) -------------------------------------------
MAXPOOL_CASES = {
    # name -> (B, in_channels, T, H, W, out_channels); odd and even extents (the pool's last window is partial for even ones)
    "conv3d_maxpool_sat": (2, 11, 5, 12, 12, 16),
    "conv3d_maxpool_odd": (1, 10, 3, 9, 7, 8),
}


def maxpool_inputs(name: str, seed: int = 520):
    """(x fp32 [B,Ci,T,H,W], upstream gradient g for the output) from numpy RandomState."""
    B, Ci, T, H, W, Co = MAXPOOL_CASES[name]
    rs = np.random.RandomState(seed)
    x = rs.randn(B, Ci, T, H, W).astype(np.float32)
    g = rs.randn(B, Co, T, (H - 1) // 2 + 1, (W - 1) // 2 + 1).astype(np.float32)
    return torch.from_numpy(x), torch.from_numpy(g)


SUBSAMPLE = 97  # stride used to thin out large tensors (fc1.weight and its grad) in the fixtures


def seq_len_of(kw: dict) -> int:
    return kw["forecast_minutes"] // 5 + kw["history_minutes"] // 5 + 1


def golden_batch(name: str, seed: int = 518) -> dict:
    """Deterministic nested batch dict for CASES[name] (int16 satellite, NaNs in PV history row 0)."""
    case = CASES[name]
    kw, B = case["model"], case["batch"]
    rs = np.random.RandomState(seed)
    C, T, S = kw["number_sat_channels"], seq_len_of(kw), kw["image_size_pixels"]
    sat = rs.randint(0, 1024, size=(B, C, T, S, S)).astype(np.int16)
    sat[rs.rand(*sat.shape) < 2e-3] = -1
    var = kw.get("output_variable", "pv_yield")
    n_sys = 128 if var == "pv_yield" else 32
    # yield time axis: 5-min steps for pv (T), 30-min steps for gsp (history_len_30 + 1 + forecast_len_30)
    n_t = T if var == "pv_yield" else (kw["history_minutes"] // 30 + 1 + kw["forecast_minutes"] // 30)
    yld = rs.rand(B, n_t, n_sys).astype(np.float32)
    legacy = yld.copy()
    legacy[:, 0, :][rs.rand(B, n_sys) < 0.05] = np.nan
    nwp = rs.randn(B, 10, 19, 2, 2).astype(np.float32)
    b = {"satellite": {"data": torch.from_numpy(sat)}, "nwp": torch.from_numpy(nwp), var: torch.from_numpy(legacy)}
    b["pv" if var == "pv_yield" else "gsp"] = {var: torch.from_numpy(yld)}
    return b


def golden_state_dict(model: torch.nn.Module, seed: int = 1234) -> dict:
    """Deterministic weights: U(-1/sqrt(fan_in), +1/sqrt(fan_in)) from numpy RandomState, in
    ``state_dict`` key order (same bound as torch's default init, probe in SURVEY.md section 8b)."""
    rs = np.random.RandomState(seed)
    sd = {}
    for k, v in model.state_dict().items():
        shape = tuple(v.shape)
        if k.endswith(".weight"):
            fan_in = int(np.prod(shape[1:]))
        else:
            w = model.state_dict()[k[: -len("bias")] + "weight"]
            fan_in = int(np.prod(tuple(w.shape)[1:]))
        bound = 1.0 / np.sqrt(fan_in)
        sd[k] = torch.from_numpy(rs.uniform(-bound, bound, size=shape).astype(np.float32))
    return sd


def thin(t: torch.Tensor) -> np.ndarray:
    """Flatten; keep everything for small tensors (<= 4096 elements), every SUBSAMPLE-th element for big ones."""
    a = t.detach().cpu().reshape(-1).numpy()
    return a if a.size <= 4096 else a[::SUBSAMPLE].copy()
