"""validation results (predict_pv_yield/models/base_model.py:121-136,222-250)."""
import os

import numpy as np
import pytest
import torch

from oracle import conv3d_oracle as O
from oracle.golden_cases import CASES, golden_batch, golden_state_dict


def test_make_validation_results_schema():
    """The columns and the row count the reference's own test asserts (tests/models/baseline/test_baseline_model_gsp.py:103-111)."""
    pd = pytest.importorskip("pandas")
    from predict_pv_yield_b200.validation import make_validation_results, save_validation_results

    B, F = 3, 4
    rs = np.random.RandomState(0)
    truths, preds, cap = rs.rand(B, F), rs.rand(B, F), 100 * rs.rand(B, F)
    t0 = np.array(["2021-06-01T12:00", "2021-06-01T12:30", "2021-06-02T09:00"], dtype="datetime64[ns]")
    df = make_validation_results(truths, preds, cap, gsp_ids=torch.tensor([7, 8, 9]), t0_datetimes_utc=t0, batch_idx=2)
    assert len(df) == B * F
    for col in ("t0_datetime_utc", "target_datetime_utc", "gsp_id", "actual_gsp_pv_outturn_mw", "forecast_gsp_pv_outturn_mw"):
        assert col in df.keys()
    row = df[(df.example_index == 1) & (df.target_datetime_utc == pd.Timestamp("2021-06-01T14:00"))].iloc[0]  # horizon 3 of example 1
    assert row.gsp_id == 8 and np.isclose(row.forecast_gsp_pv_outturn_mw, preds[1, 2]) and np.isclose(row.capacity_mwp, cap[1, 2])
    assert save_validation_results([], "unused", 0) is None


@pytest.mark.gpu
def test_validation_step_builds_the_results_table_on_the_device(tmp_path):
    pd = pytest.importorskip("pandas")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from predict_pv_yield_b200 import ops
    from predict_pv_yield_b200.models.conv3d.model import Model

    dev = torch.device("cuda:0")
    name = "prod_yaml_small"  # gsp_yield, forecast 120 min => 4 half-hour horizons
    case = CASES[name]
    m = Model(**case["model"]).to(dev)
    B = case["batch"]
    m.batch_size = B
    m.load_state_dict(golden_state_dict(m))
    batch = golden_batch(name)
    rs = np.random.RandomState(3)
    n_t = batch["gsp"]["gsp_yield"].shape[1]
    batch["gsp"]["gsp_capacity"] = torch.from_numpy((50 + 100 * rs.rand(B, n_t, 32)).astype(np.float32))
    batch["gsp"]["gsp_id"] = torch.from_numpy(rs.randint(1, 338, size=(B, 32)).astype(np.int64))
    batch["metadata"] = {"t0_datetime_utc": np.array(["2021-06-01T12:00", "2021-06-01T12:30"], dtype="datetime64[ns]")}
    dbatch = O.batch_to(batch, dev)
    m.results_file_name = str(tmp_path / "results_epoch")
    loss = m.validation_step(dbatch, 0)
    with torch.no_grad():
        y_hat = m(dbatch).cpu().numpy()
    y = batch["gsp"]["gsp_yield"][:, -4:, 0].numpy()
    cap = batch["gsp"]["gsp_capacity"][:, -4:, 0].numpy()
    assert abs(float(loss.detach()) - float(np.abs(y_hat - y).mean())) <= 1e-6
    df = m.results_dfs[0]
    assert len(df) == B * m.forecast_len_30
    got_f = df.sort_values(["example_index", "target_datetime_utc"]).forecast_gsp_pv_outturn_mw.to_numpy().reshape(B, 4)
    got_a = df.sort_values(["example_index", "target_datetime_utc"]).actual_gsp_pv_outturn_mw.to_numpy().reshape(B, 4)
    assert np.allclose(got_f, y_hat * cap, rtol=1e-6) and np.allclose(got_a, y * cap, rtol=1e-6)
    assert list(df[df.example_index == 0].gsp_id.unique()) == [int(batch["gsp"]["gsp_id"][0, 0])]
    # per-horizon metrics come from the same kernel (the reference's MAE-over-MSE key collision is kept)
    logged = m.logged_metrics
    assert abs(float(logged["MSE_forecast_horizon_1/Validation"]) - float(np.abs(y_hat - y).mean(0)[1])) <= 1e-6
    out, horizon = ops.validation_results(torch.from_numpy(y_hat).to(dev), dbatch["gsp"]["gsp_yield"][:, -4:, 0], None)
    assert np.allclose(horizon[0].cpu().numpy(), ((y_hat - y) ** 2).mean(0), rtol=1e-5)
    assert np.allclose(out[2].cpu().numpy(), 1.0)
    m.validation_epoch_end([])
    back = pd.read_csv(f"{m.results_file_name}_0.csv")
    assert len(back) == B * m.forecast_len_30 and "forecast_gsp_pv_outturn_mw" in back.keys()
