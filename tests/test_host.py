"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/pvb200.h declares,
the Model mirrors the reference interface (constructor, attributes, state_dict), the product refuses to run
on the CPU, and the plain-C oracle agrees with the torch oracle."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from oracle import conv3d_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g

    g.build()
    return True


def test_library_exports_every_declared_symbol(built):
    from predict_pv_yield_b200 import lib

    header = open(os.path.join(ROOT, "include", "pvb200.h")).read()
    declared = set(re.findall(r"\b(pvb200_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    L = lib.load()
    for name in sorted(declared):
        assert hasattr(L, name), f"libpvb200.so does not export {name}"
    assert declared == set(lib.SIGNATURES), declared ^ set(lib.SIGNATURES)
    assert L.pvb200_abi_version() == 1


def _c_kind(decl: str, named: bool = True) -> str:
    """C parameter (``named``) / return declaration -> the ctypes class it must be bound as."""
    d = re.sub(r"/\*.*?\*/", " ", decl).strip()
    if d.endswith("pvb200_stream_t") or re.search(r"\bpvb200_stream_t\s+\w+$", d):
        return "c_void_p"
    if "*" in d:
        if re.match(r"(const\s+)?char\s*\*$", d):
            return "c_char_p"
        if re.match(r"double\s*\*", d):
            return "LP_c_double"
        return "c_void_p"  # every other pointer (device or host) is passed as an address
    if named:
        d = re.sub(r"\s+\w+$", "", d)  # drop the parameter name
    d = re.sub(r"\bconst\b", "", d).strip()
    return {"int": "c_int", "long long": "c_longlong", "unsigned long long": "c_ulonglong", "size_t": "c_size_t",
            "float": "c_float", "double": "c_double", "void": "None"}[d]


def test_ctypes_signatures_match_the_header_prototypes():
    """ABI drift guard: argument count and C type of every prototype in include/pvb200.h == lib.SIGNATURES (a c_int bound
    where the header says long long, or a missing argument, corrupts the call silently)."""
    from predict_pv_yield_b200 import lib

    header = open(os.path.join(ROOT, "include", "pvb200.h")).read()
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    protos = re.findall(r"^\s*((?:const\s+)?(?:unsigned\s+)?[a-z_]+(?:\s+long)?(?:\s*\*)?)\s*(pvb200_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", header, flags=re.M)
    assert len(protos) == len(lib.SIGNATURES), (len(protos), len(lib.SIGNATURES))

    def name_of(t):
        if t is None:
            return "None"
        n = t.__name__
        return {"c_long": "c_longlong", "c_ulong": "c_ulonglong"}.get(n, n) if ctypes.sizeof(ctypes.c_long) == 8 else n

    for ret, name, args in protos:
        restype, argtypes = lib.SIGNATURES[name]
        assert name_of(restype) in (_c_kind(ret, False), {"c_size_t": "c_ulonglong"}.get(_c_kind(ret, False), "")), (name, ret, restype)
        params = [a.strip() for a in args.split(",")] if args.strip() not in ("", "void") else []
        assert len(params) == len(argtypes), (name, len(params), len(argtypes))
        for i, (a, t) in enumerate(zip(params, argtypes)):
            want, got = _c_kind(a), name_of(t)
            if want == "c_size_t" or got == "c_size_t":  # ctypes aliases c_size_t to c_ulong on LP64
                want, got = want.replace("c_size_t", "c_ulonglong"), got.replace("c_size_t", "c_ulonglong")
            if a.replace(" ", "").startswith("constpvb200_head_t*") or a.replace(" ", "").startswith("pvb200_head_t*"):
                assert got in ("c_void_p", "LP_Head"), (name, i, a, got)
                continue
            if want in ("c_void_p", "LP_c_double") and got.startswith("LP_"):  # typed host pointers (arrays of pointers, out-parameters)
                continue
            assert want == got, (name, i, a, got)


def test_head_struct_layout_matches_header(built, tmp_path):
    """sizeof/offsetof of the ctypes mirror == the C compiler's view of pvb200_head_t."""
    from predict_pv_yield_b200 import lib

    src = tmp_path / "probe.c"
    fields = [f for f, _ in lib.Head._fields_]
    body = "".join(f'printf("%zu\\n", offsetof(pvb200_head_t, {f}));' for f in fields)
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "pvb200.h"\n'
        f'int main(void){{printf("%zu\\n", sizeof(pvb200_head_t));{body}return 0;}}\n'
    )
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert out[0] == ctypes.sizeof(lib.Head)
    for f, off in zip(fields, out[1:]):
        assert getattr(lib.Head, f).offset == off, f


def test_compute_entry_points_fail_without_gpu(built):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from predict_pv_yield_b200 import lib

    L = lib.load()
    assert L.pvb200_sm_count() < 0
    assert b"CUDA" in L.pvb200_last_error() or b"device" in L.pvb200_last_error()
    # argument validation works without a device and reports through last_error
    rc = L.pvb200_sat_normalise_f32(None, None, None, None, 1, 1, 8, None)
    assert rc == 1 and b"null" in L.pvb200_last_error()


def test_model_mirrors_reference_interface():
    from predict_pv_yield_b200.models.conv3d.model import Model
    from predict_pv_yield_b200.utils import load_config

    # reference test_init: the production yaml builds (tests/models/conv3d/test_conv3d_model.py:10-15)
    m = Model(**load_config("configs/model/conv3d.yaml"))
    assert m.name == "conv3d" and Model.batch_size == 32
    assert m.cnn_output_size == 32 * 12 * 12 * 19 == 87552
    assert (m.forecast_len, m.history_len_30, m.number_of_samples_per_batch) == (4, 1, 32)
    # defaults of model.py:18-32
    d = Model()
    assert (d.include_pv_yield, d.include_nwp, d.forecast_minutes, d.history_minutes) == (True, True, 30, 60)
    assert d.number_of_conv3d_layers == 4 and d.fc3.in_features == 128 + 128 * 3 + 128


@pytest.mark.parametrize("kw", [
    dict(include_pv_yield=False, include_nwp=False, forecast_minutes=60, history_minutes=30),
    dict(include_pv_yield=True, include_nwp=True, forecast_minutes=60, history_minutes=30, image_size_pixels=16),
    dict(include_pv_yield=True, include_nwp=True, forecast_minutes=120, history_minutes=30, number_of_conv3d_layers=6,
         image_size_pixels=24, number_sat_channels=11, output_variable="gsp_yield"),
])
def test_state_dict_contract_equals_oracle(kw):
    from predict_pv_yield_b200.models.conv3d.model import Model

    torch.manual_seed(3)
    om = O.OracleModel(**kw)
    torch.manual_seed(3)
    m = Model(**kw)
    a, b = om.state_dict(), m.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype
        assert torch.equal(a[k], b[k]), k  # same construction order => same default init under a seed
    m.load_state_dict(a)  # reference-layout checkpoints load
    for attr in ("history_len_5", "forecast_len_5", "history_len_30", "forecast_len_30", "history_len_60",
                 "forecast_len_60", "forecast_len", "history_len", "number_of_samples_per_batch", "cnn_output_size"):
        assert getattr(m, attr) == getattr(om, attr), attr


def test_forward_on_cpu_fails_loudly():
    from predict_pv_yield_b200.models.conv3d.model import Model

    m = Model(include_pv_yield=False, include_nwp=False, forecast_minutes=60, history_minutes=60,
              image_size_pixels=16, number_sat_channels=11, fc1_output_features=16)
    b = O.make_synthetic_batch(2, 11, 25, 16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(b)
    assert isinstance(m.configure_optimizers(), torch.optim.Optimizer)


def test_weighted_losses_match_oracle_weights():
    from predict_pv_yield_b200.losses import WeightedLosses

    w = WeightedLosses(forecast_length=12).weights
    assert torch.equal(w, O.weighted_loss_weights(12))
    assert abs(float(w.mean()) - 1.0) < 1e-6 and abs(float(w[1] / w[0]) - 0.5) < 1e-6


def test_batchml_access_styles():
    from predict_pv_yield_b200.batch import BatchML, as_batch

    b = as_batch({"satellite": {"data": 1}, "nwp": 2})
    assert isinstance(b, BatchML) and b.satellite.data == 1 and b["nwp"] == 2 and "nwp" in b
    assert as_batch(b) is b
    with pytest.raises(KeyError):
        b["missing"]


# ---- plain-C oracle vs torch oracle -------------------------------------------------------------------
def _c_oracle():
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "libpv_oracle.so"))
    lib.ora_l1_loss.restype = ctypes.c_float
    return lib


def _fp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_c_oracle_agrees_with_torch_oracle(built):
    lib = _c_oracle()
    rs = np.random.RandomState(0)
    B, Ci, T, H, W, Co = 2, 5, 4, 7, 6, 3
    xi = rs.randint(-1, 1024, size=(B, Ci, T, H, W)).astype(np.int16)
    mean, std = O.sat_constants(12)
    mean, std = mean[:Ci].copy(), std[:Ci].copy()
    xn = np.empty(xi.shape, np.float32)
    lib.ora_sat_normalise(_fp(xi), _fp(xn), _fp(mean), _fp(std), B, Ci, ctypes.c_long(T * H * W))
    assert np.array_equal(xn.view(np.uint32), O.sat_normalise_numpy(xi, mean, std).view(np.uint32))
    w = (rs.randn(Co, Ci, 3, 3, 3) / 10).astype(np.float32)
    b = rs.randn(Co).astype(np.float32)
    y = np.empty((B, Co, T - 2, H - 2, W - 2), np.float32)
    lib.ora_conv3d_relu(_fp(xn), _fp(w), _fp(b), _fp(y), B, Ci, T, H, W, Co, 1)
    want = torch.relu(torch.nn.functional.conv3d(torch.from_numpy(xn).double(), torch.from_numpy(w).double(),
                                                 torch.from_numpy(b).double())).float().numpy()
    assert np.abs(y - want).max() <= 1e-6 * np.abs(want).max()
    # flatten order + Linear layout
    feats = y.reshape(B, -1)
    wl = (rs.randn(4, feats.shape[1]) / 10).astype(np.float32)
    bl = rs.randn(4).astype(np.float32)
    out = np.empty((B, 4), np.float32)
    lib.ora_linear(_fp(feats), _fp(wl), _fp(bl), _fp(out), B, ctypes.c_long(feats.shape[1]), 4, 0)
    want = torch.nn.functional.linear(torch.from_numpy(want).reshape(B, -1).double(), torch.from_numpy(wl).double(),
                                      torch.from_numpy(bl).double()).float().numpy()
    assert np.abs(out - want).max() <= 1e-5 * np.abs(want).max()
    t = rs.rand(B, 4).astype(np.float32)
    l1 = lib.ora_l1_loss(_fp(out), _fp(t), ctypes.c_long(out.size))
    assert abs(l1 - np.abs(out - t).mean()) <= 1e-6


def test_device_prefetcher_rejects_cpu():
    from predict_pv_yield_b200.data import DevicePrefetcher

    with pytest.raises(RuntimeError, match="CUDA"):
        DevicePrefetcher([], torch.device("cpu"))


def test_split_precision_shape_rules(built):
    """Which shapes the two-way fp16 split of the implicit GEMM takes (CTA-pair kernel, whole 16-channel steps; the rest
    runs 3xTF32), and the blocked-layout group counts behind the rule -- host logic, no device needed."""
    from predict_pv_yield_b200 import ops

    assert [ops.blocked4_groups(c) for c in (1, 8, 9, 12, 16, 24, 32)] == [2, 2, 4, 4, 4, 6, 8]
    assert ops.conv_f16x2_applies(12, 32) and ops.conv_f16x2_applies(32, 32) and ops.conv_f16x2_applies(16, 24)
    assert not ops.conv_f16x2_applies(8, 32)     # half a 16-channel step
    assert not ops.conv_f16x2_applies(24, 32)    # one and a half
    assert not ops.conv_f16x2_applies(32, 16)    # single-CTA kernel (Cout <= 16)
    assert ops.tf32x3_supported(32, 32) and not ops.tf32x3_supported(33, 32) and not ops.tf32x3_supported(32, 64)


def test_invalidate_shadow_forces_a_rebuild_of_the_bf16_copy():
    """writes through ``fc1.weight.data`` do not bump the parameter's version counter, so the staleness key
    of the tensor-core shadow cannot see them -- ``Model.invalidate_shadow()`` is the explicit switch (both models)."""
    from predict_pv_yield_b200.models.conv3d.model import Model
    from predict_pv_yield_b200.models.conv3d.model_sat_nwp import Model as SatNwp
    from oracle.golden_cases import CASES, SAT_NWP_CASES

    m = Model(**CASES["test_yaml_pv"]["model"], precision="bf16")
    w = m.fc1.weight
    assert w._pvb_shadow is m._fc1_shadow
    key = m._fc1_shadow._key(w)
    m._fc1_shadow.key = key
    w.data.mul_(0.5)  # invisible to the key ...
    assert m._fc1_shadow._key(w) == key
    with torch.no_grad():
        w.mul_(0.5)  # ... while an in-place edit of the parameter itself is not
    assert m._fc1_shadow._key(w) != key
    m._fc1_shadow.key = m._fc1_shadow._key(w)
    m.invalidate_shadow()
    assert m._fc1_shadow.key is None
    s = SatNwp(**SAT_NWP_CASES["sat_nwp_pv"]["model"], precision="bf16")
    s._fc1_shadow.key = s._nwp_fc1_shadow.key = ("x",)
    s.invalidate_shadow()
    assert s._fc1_shadow.key is None and s._nwp_fc1_shadow.key is None


def test_optimizer_checkpoints_move_between_fused_adam_and_torch_adam():
    """Same state names and group keys as ``torch.optim.Adam`` (base_model.py:255-257): a FusedAdam checkpoint resumes under
    torch's optimiser (and continues exactly like a torch run with the same moments), a torch checkpoint loads into
    FusedAdam, and settings the kernel does not implement are refused instead of ignored."""
    from predict_pv_yield_b200.optim import FusedAdam

    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(4, 3)), torch.nn.Parameter(torch.randn(5))]
    fused = FusedAdam(params, lr=5e-4)
    for q in params:  # the state FusedAdam.step() keeps: int step, two moments
        fused.state[q] = {"step": 3, "exp_avg": torch.randn_like(q) * 0.1, "exp_avg_sq": torch.rand_like(q) * 0.01}
    import copy

    sd = copy.deepcopy(fused.state_dict())  # torch's load_state_dict may alias the tensors it is handed
    grads = [torch.randn_like(q) for q in params]

    def run(opt_params, loader):
        opt = torch.optim.Adam(opt_params, lr=1.0)  # every hyper-parameter must come from the checkpoint
        loader(opt)
        for q, g in zip(opt_params, grads):
            q.grad = g.clone()
        opt.step()
        return opt

    a_params = [torch.nn.Parameter(q.detach().clone()) for q in params]
    a = run(a_params, lambda opt: opt.load_state_dict(sd))
    assert a.param_groups[0]["lr"] == 5e-4 and float(a.state[a_params[0]]["step"]) == 4.0
    b_params = [torch.nn.Parameter(q.detach().clone()) for q in params]

    def seed_state(opt):
        opt.param_groups[0]["lr"] = 5e-4
        for q, src in zip(b_params, params):
            opt.state[q] = {"step": torch.tensor(3.0), "exp_avg": fused.state[src]["exp_avg"].clone(),
                            "exp_avg_sq": fused.state[src]["exp_avg_sq"].clone()}

    run(b_params, seed_state)
    for x, y in zip(a_params, b_params):
        assert torch.equal(x, y)
    # torch -> FusedAdam
    f2 = FusedAdam([torch.nn.Parameter(q.detach().clone()) for q in params], lr=1.0)
    f2.load_state_dict(a.state_dict())
    assert f2.param_groups[0]["lr"] == 5e-4 and int(f2.state[f2.param_groups[0]["params"][0]]["step"]) == 4
    # refused, not ignored
    f2.param_groups[0]["weight_decay"] = 0.01
    with pytest.raises(RuntimeError, match="weight_decay"):
        f2.step()
