"""Host side of the hot path: torch tensors in, C-ABI calls (``lib.py``) on the current CUDA stream.

PyTorch is plumbing here (device memory, streams, autograd bookkeeping); every FLOP of the step runs
in ``libpvb200.so``.  Nothing in this file computes on the CPU and nothing falls back to torch
operators: a CPU tensor or a missing library raises.

Autograd nodes (private protocol between them is documented on each class):
  ``EncoderFn``  sat cube -> flattened last conv activation   (model.py:113-122 of the reference)
  ``HeadFn``     features (+PV history, +NWP) -> forecast     (model.py:125-154)
  ``StepLossFn`` forecast, target -> [nmae, mse, mse_exp, mae_exp]  (base_model.py:95-103)
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import lib as _lib

# ---------------------------------------------------------------------------------------------
# plumbing
# ---------------------------------------------------------------------------------------------
_workspaces: Dict[Tuple[int, str], torch.Tensor] = {}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need_cuda(t: torch.Tensor, name: str, dtype=None) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"predict_pv_yield_b200: '{name}' is on {t.device}; this implementation is CUDA (sm_100a) only "
            "and has no CPU fallback"
        )
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"predict_pv_yield_b200: '{name}' must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"predict_pv_yield_b200: '{name}' must be contiguous")


def _workspace(name: str, nbytes: int, device: torch.device) -> torch.Tensor:
    """Per-device grow-only scratch buffers (single compute stream => stream-ordered reuse is safe)."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), name)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


class KernelTimer:
    """Optional per-call device timing (CUDA events on the launching stream) with the algorithmic FLOPs and
    bytes of each call, used by bench.py for the roofline numbers.  Off unless ``set_timer`` is called."""

    def __init__(self):
        self.records = []  # (name, flops, nbytes, start_event, end_event)

    def summary(self) -> Dict[str, dict]:
        """Call after a device synchronize."""
        out: Dict[str, dict] = {}
        for name, flops, nbytes, e0, e1 in self.records:
            d = out.setdefault(name, dict(calls=0, ms=0.0, flops=0.0, bytes=0.0))
            d["calls"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["flops"] += flops
            d["bytes"] += nbytes
        return out


_timer: Optional[KernelTimer] = None


def set_timer(t: Optional[KernelTimer]) -> None:
    global _timer
    _timer = t


class _timed:
    def __init__(self, name: str, flops: float = 0.0, nbytes: float = 0.0):
        self.args = (name, flops, nbytes)

    def __enter__(self):
        if _timer is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if _timer is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _timer.records.append((*self.args, self.e0, e1))
        return False


# ---------------------------------------------------------------------------------------------
# thin operator wrappers (one C-ABI call each)
# ---------------------------------------------------------------------------------------------
def sat_normalise(x: torch.Tensor, mean: torch.Tensor, std: torch.Tensor, out_dtype=torch.float32) -> torch.Tensor:
    """(float32(x) - mean[c]) / std[c], bit-identical to netcdf_dataset.py:96-101.  x: int16 [B,C,T,H,W]."""
    L = _lib.load()
    _need_cuda(x, "satellite.data", torch.int16)
    _need_cuda(mean, "sat_mean", torch.float32)
    _need_cuda(std, "sat_std", torch.float32)
    B, Cc = x.shape[0], x.shape[1]
    thw = x[0, 0].numel()
    if mean.numel() != Cc or std.numel() != Cc:
        raise RuntimeError(f"sat_normalise: {Cc} channels but {mean.numel()} means / {std.numel()} stds")
    y = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    if out_dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError(f"sat_normalise: unsupported output dtype {out_dtype}")
    with _timed("sat_normalise", 0.0, x.numel() * (2 + y.element_size())):
        if out_dtype == torch.float32:
            rc = L.pvb200_sat_normalise_f32(_p(x), _p(y), _p(mean), _p(std), B, Cc, thw, _stream())
        else:
            rc = L.pvb200_sat_normalise_bf16(_p(x), _p(y), _p(mean), _p(std), B, Cc, thw, _stream())
    _lib.check(rc, "sat_normalise")
    return y


def conv3d_fwd(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], relu: bool = True,
               mean: Optional[torch.Tensor] = None, std: Optional[torch.Tensor] = None, pad_t: int = 0,
               pad_hw: int = 0) -> torch.Tensor:
    """relu(conv3d(x, w, b)) with 3x3x3 kernel, padding (pad_t, 0, 0) (model.py:117-120; pad_t = 1: model_sat_nwp.py:85-100).
    x fp32, or int16 with fused normalise."""
    L = _lib.load()
    i16 = x.dtype == torch.int16
    _need_cuda(x, "conv input", torch.int16 if i16 else torch.float32)
    _need_cuda(w, "conv weight", torch.float32)
    if b is not None:
        _need_cuda(b, "conv bias", torch.float32)
    B, Ci, Ti, Hi, Wi = x.shape
    Co = w.shape[0]
    if tuple(w.shape) != (Co, Ci, 3, 3, 3):
        raise RuntimeError(f"conv3d: weight shape {tuple(w.shape)} does not match input channels {Ci}")
    if min(Ti + 2 * pad_t, Hi + 2 * pad_hw, Wi + 2 * pad_hw) < 3:
        raise RuntimeError(f"conv3d: input {Ti}x{Hi}x{Wi} smaller than the 3x3x3 kernel")
    if i16 and (mean is None or std is None):
        raise RuntimeError("conv3d: int16 input needs mean/std")
    y = torch.empty((B, Co, Ti + 2 * pad_t - 2, Hi + 2 * pad_hw - 2, Wi + 2 * pad_hw - 2), dtype=torch.float32, device=x.device)
    nb = L.pvb200_conv3d_workspace_bytes(Ci, Co)
    ws = _workspace("conv", nb, x.device)
    with _timed(f"conv3d_fwd_f32[Ci={Ci}]", 2.0 * 27 * Ci * y.numel(), x.numel() * x.element_size() + 4.0 * y.numel()):
        rc = L.pvb200_conv3d_fwd_f32_pad(_p(x), int(i16), _p(mean) if i16 else None, _p(std) if i16 else None, _p(w), _p(b),
                                         _p(y), _p(ws), ws.numel(), B, Ci, Ti, Hi, Wi, Co, int(relu), pad_t, pad_hw, _stream())
    _lib.check(rc, "conv3d_fwd")
    return y


def conv3d_dgrad(gz: torch.Tensor, w: torch.Tensor, mask_src: Optional[torch.Tensor], x_shape: Sequence[int],
                 pad_t: int = 0, pad_hw: int = 0) -> torch.Tensor:
    """gx = conv_transpose3d(gz, w) * (mask_src > 0).  gz: gradient w.r.t. the conv's pre-activation output."""
    L = _lib.load()
    _need_cuda(gz, "gz", torch.float32)
    _need_cuda(w, "conv weight", torch.float32)
    B, Ci, Ti, Hi, Wi = x_shape
    Co = w.shape[0]
    if tuple(gz.shape) != (B, Co, Ti + 2 * pad_t - 2, Hi + 2 * pad_hw - 2, Wi + 2 * pad_hw - 2):
        raise RuntimeError(f"conv3d_dgrad: gz shape {tuple(gz.shape)} inconsistent with input {tuple(x_shape)}")
    if mask_src is not None:
        _need_cuda(mask_src, "mask_src", torch.float32)
        if tuple(mask_src.shape) != tuple(x_shape):
            raise RuntimeError("conv3d_dgrad: mask_src shape mismatch")
    gx = torch.empty(tuple(x_shape), dtype=torch.float32, device=gz.device)
    nb = L.pvb200_conv3d_workspace_bytes(Ci, Co)
    ws = _workspace("conv", nb, gz.device)
    with _timed(f"conv3d_dgrad_f32[Ci={Ci}]", 2.0 * 27 * Ci * gz.numel(),
                4.0 * (gz.numel() + gx.numel() * (2 if mask_src is not None else 1))):
        rc = L.pvb200_conv3d_dgrad_f32_pad(_p(gz), _p(w), _p(mask_src), _p(gx), _p(ws), ws.numel(), B, Ci, Ti, Hi, Wi, Co, pad_t,
                                           pad_hw, _stream())
    _lib.check(rc, "conv3d_dgrad")
    return gx


def conv3d_wgrad(x: torch.Tensor, gz: torch.Tensor, mean: Optional[torch.Tensor] = None,
                 std: Optional[torch.Tensor] = None, pad_t: int = 0, pad_hw: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """(dw [Co,Ci,3,3,3], db [Co]) from the layer input x and the pre-activation gradient gz."""
    L = _lib.load()
    i16 = x.dtype == torch.int16
    _need_cuda(x, "conv input", torch.int16 if i16 else torch.float32)
    _need_cuda(gz, "gz", torch.float32)
    B, Ci, Ti, Hi, Wi = x.shape
    Co = gz.shape[1]
    if tuple(gz.shape) != (B, Co, Ti + 2 * pad_t - 2, Hi + 2 * pad_hw - 2, Wi + 2 * pad_hw - 2):
        raise RuntimeError(f"conv3d_wgrad: gz shape {tuple(gz.shape)} inconsistent with input {tuple(x.shape)}")
    dw = torch.empty((Co, Ci, 3, 3, 3), dtype=torch.float32, device=x.device)
    db = torch.empty((Co,), dtype=torch.float32, device=x.device)
    nb = L.pvb200_conv3d_wgrad_workspace_bytes(Ci, Co)
    ws = _workspace("wgrad", nb, x.device)
    with _timed(f"conv3d_wgrad_f32[Ci={Ci}]", 2.0 * 27 * Ci * gz.numel(), x.numel() * x.element_size() + 4.0 * gz.numel()):
        rc = L.pvb200_conv3d_wgrad_f32_pad(_p(x), int(i16), _p(mean) if i16 else None, _p(std) if i16 else None, _p(gz), _p(dw),
                                           _p(db), _p(ws), ws.numel(), B, Ci, Ti, Hi, Wi, Co, pad_t, pad_hw, _stream())
    _lib.check(rc, "conv3d_wgrad")
    return dw, db


# ---- bf16 tensor-core path: blocked [B][Cg][T][H][W][8] activations -------------------------------------------
def blocked_groups(C: int) -> int:
    return int(_lib.load().pvb200_blocked_channel_groups(C))


def to_blocked_bf16(x: torch.Tensor, pad: int = 0) -> torch.Tensor:
    """[B,C,T,H,W] fp32 -> blocked bf16 [B,Cg,T+2p,H+2p,W+2p,8] (zero border, zero pad channels)."""
    L = _lib.load()
    _need_cuda(x, "x", torch.float32)
    B, Cc, T, H, W = x.shape
    Cg = blocked_groups(Cc)
    alloc = torch.zeros if pad > 0 else torch.empty
    y = alloc((B, Cg, T + 2 * pad, H + 2 * pad, W + 2 * pad, 8), dtype=torch.bfloat16, device=x.device)
    with _timed("nc_to_blocked_bf16", 0.0, 4.0 * x.numel() + 2.0 * y.numel()):
        rc = L.pvb200_nc_to_blocked_bf16(_p(x), _p(y), B, Cc, T, H, W, pad, _stream())
    _lib.check(rc, "nc_to_blocked_bf16")
    return y


def from_blocked_bf16(xb: torch.Tensor, C: int) -> torch.Tensor:
    """blocked bf16 [B,Cg,T,H,W,8] -> [B,C,T,H,W] fp32."""
    L = _lib.load()
    _need_cuda(xb, "xb", torch.bfloat16)
    B, Cg, T, H, W, e = xb.shape
    if e != 8 or Cg != blocked_groups(C):
        raise RuntimeError("from_blocked_bf16: shape does not match the channel count")
    y = torch.empty((B, C, T, H, W), dtype=torch.float32, device=xb.device)
    with _timed("blocked_to_nc_f32", 0.0, 2.0 * xb.numel() + 4.0 * y.numel()):
        rc = L.pvb200_blocked_to_nc_f32(_p(xb), _p(y), B, C, T, H, W, _stream())
    _lib.check(rc, "blocked_to_nc_f32")
    return y


def conv3d_fwd_bf16(xb: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], relu: bool = True,
                    out_pad: int = 0, pad_t: int = 0) -> torch.Tensor:
    """Blocked bf16 conv3d 3x3x3 (+bias, +ReLU) on the tensor cores; fp32 master weights [Co,Ci,3,3,3]."""
    L = _lib.load()
    _need_cuda(xb, "xb", torch.bfloat16)
    _need_cuda(w, "conv weight", torch.float32)
    B, Cg, Ti, Hi, Wi, e = xb.shape
    Co, Ci = w.shape[0], w.shape[1]
    if e != 8 or Cg != blocked_groups(Ci):
        raise RuntimeError(f"conv3d_fwd_bf16: input has {Cg} channel groups, weight expects Cin={Ci}")
    alloc = torch.zeros if out_pad > 0 else torch.empty
    yb = alloc((B, blocked_groups(Co), Ti + 2 * pad_t - 2 + 2 * out_pad, Hi - 2 + 2 * out_pad, Wi - 2 + 2 * out_pad, 8),
               dtype=torch.bfloat16, device=xb.device)
    ws = _workspace("conv_bf16", L.pvb200_conv3d_bf16_workspace_bytes(Ci, Co), xb.device)
    npos = B * (Ti + 2 * pad_t - 2) * (Hi - 2) * (Wi - 2)
    with _timed(f"conv3d_fwd_bf16[Ci={Ci}]", 2.0 * 27 * Ci * Co * npos, 2.0 * (xb.numel() + 8 * blocked_groups(Co) * npos)):
        rc = L.pvb200_conv3d_fwd_bf16_tpad(_p(xb), _p(w), _p(b), _p(yb), _p(ws), ws.numel(), B, Ci, Ti, Hi, Wi, Co, int(relu),
                                           out_pad, pad_t, _stream())
    _lib.check(rc, "conv3d_fwd_bf16")
    return yb


def conv3d_dgrad_bf16(gz_padded: torch.Tensor, w: torch.Tensor, mask_src: Optional[torch.Tensor], out_pad: int = 0,
                      also_gzw: bool = False, persistent: bool = False, pad_t: int = 0):
    """gx (blocked bf16, optionally written into a padded tensor) from gz zero-padded by 2 on T,H,W.
    ``also_gzw``: additionally return gx in the weight-gradient operand layout of the layer below."""
    L = _lib.load()
    _need_cuda(gz_padded, "gz_padded", torch.bfloat16)
    _need_cuda(w, "conv weight", torch.float32)
    B, Cgo, Tp, Hp, Wp, e = gz_padded.shape
    Co, Ci = w.shape[0], w.shape[1]
    Ti, Hi, Wi = Tp - 2 - 2 * pad_t, Hp - 2, Wp - 2
    if e != 8 or Cgo != blocked_groups(Co):
        raise RuntimeError("conv3d_dgrad_bf16: gz channel groups do not match the weight")
    Cgi = blocked_groups(Ci)
    if mask_src is not None:
        _need_cuda(mask_src, "mask_src", torch.bfloat16)
        if tuple(mask_src.shape) != (B, Cgi, Ti, Hi, Wi, 8):
            raise RuntimeError("conv3d_dgrad_bf16: mask_src shape mismatch")
    # zero-bordered outputs live in persistent buffers (one per shape): the kernel rewrites every valid position, the
    # border / wrap columns stay zero from the first allocation, so there is no per-step memset (6 x ~100 MB per step)
    gx_shape = (B, Cgi, Ti + 2 * out_pad, Hi + 2 * out_pad, Wi + 2 * out_pad, 8)
    if out_pad > 0 and persistent:
        gx = _zero_bordered("gx_pad", gx_shape, gz_padded.device)
    else:
        gx = (torch.zeros if out_pad > 0 else torch.empty)(gx_shape, dtype=torch.bfloat16, device=gz_padded.device)
    gzw = None
    if also_gzw:
        QP = int(L.pvb200_conv3d_wgrad_bf16_gz_plane(Hi + 2, Wi + 2))
        if persistent:
            # keyed by the plane GEOMETRY too: two layers whose planes round up to the same QP must not share a buffer
            # (the positions that stay zero differ with the pitch)
            gzw = _zero_bordered(f"gzw_{Hi}x{Wi}", (B, Cgi, Ti, QP, 8), gz_padded.device)
        else:
            gzw = torch.zeros((B, Cgi, Ti, QP, 8), dtype=torch.bfloat16, device=gz_padded.device)
    ws = _workspace("conv_bf16", L.pvb200_conv3d_bf16_workspace_bytes(Ci, Co), gz_padded.device)
    npos = B * Ti * Hi * Wi
    with _timed(f"conv3d_dgrad_bf16[Ci={Ci}]", 2.0 * 27 * Ci * Co * npos,
                2.0 * (gz_padded.numel() + 8 * Cgi * npos * ((2 if mask_src is not None else 1) + (1 if also_gzw else 0)))):
        rc = L.pvb200_conv3d_dgrad_bf16_tpad(_p(gz_padded), _p(w), _p(mask_src), _p(gx), _p(gzw), _p(ws), ws.numel(), B, Ci, Ti, Hi,
                                             Wi, Co, out_pad, pad_t, _stream())
    _lib.check(rc, "conv3d_dgrad_bf16")
    return (gx, gzw) if also_gzw else gx


def sat_normalise_blocked_bf16(x: torch.Tensor, mean: torch.Tensor, std: torch.Tensor) -> torch.Tensor:
    """int16 [B,C,T,H,W] -> normalised blocked bf16 [B,Cg,T,H,W,8] (a1 fused with the layout change)."""
    L = _lib.load()
    _need_cuda(x, "satellite.data", torch.int16)
    _need_cuda(mean, "sat_mean", torch.float32)
    _need_cuda(std, "sat_std", torch.float32)
    B, Cc, T, H, W = x.shape
    y = torch.empty((B, blocked_groups(Cc), T, H, W, 8), dtype=torch.bfloat16, device=x.device)
    with _timed("sat_normalise_blocked_bf16", 0.0, 2.0 * x.numel() + 2.0 * y.numel()):
        rc = L.pvb200_sat_normalise_blocked_bf16(_p(x), _p(y), _p(mean), _p(std), B, Cc, T, H, W, _stream())
    _lib.check(rc, "sat_normalise_blocked_bf16")
    return y


def to_gzw_bf16(gz: torch.Tensor) -> torch.Tensor:
    """[B,Co,To,Ho,Wo] fp32 -> the wgrad operand layout: blocked bf16 with the input pitch, [B,Cg,To,QP,8]."""
    L = _lib.load()
    _need_cuda(gz, "gz", torch.float32)
    B, Co, To, Ho, Wo = gz.shape
    QP = int(L.pvb200_conv3d_wgrad_bf16_gz_plane(Ho + 2, Wo + 2))
    out = torch.empty((B, blocked_groups(Co), To, QP, 8), dtype=torch.bfloat16, device=gz.device)
    with _timed("nc_to_gzw_bf16", 0.0, 4.0 * gz.numel() + 2.0 * out.numel()):
        rc = L.pvb200_nc_to_gzw_bf16(_p(gz), _p(out), B, Co, To, Ho, Wo, _stream())
    _lib.check(rc, "nc_to_gzw_bf16")
    return out


def conv3d_wgrad_bf16(xb: torch.Tensor, gzw: torch.Tensor, Ci: int, Co: int, pad_t: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """(dw [Co,Ci,3,3,3], db [Co]) fp32 from blocked bf16 x [B,Cg,Ti,Hi,Wi,8] and gzw [B,Cg,To,QP,8]."""
    L = _lib.load()
    _need_cuda(xb, "xb", torch.bfloat16)
    _need_cuda(gzw, "gzw", torch.bfloat16)
    B, Cgx, Ti, Hi, Wi, e = xb.shape
    QP = int(L.pvb200_conv3d_wgrad_bf16_gz_plane(Hi, Wi))
    if e != 8 or Cgx != blocked_groups(Ci) or tuple(gzw.shape) != (B, blocked_groups(Co), Ti + 2 * pad_t - 2, QP, 8):
        raise RuntimeError(f"conv3d_wgrad_bf16: shapes {tuple(xb.shape)} / {tuple(gzw.shape)} inconsistent")
    dw = torch.empty((Co, Ci, 3, 3, 3), dtype=torch.float32, device=xb.device)
    db = torch.empty((Co,), dtype=torch.float32, device=xb.device)
    ws = _workspace("wgrad_bf16", L.pvb200_conv3d_wgrad_bf16_workspace_bytes(Ci, Co), xb.device)
    npos = B * (Ti + 2 * pad_t - 2) * (Hi - 2) * (Wi - 2)
    with _timed(f"conv3d_wgrad_bf16[Ci={Ci}]", 2.0 * 27 * Ci * Co * npos, 2.0 * (xb.numel() + gzw.numel())):
        rc = L.pvb200_conv3d_wgrad_bf16_tpad(_p(xb), _p(gzw), _p(dw), _p(db), _p(ws), ws.numel(), B, Ci, Ti, Hi, Wi, Co, pad_t,
                                             _stream())
    _lib.check(rc, "conv3d_wgrad_bf16")
    return dw, db


# ---- fp32 mode on the tensor cores (3xTF32): blocked fp32 [B][G][T][H][W][4] activations ----------------------------
def blocked4_groups(C: int) -> int:
    return int(_lib.load().pvb200_blocked4_channel_groups(C))


def tf32x3_supported(Ci: int, Co: int) -> bool:
    """Shapes the 3xTF32 tensor-core convolutions take (the fp32 direct kernels cover everything else)."""
    return Ci <= 32 and Co <= 32


def conv_f16x2_applies(Ci: int, Co: int) -> bool:
    """Shapes for which the implicit GEMM takes the two-way fp16 split when it is given the input's largest magnitude
    (CTA-pair kernel, 16-channel steps); other shapes run 3xTF32.  ``Ci`` / ``Co``: the kernel's input / output role."""
    return Co > 16 and blocked4_groups(Ci) % 4 == 0


def to_blocked_f32(x: torch.Tensor, pad: int = 0, persistent: bool = False, amax: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[B,C,T,H,W] fp32 -> blocked fp32 [B,G,T+2p,H+2p,W+2p,4] (zero border, zero pad channels).  ``amax`` (here and in the
    operators below): one zeroed fp32 device element that receives the largest magnitude written."""
    L = _lib.load()
    _need_cuda(x, "x", torch.float32)
    B, Cc, T, H, W = x.shape
    G = blocked4_groups(Cc)
    shape = (B, G, T + 2 * pad, H + 2 * pad, W + 2 * pad, 4)
    if pad > 0 and persistent:
        y = _zero_bordered("blk4_pad", shape, x.device, torch.float32)
    else:
        y = (torch.zeros if pad > 0 else torch.empty)(shape, dtype=torch.float32, device=x.device)
    with _timed("nc_to_blocked_f32", 0.0, 4.0 * x.numel() + 16.0 * G * B * T * H * W):
        rc = L.pvb200_nc_to_blocked_f32(_p(x), _p(y), B, Cc, T, H, W, pad, _p(amax), _stream())
    _lib.check(rc, "nc_to_blocked_f32")
    return y


def from_blocked_f32(xb: torch.Tensor, C: int) -> torch.Tensor:
    """blocked fp32 [B,G,T,H,W,4] -> [B,C,T,H,W] fp32."""
    L = _lib.load()
    _need_cuda(xb, "xb", torch.float32)
    B, G, T, H, W, e = xb.shape
    if e != 4 or G != blocked4_groups(C):
        raise RuntimeError("from_blocked_f32: shape does not match the channel count")
    y = torch.empty((B, C, T, H, W), dtype=torch.float32, device=xb.device)
    with _timed("blocked_f32_to_nc", 0.0, 4.0 * xb.numel() + 4.0 * y.numel()):
        rc = L.pvb200_blocked_f32_to_nc(_p(xb), _p(y), B, C, T, H, W, _stream())
    _lib.check(rc, "blocked_f32_to_nc")
    return y


def sat_normalise_blocked_f32(x: torch.Tensor, mean: torch.Tensor, std: torch.Tensor, amax: Optional[torch.Tensor] = None) -> torch.Tensor:
    """int16 [B,C,T,H,W] -> normalised blocked fp32 [B,G,T,H,W,4] (a1 fused with the layout change, bit-identical)."""
    L = _lib.load()
    _need_cuda(x, "satellite.data", torch.int16)
    _need_cuda(mean, "sat_mean", torch.float32)
    _need_cuda(std, "sat_std", torch.float32)
    B, Cc, T, H, W = x.shape
    if mean.numel() != Cc or std.numel() != Cc:
        raise RuntimeError(f"sat_normalise: {Cc} channels but {mean.numel()} means / {std.numel()} stds")
    y = torch.empty((B, blocked4_groups(Cc), T, H, W, 4), dtype=torch.float32, device=x.device)
    with _timed("sat_normalise_blocked_f32", 0.0, 2.0 * x.numel() + 4.0 * y.numel()):
        rc = L.pvb200_sat_normalise_blocked_f32(_p(x), _p(y), _p(mean), _p(std), B, Cc, T, H, W, _p(amax), _stream())
    _lib.check(rc, "sat_normalise_blocked_f32")
    return y


def conv3d_fwd_tf32x3(xb: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], relu: bool = True, out_pad: int = 0,
                      pad_t: int = 0, want_blk: bool = True, want_nc: bool = False, amax: Optional[torch.Tensor] = None,
                      amax_in: Optional[torch.Tensor] = None):
    """relu(conv3d(x, w, b)) on the tensor cores at fp32 accuracy (3xTF32).  xb blocked fp32 [B,G,Ti,Hi,Wi,4]; returns
    (blocked copy | None, NCDHW copy | None)."""
    L = _lib.load()
    _need_cuda(xb, "xb", torch.float32)
    _need_cuda(w, "conv weight", torch.float32)
    B, G, Ti, Hi, Wi, e = xb.shape
    Co, Ci = w.shape[0], w.shape[1]
    if e != 4 or G != blocked4_groups(Ci):
        raise RuntimeError(f"conv3d_fwd_tf32x3: input has {G} channel groups, weight expects Cin={Ci}")
    To, Ho, Wo = Ti + 2 * pad_t - 2, Hi - 2, Wi - 2
    GO = blocked4_groups(Co)
    y_blk = y_nc = None
    if want_blk:
        y_blk = (torch.zeros if out_pad > 0 else torch.empty)((B, GO, To + 2 * out_pad, Ho + 2 * out_pad, Wo + 2 * out_pad, 4),
                                                              dtype=torch.float32, device=xb.device)
    if want_nc:
        y_nc = torch.empty((B, Co, To, Ho, Wo), dtype=torch.float32, device=xb.device)
    ws = _workspace("conv_tf32x3", L.pvb200_conv3d_tf32x3_workspace_bytes(Ci, Co), xb.device)
    npos = B * To * Ho * Wo
    kind = "f16x2" if (amax_in is not None and conv_f16x2_applies(Ci, Co)) else "tf32x3"
    with _timed(f"conv3d_fwd_{kind}[Ci={Ci}]", 2.0 * 27 * Ci * Co * npos,
                4.0 * xb.numel() + 4.0 * npos * (4 * GO * int(want_blk) + Co * int(want_nc))):
        rc = L.pvb200_conv3d_fwd_tf32x3(_p(xb), _p(w), _p(b), _p(y_blk), _p(y_nc), _p(ws), ws.numel(), B, Ci, Ti, Hi, Wi, Co,
                                        int(relu), out_pad, pad_t, _p(amax_in), _p(amax), _stream())
    _lib.check(rc, "conv3d_fwd_tf32x3")
    return y_blk, y_nc


def conv3d_dgrad_tf32x3(gz_padded: torch.Tensor, w: torch.Tensor, mask_blk: Optional[torch.Tensor], out_pad: int = 0,
                        want_blk: bool = True, want_nc: bool = False, persistent: bool = False, pad_t: int = 0,
                        amax: Optional[torch.Tensor] = None, amax_in: Optional[torch.Tensor] = None):
    """gx = conv_transpose3d(gz, w) * (mask > 0) on the tensor cores (3xTF32) from gz blocked fp32 zero-padded by 2 on
    T, H, W.  Returns (blocked copy | None, NCDHW copy | None)."""
    L = _lib.load()
    _need_cuda(gz_padded, "gz_padded", torch.float32)
    _need_cuda(w, "conv weight", torch.float32)
    B, GOz, Tp, Hp, Wp, e = gz_padded.shape
    Co, Ci = w.shape[0], w.shape[1]
    Ti, Hi, Wi = Tp - 2 - 2 * pad_t, Hp - 2, Wp - 2
    if e != 4 or GOz != blocked4_groups(Co):
        raise RuntimeError("conv3d_dgrad_tf32x3: gz channel groups do not match the weight")
    GI = blocked4_groups(Ci)
    if mask_blk is not None:
        _need_cuda(mask_blk, "mask_blk", torch.float32)
        if tuple(mask_blk.shape) != (B, GI, Ti, Hi, Wi, 4):
            raise RuntimeError("conv3d_dgrad_tf32x3: mask shape mismatch")
    gx_blk = gx_nc = None
    if want_blk:
        shape = (B, GI, Ti + 2 * out_pad, Hi + 2 * out_pad, Wi + 2 * out_pad, 4)
        if out_pad > 0 and persistent:
            gx_blk = _zero_bordered("gx4_pad", shape, gz_padded.device, torch.float32)
        else:
            gx_blk = (torch.zeros if out_pad > 0 else torch.empty)(shape, dtype=torch.float32, device=gz_padded.device)
    if want_nc:
        gx_nc = torch.empty((B, Ci, Ti, Hi, Wi), dtype=torch.float32, device=gz_padded.device)
    ws = _workspace("conv_tf32x3", L.pvb200_conv3d_tf32x3_workspace_bytes(Ci, Co), gz_padded.device)
    npos = B * Ti * Hi * Wi
    kind = "f16x2" if (amax_in is not None and conv_f16x2_applies(Co, Ci)) else "tf32x3"
    with _timed(f"conv3d_dgrad_{kind}[Ci={Ci}]", 2.0 * 27 * Ci * Co * npos,
                4.0 * gz_padded.numel() + 4.0 * npos * (4 * GI * (int(want_blk) + int(mask_blk is not None)) + Ci * int(want_nc))):
        rc = L.pvb200_conv3d_dgrad_tf32x3(_p(gz_padded), _p(w), _p(mask_blk), _p(gx_blk), _p(gx_nc), _p(ws), ws.numel(), B, Ci, Ti,
                                          Hi, Wi, Co, out_pad, pad_t, _p(amax_in), _p(amax), _stream())
    _lib.check(rc, "conv3d_dgrad_tf32x3")
    return gx_blk, gx_nc


def wgrad_bf16x3_supported(Ci: int, Co: int, Hi: int, Wi: int) -> bool:
    return bool(_lib.load().pvb200_conv3d_wgrad_bf16x3_supported(Ci, Co, Hi, Wi))


def absmax_f32(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out[0] = max(out[0], max |x|) on the device (``out``: one fp32 element the caller has zeroed)."""
    L = _lib.load()
    _need_cuda(x, "x", torch.float32)
    _need_cuda(out, "out", torch.float32)
    if not x.is_contiguous():
        raise RuntimeError("absmax_f32: x must be contiguous")
    with _timed("absmax_f32", 0.0, 4.0 * x.numel()):
        rc = L.pvb200_absmax_f32(_p(x), x.numel(), _p(out), _stream())
    _lib.check(rc, "absmax_f32")
    return out


def conv3d_wgrad_bf16x3(xb: torch.Tensor, gzb: torch.Tensor, Ci: int, Co: int, gz_pad: int = 0,
                        pad_t: int = 0, amax: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(dw [Co,Ci,3,3,3], db [Co]) on the tensor cores from x blocked fp32 [B,G,Ti,Hi,Wi,4] and the pre-activation
    gradient blocked fp32, zero-padded by ``gz_pad`` on T, H, W (2 = the tensor the data gradient reads).  Three-way bf16
    split, or -- when ``amax`` = (max |x|, max |gz|), two one-element device tensors, is given -- the two-way fp16 split of
    the scaled operands."""
    L = _lib.load()
    _need_cuda(xb, "xb", torch.float32)
    _need_cuda(gzb, "gzb", torch.float32)
    B, G, Ti, Hi, Wi, e = xb.shape
    To, Ho, Wo = Ti + 2 * pad_t - 2, Hi - 2, Wi - 2
    want = (B, blocked4_groups(Co), To + 2 * gz_pad, Ho + 2 * gz_pad, Wo + 2 * gz_pad, 4)
    if e != 4 or G != blocked4_groups(Ci) or tuple(gzb.shape) != want:
        raise RuntimeError(f"conv3d_wgrad_bf16x3: shapes {tuple(xb.shape)} / {tuple(gzb.shape)} inconsistent (expected gz {want})")
    dw = torch.empty((Co, Ci, 3, 3, 3), dtype=torch.float32, device=xb.device)
    db = torch.empty((Co,), dtype=torch.float32, device=xb.device)
    ws = _workspace("wgrad_bf16x3", L.pvb200_conv3d_wgrad_bf16x3_workspace_bytes(), xb.device)
    npos = B * To * Ho * Wo
    with _timed(f"conv3d_wgrad_{'bf16x3' if amax is None else 'f16x2'}[Ci={Ci}]", 2.0 * 27 * Ci * Co * npos, 4.0 * xb.numel() + 16.0 * blocked4_groups(Co) * npos):
        if amax is None:
            rc = L.pvb200_conv3d_wgrad_bf16x3(_p(xb), _p(gzb), gz_pad, _p(dw), _p(db), _p(ws), ws.numel(), B, Ci, Ti, Hi, Wi, Co, pad_t,
                                              _stream())
        else:
            _need_cuda(amax[0], "amax_x", torch.float32)
            _need_cuda(amax[1], "amax_gz", torch.float32)
            rc = L.pvb200_conv3d_wgrad_f16x2(_p(xb), _p(gzb), gz_pad, _p(amax[0]), _p(amax[1]), _p(dw), _p(db), _p(ws), ws.numel(), B, Ci,
                                             Ti, Hi, Wi, Co, pad_t, _stream())
    _lib.check(rc, "conv3d_wgrad_bf16x3")
    return dw, db


def wgrad_bf16_rows_supported(Ci: int, Co: int, Hi: int, Wi: int) -> bool:
    return bool(_lib.load().pvb200_conv3d_wgrad_bf16_rows_supported(Ci, Co, Hi, Wi))


def conv3d_wgrad_bf16_rows(xb: torch.Tensor, gzb: torch.Tensor, Ci: int, Co: int, gz_pad: int = 0,
                           pad_t: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """(dw [Co,Ci,3,3,3], db [Co]) fp32 from blocked bf16 x [B,Cg,Ti,Hi,Wi,8] and the pre-activation gradient blocked bf16
    zero-padded by ``gz_pad`` on T, H, W (row-step kernel: TMA tensor maps straight into the operand layout)."""
    L = _lib.load()
    _need_cuda(xb, "xb", torch.bfloat16)
    _need_cuda(gzb, "gzb", torch.bfloat16)
    B, Cgx, Ti, Hi, Wi, e = xb.shape
    To, Ho, Wo = Ti + 2 * pad_t - 2, Hi - 2, Wi - 2
    want = (B, blocked_groups(Co), To + 2 * gz_pad, Ho + 2 * gz_pad, Wo + 2 * gz_pad, 8)
    if e != 8 or Cgx != blocked_groups(Ci) or tuple(gzb.shape) != want:
        raise RuntimeError(f"conv3d_wgrad_bf16_rows: shapes {tuple(xb.shape)} / {tuple(gzb.shape)} inconsistent (expected gz {want})")
    dw = torch.empty((Co, Ci, 3, 3, 3), dtype=torch.float32, device=xb.device)
    db = torch.empty((Co,), dtype=torch.float32, device=xb.device)
    ws = _workspace("wgrad_bf16_rows", L.pvb200_conv3d_wgrad_bf16_rows_workspace_bytes(), xb.device)
    npos = B * To * Ho * Wo
    with _timed(f"conv3d_wgrad_bf16[Ci={Ci}]", 2.0 * 27 * Ci * Co * npos, 2.0 * (xb.numel() + 8 * blocked_groups(Co) * npos)):
        rc = L.pvb200_conv3d_wgrad_bf16_rows(_p(xb), _p(gzb), gz_pad, _p(dw), _p(db), _p(ws), ws.numel(), B, Ci, Ti, Hi, Wi, Co, pad_t,
                                             _stream())
    _lib.check(rc, "conv3d_wgrad_bf16_rows")
    return dw, db


def adam_step(params: List[torch.Tensor], grads: List[torch.Tensor], exp_avg: List[torch.Tensor],
              exp_avg_sq: List[torch.Tensor], lr: float, beta1: float, beta2: float, eps: float, step: int,
              grad_scale: float = 1.0) -> None:
    """One multi-tensor Adam update in place (torch.optim.Adam arithmetic, base_model.py:255-257)."""
    L = _lib.load()
    n = len(params)
    for i in range(n):
        for t, nm in ((params[i], "param"), (grads[i], "grad"), (exp_avg[i], "exp_avg"), (exp_avg_sq[i], "exp_avg_sq")):
            _need_cuda(t, nm, torch.float32)
        if not (params[i].numel() == grads[i].numel() == exp_avg[i].numel() == exp_avg_sq[i].numel()):
            raise RuntimeError("adam_step: size mismatch")
    arr = lambda ts: (C.c_void_p * n)(*[t.data_ptr() for t in ts])  # noqa: E731
    numel = (C.c_longlong * n)(*[p.numel() for p in params])
    with _timed("adam_step_f32", 0.0, 28.0 * sum(p.numel() for p in params)):
        rc = L.pvb200_adam_step_f32(n, arr(params), arr(grads), arr(exp_avg), arr(exp_avg_sq), numel, lr, beta1, beta2,
                                    eps, step, grad_scale, _stream())
    _lib.check(rc, "adam_step")


# ---------------------------------------------------------------------------------------------
# autograd nodes
# ---------------------------------------------------------------------------------------------
class EncoderFn(torch.autograd.Function):
    """Conv3d stack.  forward(sat, mean, std, w0, b0, w1, b1, ...) -> features [B, cnn_output_size].

    ``sat`` is int16 (normalisation fused into layer 0's loads) or already-normalised fp32.
    PRIVATE PROTOCOL: the incoming gradient must already carry the ReLU mask of the last layer
    (``HeadFn`` applies it inside its fc1 data-gradient kernel), i.e. it is the gradient w.r.t. the last
    layer's pre-activation.  Layer l's data-gradient kernel fuses the mask of layer l-1.
    """

    @staticmethod
    def forward(ctx, sat, mean, std, *wb):
        n_layers = len(wb) // 2
        acts = []
        # int16 cube: one streaming pass of the normalise kernel (128-bit loads, 6 B/element), so that layer 0's
        # forward and weight gradient use the cp.async-pipelined fp32 loaders (measured 1.5x faster than converting
        # inside their staging loops); the conv kernels keep their fused-int16 loaders for callers that want them.
        x = sat_normalise(sat, mean, std) if sat.dtype == torch.int16 else sat
        x0 = x
        for l in range(n_layers):
            x = conv3d_fwd(x, wb[2 * l], wb[2 * l + 1], relu=True)
            acts.append(x)
        ctx.save_for_backward(x0, mean, std, *wb, *acts)
        ctx.n_layers = n_layers
        return acts[-1].view(sat.shape[0], -1)

    @staticmethod
    def backward(ctx, g):
        n = ctx.n_layers
        saved = ctx.saved_tensors
        sat, mean, std = saved[0], saved[1], saved[2]
        wb = saved[3: 3 + 2 * n]
        acts = saved[3 + 2 * n:]
        gz = g.contiguous().view(acts[-1].shape)
        grads: List[Optional[torch.Tensor]] = [None] * (2 * n)
        for l in range(n - 1, -1, -1):
            x = sat if l == 0 else acts[l - 1]
            dw, db = conv3d_wgrad(x, gz)
            grads[2 * l], grads[2 * l + 1] = dw, db
            if l > 0:
                gz = conv3d_dgrad(gz, wb[2 * l], acts[l - 1], acts[l - 1].shape)
        return (None, None, None, *grads)


# fp32 mode, tensor-core weight gradient: two-way fp16 split of the scaled operands (three products) instead of the three-way
# bf16 split (six); the largest magnitudes it scales by come out of the kernels that write the tensors
FP32_WGRAD_F16X2 = True
# the same split in the forward / data-gradient implicit GEMM (instead of 3xTF32): K = 16 channels per MMA, half the MMAs
FP32_CONV_F16X2 = True


class EncoderTf32Fn(torch.autograd.Function):
    """Conv3d stack of the fp32 mode on the tensor cores (3xTF32: forward, data gradient AND weight gradient at fp32-class
    accuracy, csrc/conv3d_igemm_tf32x3.cu / conv3d_wgrad_bf16x3.cu).

    forward(sat, mean, std, w0, b0, ...) -> features fp32 [B, cnn_output_size] (NCDHW flatten order, model.py:122).
    Activations and gradients live blocked ([B,G,T,H,W,4], what the tensor-core kernels read); the last activation is
    also written in NCDHW for the fp32 head.  A layer whose rows do not fit the tensor-core weight gradient's shared
    memory (128-wide planes) falls back to the fp32 FMA weight gradient, for which the producing kernels' epilogues write
    an extra NCDHW copy (no conversion pass).  Same private protocol as ``EncoderFn``: the incoming gradient already
    carries the ReLU mask of the last layer."""

    @staticmethod
    def forward(ctx, sat, mean, std, *wb):
        n = len(wb) // 2
        B, _, T, H, W = sat.shape
        # per layer: does the tensor-core weight gradient take it?  (input plane of layer l: H - 2l)
        tc_w = [wgrad_bf16x3_supported(wb[2 * l].shape[1], wb[2 * l].shape[0], H - 2 * l, W - 2 * l) for l in range(n)]
        # amax[l] = max |input of layer l|, amax[n + l] = max |gradient w.r.t. layer l's pre-activation| (filled in backward)
        amax = torch.zeros((2 * n,), dtype=torch.float32, device=sat.device) if (FP32_WGRAD_F16X2 or FP32_CONV_F16X2) else None
        am = (lambda i: amax[i:i + 1]) if amax is not None else (lambda i: None)
        am_in = am if FP32_CONV_F16X2 else (lambda i: None)
        if sat.dtype == torch.int16:
            x_blk = sat_normalise_blocked_f32(sat, mean, std, amax=am(0))
            x_nc = None if tc_w[0] else sat_normalise(sat, mean, std)
        else:
            x_blk = to_blocked_f32(sat, amax=am(0))
            x_nc = None if tc_w[0] else sat
        ins_blk, ins_nc = [x_blk], [x_nc]  # input of layer l, blocked / NCDHW (None when not needed)
        y_nc = None
        for l in range(n):
            last = l == n - 1
            want_nc = last or not tc_w[l + 1]
            y_blk, y_nc = conv3d_fwd_tf32x3(ins_blk[l], wb[2 * l], wb[2 * l + 1], relu=True, want_blk=not last, want_nc=want_nc,
                                            amax=None if last else am(l + 1), amax_in=am_in(l))
            if not last:
                ins_blk.append(y_blk)
                ins_nc.append(y_nc if not tc_w[l + 1] else None)
        keep = [t for t in ins_nc if t is not None]
        ctx.save_for_backward(*wb, *ins_blk, *keep)
        ctx.n_layers, ctx.tc_w, ctx.amax, ctx.f16 = n, tc_w, amax, (FP32_WGRAD_F16X2, FP32_CONV_F16X2)
        ctx.nc_index = [i for i, t in enumerate(ins_nc) if t is not None]
        ctx.out_shape = tuple(y_nc.shape)
        return y_nc.view(B, -1)

    @staticmethod
    def backward(ctx, g):
        n, tc_w = ctx.n_layers, ctx.tc_w
        saved = ctx.saved_tensors
        wb = saved[: 2 * n]
        ins_blk = saved[2 * n: 3 * n]
        ins_nc: List[Optional[torch.Tensor]] = [None] * n
        for i, t in zip(ctx.nc_index, saved[3 * n:]):
            ins_nc[i] = t
        gz_nc = g.contiguous().view(ctx.out_shape)
        gz_blk, gz_pad = None, 0
        amax = ctx.amax
        am = (lambda i: amax[i:i + 1]) if amax is not None else (lambda i: None)
        if n > 1 or tc_w[n - 1]:
            gz_blk, gz_pad = to_blocked_f32(gz_nc, pad=2, persistent=True, amax=am(2 * n - 1)), 2
        grads: List[Optional[torch.Tensor]] = [None] * (2 * n)
        for l in range(n - 1, -1, -1):
            Co, Ci = wb[2 * l].shape[0], wb[2 * l].shape[1]
            if tc_w[l]:
                dw, db = conv3d_wgrad_bf16x3(ins_blk[l], gz_blk, Ci, Co, gz_pad=gz_pad,
                                             amax=(am(l), am(n + l)) if ctx.f16[0] else None)
            else:
                dw, db = conv3d_wgrad(ins_nc[l], gz_nc)
            grads[2 * l], grads[2 * l + 1] = dw, db
            if l > 0:
                # gradient w.r.t. layer l-1's pre-activation: blocked + padded by 2 when a data gradient still follows
                # (l-1 >= 1), blocked when layer l-1's weight gradient runs on the tensor cores, NCDHW when it does not
                nxt_pad = 2 if l - 1 >= 1 else 0
                want_blk = (l - 1 >= 1) or tc_w[l - 1]
                gz_blk, gz_nc = conv3d_dgrad_tf32x3(gz_blk, wb[2 * l], ins_blk[l], out_pad=nxt_pad if want_blk else 0,
                                                    want_blk=want_blk, want_nc=not tc_w[l - 1], persistent=True, amax=am(n + l - 1),
                                                    amax_in=am(n + l) if ctx.f16[1] else None)
                gz_pad = nxt_pad
        return (None, None, None, *grads)


class EncoderBf16Fn(torch.autograd.Function):
    """Conv3d stack in bf16 on the tensor cores (tcgen05 implicit GEMM), fp32 master weights.

    forward(sat, mean, std, w0, b0, ...) -> features fp32 [B, cnn_output_size] in the reference's NCDHW flatten order.
    Activations are kept blocked bf16 ([B,Cg,T,H,W,8]).  Same private protocol as ``EncoderFn``: the incoming gradient
    already carries the ReLU mask of the last layer.  Backward, per layer: weight gradient (tensor cores, MN-major
    operands) from the blocked input and the gradient in "gzw" layout; data gradient (same implicit-GEMM kernel on the
    zero-padded gradient) writes the next gradient in BOTH layouts it is needed in, with the ReLU mask fused."""

    @staticmethod
    def forward(ctx, link, sat, mean, std, *wb):
        """``link``: None -> return fp32 NCDHW features (consumed by the fp32 ``HeadFn``); a dict -> return the last
        activation itself in blocked bf16 (consumed by ``HeadBf16Fn``, which hands the gradient back through
        ``link["gz"]`` already in the two layouts the conv backward needs)."""
        n_layers = len(wb) // 2
        pad_t = int(link.get("pad_t", 0)) if link is not None else 0  # time padding of every layer (conv3d_sat_nwp towers: 1)
        if sat.dtype == torch.int16:
            x = sat_normalise_blocked_bf16(sat, mean, std)
        else:
            x = to_blocked_bf16(sat)
        acts = [x]
        for l in range(n_layers):
            x = conv3d_fwd_bf16(x, wb[2 * l], wb[2 * l + 1], relu=True, pad_t=pad_t)
            acts.append(x)
        ctx.save_for_backward(*wb, *acts)
        ctx.n_layers = n_layers
        ctx.pad_t = pad_t
        ctx.channels = [wb[2 * l].shape[1] for l in range(n_layers)] + [wb[-2].shape[0]]
        ctx.link = link
        if link is not None:
            # does the last layer's weight gradient need the gradient in the round-1 kernel's "gzw" layout as well?
            link["need_gzw"] = not wgrad_bf16_rows_supported(ctx.channels[-2], ctx.channels[-1], acts[-2].shape[3], acts[-2].shape[4])
            return acts[-1]
        feats = from_blocked_bf16(acts[-1], ctx.channels[-1])
        return feats.view(sat.shape[0], -1)

    @staticmethod
    def backward(ctx, g):
        n = ctx.n_layers
        saved = ctx.saved_tensors
        wb, acts = saved[: 2 * n], saved[2 * n:]
        ch = ctx.channels
        B, _, To, Ho, Wo, _ = acts[-1].shape
        if ctx.link is not None:
            if "gz" not in ctx.link:
                raise RuntimeError("EncoderBf16Fn: the blocked activation must be consumed by HeadBf16Fn (private protocol)")
            gz_pad, gzw = ctx.link.pop("gz")
        else:
            g = g.contiguous().view(B, ch[-1], To, Ho, Wo)
            gz_pad = to_blocked_bf16(g, pad=2) if n > 1 else None
            gzw = to_gzw_bf16(g)
        grads: List[Optional[torch.Tensor]] = [None] * (2 * n)
        pad_t = ctx.pad_t
        # per layer: the row-step weight gradient (round 2: tensor maps straight into the operand layout, reads the same
        # zero-padded gradient tensor as the data gradient) or, for planes wider than 64, the round-1 tile kernel with its
        # own "gzw" operand layout
        rows = [wgrad_bf16_rows_supported(ch[l], ch[l + 1], acts[l].shape[3], acts[l].shape[4]) for l in range(n)]
        gz_blk, gz_blk_pad = gz_pad, 2
        if gz_blk is None and rows[n - 1]:
            gz_blk, gz_blk_pad = to_blocked_bf16(g, pad=0), 0
        for l in range(n - 1, -1, -1):
            if rows[l]:
                dw, db = conv3d_wgrad_bf16_rows(acts[l], gz_blk, ch[l], ch[l + 1], gz_pad=gz_blk_pad, pad_t=pad_t)
            else:
                dw, db = conv3d_wgrad_bf16(acts[l], gzw, ch[l], ch[l + 1], pad_t=pad_t)
            grads[2 * l], grads[2 * l + 1] = dw, db
            if l > 0:
                need_gzw = not rows[l - 1]
                # the gradient w.r.t. layer 0's output only feeds layer 0's weight gradient: no zero border needed
                out_pad = 2 if l > 1 else 0
                res = conv3d_dgrad_bf16(gz_pad, wb[2 * l], acts[l], out_pad=out_pad, also_gzw=need_gzw, persistent=True, pad_t=pad_t)
                gx, gzw = res if need_gzw else (res, None)
                gz_blk, gz_blk_pad = gx, out_pad
                gz_pad = gx if l > 1 else None
        return (None, None, None, None, *grads)


class HeadFn(torch.autograd.Function):
    """Fully connected head.  forward(feats, pv_hist|None, nwp|None, w1,b1,w2,b2,wn|None,bn|None,w3,b3,w4,b4).

    ``pv_hist``: [B, nt, ns] view with unit stride on the last axis (NaNs allowed: nan_to_num fused);
    ``nwp``: [B, NNWP] contiguous.  The returned gradient w.r.t. ``feats`` carries the ReLU mask
    ``feats > 0`` (see ``EncoderFn``)."""

    @staticmethod
    def _desc(feats, pv_hist, nwp, w1, b1, w2, b2, wn, bn, w3, b3, w4, b4):
        h = _lib.Head()
        h.struct_size = C.sizeof(_lib.Head)
        h.B, h.K1 = feats.shape[0], feats.shape[1]
        h.F1, h.F2, h.F3, h.FO = w1.shape[0], w2.shape[0], w3.shape[0], w4.shape[0]
        if pv_hist is not None:
            h.NPV = pv_hist.shape[1] * pv_hist.shape[2]
            h.pv_ns = pv_hist.shape[2]
            h.pv_sb, h.pv_st = pv_hist.stride(0), pv_hist.stride(1)
            h.pv = pv_hist.data_ptr()
        else:
            h.NPV, h.pv_ns, h.pv_sb, h.pv_st, h.pv = 0, 0, 0, 0, None
        if nwp is not None:
            h.NNWP, h.FNWP = nwp.shape[1], wn.shape[0]
            h.nwp, h.wn, h.bn = nwp.data_ptr(), wn.data_ptr(), bn.data_ptr()
        else:
            h.NNWP, h.FNWP = 0, 0
        ncat = h.F2 + h.NPV + (h.FNWP if nwp is not None else 0)
        if tuple(w1.shape) != (h.F1, h.K1) or tuple(w2.shape) != (h.F2, h.F1) or tuple(w3.shape) != (h.F3, ncat) \
                or tuple(w4.shape) != (h.FO, h.F3) or (nwp is not None and tuple(wn.shape) != (h.FNWP, h.NNWP)):
            raise RuntimeError("head: weight shapes inconsistent with the inputs")
        h.w1, h.b1, h.w2, h.b2 = w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr()
        h.w3, h.b3, h.w4, h.b4 = w3.data_ptr(), b3.data_ptr(), w4.data_ptr(), b4.data_ptr()
        h.x = feats.data_ptr()
        return h, ncat

    @staticmethod
    def forward(ctx, feats, pv_hist, nwp, w1, b1, w2, b2, wn, bn, w3, b3, w4, b4):
        L = _lib.load()
        _need_cuda(feats, "features", torch.float32)
        for t, nm in ((w1, "fc1.weight"), (b1, "fc1.bias"), (w2, "fc2.weight"), (b2, "fc2.bias"), (w3, "fc3.weight"),
                      (b3, "fc3.bias"), (w4, "fc4.weight"), (b4, "fc4.bias")):
            _need_cuda(t, nm, torch.float32)
        if pv_hist is not None:
            if not pv_hist.is_cuda or pv_hist.dtype != torch.float32 or pv_hist.stride(2) != 1:
                raise RuntimeError("head: pv history must be a CUDA fp32 [B,nt,ns] view with unit last stride")
        if nwp is not None:
            _need_cuda(nwp, "nwp", torch.float32)
            _need_cuda(wn, "fc_nwp.weight", torch.float32)
            _need_cuda(bn, "fc_nwp.bias", torch.float32)
        h, ncat = HeadFn._desc(feats, pv_hist, nwp, w1, b1, w2, b2, wn, bn, w3, b3, w4, b4)
        if w1.requires_grad and getattr(w1, "_pvb_shadow", None) is not None:
            w1._pvb_shadow.trained_through = False  # this step reads the fp32 master, not the shadow
        dev, B = feats.device, feats.shape[0]
        h1 = torch.empty((B, h.F1), dtype=torch.float32, device=dev)
        cat = torch.empty((B, ncat), dtype=torch.float32, device=dev)
        h3 = torch.empty((B, h.F3), dtype=torch.float32, device=dev)
        out = torch.empty((B, h.FO), dtype=torch.float32, device=dev)
        h.h1, h.cat, h.h3, h.out = h1.data_ptr(), cat.data_ptr(), h3.data_ptr(), out.data_ptr()
        ws = _workspace("head", L.pvb200_head_fwd_workspace_bytes(B, h.F1, h.K1), dev)
        h.workspace, h.workspace_bytes = ws.data_ptr(), ws.numel()
        with _timed("head_fwd_f32", 2.0 * B * h.F1 * h.K1, 4.0 * (B * h.K1 + h.F1 * h.K1)):
            rc = L.pvb200_head_fwd_f32(C.byref(h), _stream())
        _lib.check(rc, "head_fwd")
        ctx.save_for_backward(feats, pv_hist, nwp, w1, b1, w2, b2, wn, bn, w3, b3, w4, b4, h1, cat, h3)
        return out

    @staticmethod
    def backward(ctx, g_out):
        L = _lib.load()
        feats, pv_hist, nwp, w1, b1, w2, b2, wn, bn, w3, b3, w4, b4, h1, cat, h3 = ctx.saved_tensors
        h, ncat = HeadFn._desc(feats, pv_hist, nwp, w1, b1, w2, b2, wn, bn, w3, b3, w4, b4)
        dev, B = feats.device, feats.shape[0]
        g_out = g_out.contiguous()
        e = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)  # noqa: E731
        g_h3, g_cat, g_h1 = e(B, h.F3), e(B, ncat), e(B, h.F1)
        need_gx = ctx.needs_input_grad[0]
        g_x = e(B, h.K1) if need_gx else None
        dw1, db1, dw2, db2 = torch.empty_like(w1), torch.empty_like(b1), torch.empty_like(w2), torch.empty_like(b2)
        dw3, db3, dw4, db4 = torch.empty_like(w3), torch.empty_like(b3), torch.empty_like(w4), torch.empty_like(b4)
        dwn = torch.empty_like(wn) if nwp is not None else None
        dbn = torch.empty_like(bn) if nwp is not None else None
        h.h1, h.cat, h.h3 = h1.data_ptr(), cat.data_ptr(), h3.data_ptr()
        h.g_out, h.g_h3, h.g_cat, h.g_h1, h.g_x = g_out.data_ptr(), g_h3.data_ptr(), g_cat.data_ptr(), g_h1.data_ptr(), _p(g_x)
        h.dw1, h.db1, h.dw2, h.db2 = dw1.data_ptr(), db1.data_ptr(), dw2.data_ptr(), db2.data_ptr()
        h.dw3, h.db3, h.dw4, h.db4 = dw3.data_ptr(), db3.data_ptr(), dw4.data_ptr(), db4.data_ptr()
        h.dwn, h.dbn = _p(dwn), _p(dbn)
        # fc1 wgrad: read x, write dW1; fc1 dgrad: read W1 + x (mask), write gx
        with _timed("head_bwd_f32", (4.0 if need_gx else 2.0) * B * h.F1 * h.K1,
                    4.0 * (B * h.K1 + h.F1 * h.K1) + (4.0 * (h.F1 * h.K1 + 2 * B * h.K1) if need_gx else 0.0)):
            rc = L.pvb200_head_bwd_f32(C.byref(h), _stream())
        _lib.check(rc, "head_bwd")
        return g_x, None, None, dw1, db1, dw2, db2, dwn, dbn, dw3, db3, dw4, db4


_persistent: Dict[Tuple, torch.Tensor] = {}


def _zero_bordered(name: str, shape: Tuple[int, ...], device: torch.device, dtype=torch.bfloat16) -> torch.Tensor:
    """A buffer allocated ONCE with zeros and reused every step: the kernels that fill it only ever write the
    interior (valid positions), so the zero border / wrap columns never need a per-step memset."""
    key = (name, tuple(shape), device.index, dtype)
    buf = _persistent.get(key)
    if buf is None:
        buf = torch.zeros(shape, dtype=dtype, device=device)
        _persistent[key] = buf
    return buf


class Fc1Shadow:
    """bf16 shadow of ``fc1.weight`` in the tensor-core layout [Cg*T*H*W][128][8], refreshed lazily.

    The shadow is valid for one (storage, version) of the master weight: ``_version`` changes when torch writes the
    parameter in place (``load_state_dict``, manual edits), ``_pvb_gen`` when ``FusedAdam`` updates it through the C ABI
    (which torch cannot see).  ``FusedAdam`` refreshes the shadow itself, inside the Adam pass, when it is attached."""

    def __init__(self):
        self.buf: Optional[torch.Tensor] = None
        self.geom: Optional[Tuple[int, int, int, int]] = None
        self.key = None
        self.ready: Optional[torch.cuda.Event] = None  # set while a sharded refresh is in flight on the comm stream
        # did the last grad-enabled forward read fc1 THROUGH this shadow?  Only then may data parallelism shard the
        # optimiser of the master weight by rows (dp.GradientExchange._can_shard): a forward through the fp32 head reads
        # the master itself, whose foreign rows would be stale.
        self.trained_through = False
        self._send: Optional[torch.Tensor] = None
        self._gathered: Optional[torch.Tensor] = None

    @staticmethod
    def _key(w1: torch.Tensor):
        return (w1.data_ptr(), w1._version, getattr(w1, "_pvb_gen", 0))

    def ensure(self, w1: torch.Tensor, F1: int, Cg: int, T: int, H: int, W: int) -> torch.Tensor:
        L = _lib.load()
        geom = (Cg, T, H, W)
        if self.buf is None or self.geom != geom or self.buf.device != w1.device:
            self.buf = torch.empty(L.pvb200_fc1_bf16_shadow_bytes(Cg, T, H, W), dtype=torch.uint8, device=w1.device)
            self.geom, self.key = geom, None
        if self.ready is not None:  # refreshed by the sharded optimiser step on the communication stream
            torch.cuda.current_stream(w1.device).wait_event(self.ready)
            self.ready = None
        if self.key != self._key(w1):
            with _timed("fc1_make_shadow_bf16", 0.0, 6.0 * w1.numel()):
                rc = L.pvb200_fc1_make_shadow_bf16(_p(w1), _p(self.buf), F1, Cg, T, H, W, _stream())
            _lib.check(rc, "fc1_make_shadow_bf16")
            self.key = self._key(w1)
        return self.buf

    def adam_step(self, w1: "torch.nn.Parameter", grad: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, lr: float,
                  beta1: float, beta2: float, eps: float, step: int, grad_scale: float) -> bool:
        """Adam update of ``w1`` that rewrites the shadow in the same pass; False if no shadow geometry is known yet."""
        if self.buf is None or self.geom is None or self.buf.device != w1.device:
            return False
        L = _lib.load()
        Cg, T, H, W = self.geom
        if w1.shape[1] != Cg * 8 * T * H * W or w1.shape[0] > 128:
            return False
        for t, nm in ((w1.data, "param"), (grad, "grad"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq")):
            _need_cuda(t, nm, torch.float32)
        with _timed("adam_fc1_shadow", 0.0, 30.0 * w1.numel()):
            rc = L.pvb200_adam_fc1_shadow(_p(w1), _p(grad), _p(exp_avg), _p(exp_avg_sq), _p(self.buf), w1.shape[0], Cg, T, H, W,
                                          lr, beta1, beta2, eps, step, grad_scale, _stream())
        _lib.check(rc, "adam_fc1_shadow")
        w1._pvb_gen = getattr(w1, "_pvb_gen", 0) + 1
        self.key = self._key(w1)
        return True


    def adam_step_sharded(self, w1: "torch.nn.Parameter", grad: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor,
                          lr: float, beta1: float, beta2: float, eps: float, step: int, grad_scale: float, rank: int, world: int,
                          group, comm_stream: "torch.cuda.Stream") -> bool:
        """Data-parallel step with the optimiser sharded by output feature: this rank updates rows
        [rank*F1/world, (rank+1)*F1/world) of ``w1`` (their gradient rows must already hold the reduce-scattered sum),
        the bf16 copies of all ranks' rows are all-gathered on ``comm_stream`` and interleaved into the shadow there, so
        only the 1/world Adam pass sits on the compute stream; the next forward waits on ``self.ready``.
        The fp32 master rows owned by OTHER ranks go stale (``GradientExchange.gather_master_weights`` refreshes them)."""
        import torch.distributed as dist

        if self.buf is None or self.geom is None or self.buf.device != w1.device:
            return False
        L = _lib.load()
        Cg, T, H, W = self.geom
        F1 = w1.shape[0]
        if w1.shape[1] != Cg * 8 * T * H * W or F1 > 128 or F1 % world != 0:
            return False
        nrows = F1 // world
        KG = Cg * T * H * W
        dev = w1.device
        if self._send is None or self._send.numel() != KG * nrows * 8 or self._send.device != dev:
            self._send = torch.empty(KG * nrows * 8, dtype=torch.bfloat16, device=dev)
            self._gathered = torch.empty(world * KG * nrows * 8, dtype=torch.bfloat16, device=dev)
        for t, nm in ((w1.data, "param"), (grad, "grad"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq")):
            _need_cuda(t, nm, torch.float32)
        with _timed("adam_fc1_shadow_rows", 0.0, 30.0 * w1.numel() / world):
            rc = L.pvb200_adam_fc1_shadow_rows(_p(w1), _p(grad), _p(exp_avg), _p(exp_avg_sq), _p(self._send), F1, Cg, T, H, W,
                                               rank * nrows, nrows, lr, beta1, beta2, eps, step, grad_scale, _stream())
        _lib.check(rc, "adam_fc1_shadow_rows")
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(dev))
        comm_stream.wait_event(done)
        with torch.cuda.stream(comm_stream):
            dist.all_gather_into_tensor(self._gathered, self._send, group=group)
            rc = L.pvb200_fc1_shadow_from_shards(_p(self._gathered), _p(self.buf), world, nrows, Cg, T, H, W,
                                                 comm_stream.cuda_stream)
            _lib.check(rc, "fc1_shadow_from_shards")
            self.ready = torch.cuda.Event()
            self.ready.record(comm_stream)
        w1._pvb_gen = getattr(w1, "_pvb_gen", 0) + 1
        self.key = self._key(w1)
        return True


class HeadBf16Fn(torch.autograd.Function):
    """FC head of the bf16 mode: fc1 as weight-streaming tcgen05 GEMMs over a bf16 shadow of the fp32 master weight,
    the rest of the head (fc2, concat, fc_nwp, fc3, fc4) in the fused fp32 tail kernels.

    forward(link, act, pv_hist|None, nwp|None, w1,b1,w2,b2,wn|None,bn|None,w3,b3,w4,b4); ``act`` is the last conv
    activation in blocked bf16 [B,Cg,T,H,W,8] from ``EncoderBf16Fn``.  The gradient w.r.t. ``act`` is handed to the
    encoder through ``link["gz"]`` (ReLU mask fused, both layouts); the tensor returned to autograd is a placeholder."""

    @staticmethod
    def forward(ctx, link, act, pv_hist, nwp, w1, b1, w2, b2, wn, bn, w3, b3, w4, b4):
        L = _lib.load()
        _need_cuda(act, "activation", torch.bfloat16)
        B, Cg, T, H, W, _ = act.shape
        feats_stub = act.view(B, -1)  # only its shape is used by the descriptor
        h, ncat = HeadFn._desc(feats_stub, pv_hist, nwp, w1, b1, w2, b2, wn, bn, w3, b3, w4, b4)
        h.x = None
        dev = act.device
        shadow_state = link.get("shadow") if link is not None else None
        if shadow_state is None:
            shadow_state = Fc1Shadow()
        shadow = shadow_state.ensure(w1, h.F1, Cg, T, H, W)
        if w1.requires_grad:
            shadow_state.trained_through = True
        ctx.shadow = shadow  # the very buffer the forward used (kept alive for the data gradient)
        S = int(L.pvb200_fc1_fwd_bf16_splits())
        partial = _workspace("head", S * B * h.F1 * 4, dev)
        with _timed("fc1_fwd_bf16", 2.0 * B * h.F1 * h.K1, 2.0 * (B * h.K1 + 128 * h.K1)):
            rc = L.pvb200_fc1_fwd_bf16(_p(act), _p(shadow), _p(partial), B, h.F1, Cg, T, H, W, _stream())
        _lib.check(rc, "fc1_fwd_bf16")
        h1 = torch.empty((B, h.F1), dtype=torch.float32, device=dev)
        cat = torch.empty((B, ncat), dtype=torch.float32, device=dev)
        h3 = torch.empty((B, h.F3), dtype=torch.float32, device=dev)
        out = torch.empty((B, h.FO), dtype=torch.float32, device=dev)
        h.h1, h.cat, h.h3, h.out = h1.data_ptr(), cat.data_ptr(), h3.data_ptr(), out.data_ptr()
        h.workspace, h.workspace_bytes = partial.data_ptr(), partial.numel()
        _lib.check(L.pvb200_head_tail_fwd_f32(C.byref(h), S, _stream()), "head_tail_fwd")
        ctx.save_for_backward(act, pv_hist, nwp, w1, b1, w2, b2, wn, bn, w3, b3, w4, b4, h1, cat, h3)
        ctx.link = link
        return out

    @staticmethod
    def backward(ctx, g_out):
        L = _lib.load()
        act, pv_hist, nwp, w1, b1, w2, b2, wn, bn, w3, b3, w4, b4, h1, cat, h3 = ctx.saved_tensors
        B, Cg, T, H, W, _ = act.shape
        h, ncat = HeadFn._desc(act.view(B, -1), pv_hist, nwp, w1, b1, w2, b2, wn, bn, w3, b3, w4, b4)
        h.x = None
        dev = act.device
        g_out = g_out.contiguous()
        e = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)  # noqa: E731
        g_h3, g_cat, g_h1 = e(B, h.F3), e(B, ncat), e(B, h.F1)
        dw1, db1, dw2, db2 = torch.empty_like(w1), torch.empty_like(b1), torch.empty_like(w2), torch.empty_like(b2)
        dw3, db3, dw4, db4 = torch.empty_like(w3), torch.empty_like(b3), torch.empty_like(w4), torch.empty_like(b4)
        dwn = torch.empty_like(wn) if nwp is not None else None
        dbn = torch.empty_like(bn) if nwp is not None else None
        h.h1, h.cat, h.h3 = h1.data_ptr(), cat.data_ptr(), h3.data_ptr()
        h.g_out, h.g_h3, h.g_cat, h.g_h1, h.g_x = g_out.data_ptr(), g_h3.data_ptr(), g_cat.data_ptr(), g_h1.data_ptr(), None
        h.dw1, h.db1, h.dw2, h.db2 = None, db1.data_ptr(), dw2.data_ptr(), db2.data_ptr()
        h.dw3, h.db3, h.dw4, h.db4 = dw3.data_ptr(), db3.data_ptr(), dw4.data_ptr(), db4.data_ptr()
        h.dwn, h.dbn = _p(dwn), _p(dbn)
        _lib.check(L.pvb200_head_tail_bwd_f32(C.byref(h), _stream()), "head_tail_bwd")
        with _timed("fc1_wgrad_bf16", 2.0 * B * h.F1 * h.K1, 2.0 * B * h.K1 + 4.0 * h.F1 * h.K1):
            rc = L.pvb200_fc1_wgrad_bf16(_p(g_h1), _p(act), _p(dw1), B, h.F1, Cg, T, H, W, _stream())
        _lib.check(rc, "fc1_wgrad_bf16")
        g_act = None
        if ctx.needs_input_grad[1]:
            QP = int(L.pvb200_conv3d_wgrad_bf16_gz_plane(H + 2, W + 2))
            gz_pad = _zero_bordered("gz_pad_head", (B, Cg, T + 4, H + 4, W + 4, 8), dev)
            need_gzw = ctx.link is None or ctx.link.get("need_gzw", True)
            gzw = _zero_bordered(f"gzw_head_{H}x{W}", (B, Cg, T, QP, 8), dev) if need_gzw else None
            shadow = ctx.shadow
            with _timed("fc1_dgrad_bf16", 2.0 * B * h.F1 * h.K1, 2.0 * (128 * h.K1 + (4 if need_gzw else 3) * B * h.K1)):
                rc = L.pvb200_fc1_dgrad_bf16(_p(g_h1), _p(shadow), _p(act), _p(gz_pad), _p(gzw), B, h.F1, Cg, T, H, W, _stream())
            _lib.check(rc, "fc1_dgrad_bf16")
            if ctx.link is None:
                raise RuntimeError("HeadBf16Fn: no link to hand the activation gradient to EncoderBf16Fn")
            ctx.link["gz"] = (gz_pad, gzw)
            g_act = _zero_bordered("g_act_placeholder", tuple(act.shape), act.device, act.dtype)  # zeros, never written: the real gradient travels through the link
        return None, g_act, None, None, dw1, db1, dw2, db2, dwn, dbn, dw3, db3, dw4, db4


class TowerFn(torch.autograd.Function):
    """Conv3d + ReLU stack with time padding (pad_t, 0, 0) in fp32: the satellite / NWP towers of conv3d_sat_nwp
    (model_sat_nwp.py:85-100,130-143,187-190,237-240).  forward(pad_t, x, mean|None, std|None, w0, b0, ...) -> the last
    activation flattened to [B, C*T*H*W] (NCDHW order, model_sat_nwp.py:192).  Same private protocol as ``EncoderFn``:
    the incoming gradient already carries the ReLU mask of the last layer (``LinearFn(mask_input=True)``).  ``x`` needs no
    gradient (it is data)."""

    @staticmethod
    def forward(ctx, pad_t, x, mean, std, *wb):
        n_layers = len(wb) // 2
        if x.dtype == torch.int16:
            x = sat_normalise(x, mean, std)
        elif x.dtype != torch.float32:
            raise RuntimeError(f"TowerFn: input must be int16 or float32, got {x.dtype}")
        x0 = x.contiguous()
        acts = []
        a = x0
        for l in range(n_layers):
            a = conv3d_fwd(a, wb[2 * l], wb[2 * l + 1], relu=True, pad_t=pad_t)
            acts.append(a)
        ctx.save_for_backward(x0, *wb, *acts)
        ctx.n_layers, ctx.pad_t = n_layers, pad_t
        return acts[-1].view(x0.shape[0], -1)

    @staticmethod
    def backward(ctx, g):
        n, pad_t = ctx.n_layers, ctx.pad_t
        saved = ctx.saved_tensors
        x0 = saved[0]
        wb = saved[1: 1 + 2 * n]
        acts = saved[1 + 2 * n:]
        gz = g.contiguous().view(acts[-1].shape)
        grads: List[Optional[torch.Tensor]] = [None] * (2 * n)
        for l in range(n - 1, -1, -1):
            xin = x0 if l == 0 else acts[l - 1]
            grads[2 * l], grads[2 * l + 1] = conv3d_wgrad(xin, gz, pad_t=pad_t)
            if l > 0:
                gz = conv3d_dgrad(gz, wb[2 * l], acts[l - 1], acts[l - 1].shape, pad_t=pad_t)
        return (None, None, None, None, *grads)


class Conv3dMaxPoolFn(torch.autograd.Function):
    """MaxPool3d(3, stride (1,2,2), pad 1)(Conv3d(x, w, b, padding 1)) -- the Conv3dMaxPool block of the Perceiver hybrid
    (perceiver_conv3d_nwp_sat.py:42-57; no ReLU between the two).  forward(x [B,Ci,T,H,W] fp32, w, b) -> [B,Co,T,Ho,Wo];
    x may require grad (the block is a front-end, but nothing forces it to be the first layer)."""

    @staticmethod
    def forward(ctx, x, w, b):
        L = _lib.load()
        x = x.contiguous()
        y = conv3d_fwd(x, w, b, relu=False, pad_t=1, pad_hw=1)
        B, Co, T, H, W = y.shape
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        out = torch.empty((B, Co, T, Ho, Wo), dtype=torch.float32, device=x.device)
        arg = torch.empty((B, Co, T, Ho, Wo), dtype=torch.int32, device=x.device)
        with _timed("maxpool3d_fwd_f32", 0.0, 4.0 * y.numel() + 8.0 * out.numel()):
            rc = L.pvb200_maxpool3d_fwd_f32(_p(y), _p(out), _p(arg), B * Co, T, H, W, _stream())
        _lib.check(rc, "maxpool3d_fwd")
        ctx.save_for_backward(x, w, arg)
        ctx.conv_shape = tuple(y.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        L = _lib.load()
        x, w, arg = ctx.saved_tensors
        B, Co, T, H, W = ctx.conv_shape
        g = g.contiguous()
        gz = torch.empty(ctx.conv_shape, dtype=torch.float32, device=g.device)
        with _timed("maxpool3d_bwd_f32", 0.0, 4.0 * gz.numel() + 8.0 * g.numel()):
            rc = L.pvb200_maxpool3d_bwd_f32(_p(g), _p(arg), _p(gz), B * Co, T, H, W, _stream())
        _lib.check(rc, "maxpool3d_bwd")
        dw, db = conv3d_wgrad(x, gz, pad_t=1, pad_hw=1)
        gx = conv3d_dgrad(gz, w, None, x.shape, pad_t=1, pad_hw=1) if ctx.needs_input_grad[0] else None
        return gx, dw, db


class Fc1Bf16Fn(torch.autograd.Function):
    """relu(fc1(flatten(act))) on the tensor cores for a tower of conv3d_sat_nwp in bf16 mode (model_sat_nwp.py:192-195,
    242-244): weight-streaming tcgen05 GEMMs over the bf16 shadow of the fp32 master weight, exactly as ``HeadBf16Fn``
    does for the single-tower model.  forward(link, act [B,Cg,T,H,W,8] bf16, w1 [F1,K1], b1) -> [B,F1] fp32.  The
    gradient w.r.t. ``act`` is handed to ``EncoderBf16Fn`` through ``link["gz"]`` (ReLU mask fused, both layouts)."""

    @staticmethod
    def forward(ctx, link, act, w1, b1):
        L = _lib.load()
        _need_cuda(act, "activation", torch.bfloat16)
        _need_cuda(w1, "fc1.weight", torch.float32)
        _need_cuda(b1, "fc1.bias", torch.float32)
        B, Cg, T, H, W, _ = act.shape
        F1 = w1.shape[0]
        if w1.shape[1] != Cg * 8 * T * H * W:
            raise RuntimeError(f"Fc1Bf16Fn: weight {tuple(w1.shape)} does not match the activation {tuple(act.shape)}")
        shadow = link["shadow"].ensure(w1, F1, Cg, T, H, W)
        if w1.requires_grad:
            link["shadow"].trained_through = True
        ctx.shadow = shadow
        S = int(L.pvb200_fc1_fwd_bf16_splits())
        partial = _workspace("head", S * B * F1 * 4, act.device)
        K1 = w1.shape[1]
        with _timed("fc1_fwd_bf16", 2.0 * B * F1 * K1, 2.0 * (B * K1 + 128 * K1)):
            rc = L.pvb200_fc1_fwd_bf16(_p(act), _p(shadow), _p(partial), B, F1, Cg, T, H, W, _stream())
        _lib.check(rc, "fc1_fwd_bf16")
        h1 = torch.empty((B, F1), dtype=torch.float32, device=act.device)
        _lib.check(L.pvb200_linear_finish_f32(_p(partial), S, _p(b1), _p(h1), F1, B, F1, 1, _stream()), "linear_finish")
        ctx.save_for_backward(act, w1, h1)
        ctx.link = link
        return h1

    @staticmethod
    def backward(ctx, g):
        L = _lib.load()
        act, w1, h1 = ctx.saved_tensors
        B, Cg, T, H, W, _ = act.shape
        F1, K1 = w1.shape
        dev = act.device
        g = g.contiguous()
        g_pre = torch.empty((B, F1), dtype=torch.float32, device=dev)
        db1 = torch.empty((F1,), dtype=torch.float32, device=dev)
        _lib.check(L.pvb200_linear_gpre_f32(_p(g), F1, _p(h1), F1, _p(g_pre), _p(db1), B, F1, _stream()), "linear_gpre")
        dw1 = torch.empty_like(w1)
        with _timed("fc1_wgrad_bf16", 2.0 * B * F1 * K1, 2.0 * B * K1 + 4.0 * F1 * K1):
            rc = L.pvb200_fc1_wgrad_bf16(_p(g_pre), _p(act), _p(dw1), B, F1, Cg, T, H, W, _stream())
        _lib.check(rc, "fc1_wgrad_bf16")
        g_act = None
        if ctx.needs_input_grad[1]:
            tag = ctx.link.get("tag", "tower")
            QP = int(L.pvb200_conv3d_wgrad_bf16_gz_plane(H + 2, W + 2))
            gz_pad = _zero_bordered("gz_pad_" + tag, (B, Cg, T + 4, H + 4, W + 4, 8), dev)
            gzw = _zero_bordered(f"gzw_{tag}_{H}x{W}", (B, Cg, T, QP, 8), dev) if ctx.link.get("need_gzw", True) else None
            with _timed("fc1_dgrad_bf16", 2.0 * B * F1 * K1, 2.0 * (128 * K1 + 4 * B * K1)):
                rc = L.pvb200_fc1_dgrad_bf16(_p(g_pre), _p(ctx.shadow), _p(act), _p(gz_pad), _p(gzw), B, F1, Cg, T, H, W, _stream())
            _lib.check(rc, "fc1_dgrad_bf16")
            ctx.link["gz"] = (gz_pad, gzw)
            g_act = _zero_bordered("g_act_placeholder", tuple(act.shape), act.device, act.dtype)  # zeros, never written: the real gradient travels through the link
        return None, g_act, dw1, db1


class LinearFn(torch.autograd.Function):
    """y = act(x W^T + b) through ``pvb200_linear_{fwd,bwd}_f32`` (nn.Linear layers of model_sat_nwp.py:102-172).

    forward(x [B,K], w [N,K], b [N], relu, mask_input).  ``mask_input``: x is a post-ReLU activation whose producer
    expects the gradient of its PRE-activation (the conv towers): gx is multiplied by (x > 0)."""

    @staticmethod
    def forward(ctx, x, w, b, relu, mask_input):
        L = _lib.load()
        for t, nm in ((x, "input"), (w, "weight"), (b, "bias")):
            _need_cuda(t, nm, torch.float32)
        if x.dim() != 2 or x.stride(1) != 1:
            raise RuntimeError("LinearFn: input must be a 2-D tensor with unit column stride")
        B, K = x.shape
        N = w.shape[0]
        if tuple(w.shape) != (N, K) or tuple(b.shape) != (N,):
            raise RuntimeError(f"LinearFn: weight {tuple(w.shape)} / bias {tuple(b.shape)} do not match input features {K}")
        if w.requires_grad and getattr(w, "_pvb_shadow", None) is not None:
            w._pvb_shadow.trained_through = False  # this step reads the fp32 master, not the shadow
        w = w.contiguous()
        y = torch.empty((B, N), dtype=torch.float32, device=x.device)
        ws = _workspace("linear", L.pvb200_linear_workspace_bytes(B, N, K), x.device)
        with _timed(f"linear_fwd_f32[K={K}]", 2.0 * B * N * K, 4.0 * (B * K + N * K)):
            rc = L.pvb200_linear_fwd_f32(_p(x), x.stride(0), _p(w), _p(b), _p(y), N, B, K, N, int(relu), _p(ws), ws.numel(), _stream())
        _lib.check(rc, "linear_fwd")
        ctx.save_for_backward(x, w, y)
        ctx.relu, ctx.mask_input = bool(relu), bool(mask_input)
        return y

    @staticmethod
    def backward(ctx, g):
        L = _lib.load()
        x, w, y = ctx.saved_tensors
        B, K = x.shape
        N = w.shape[0]
        g = g.contiguous()
        dw = torch.empty_like(w)
        db = torch.empty((N,), dtype=torch.float32, device=x.device)
        gx = torch.empty((B, K), dtype=torch.float32, device=x.device) if ctx.needs_input_grad[0] else None
        ws = _workspace("linear", L.pvb200_linear_workspace_bytes(B, N, K), x.device)
        with _timed(f"linear_bwd_f32[K={K}]", 4.0 * B * N * K, 4.0 * (2 * B * K + 2 * N * K)):
            rc = L.pvb200_linear_bwd_f32(_p(x), x.stride(0), _p(w), _p(y) if ctx.relu else None, N, _p(g), N, _p(gx), K,
                                         int(ctx.mask_input), _p(dw), _p(db), B, K, N, _p(ws), ws.numel(), _stream())
        _lib.check(rc, "linear_bwd")
        return gx, dw, db, None, None


class EmbeddingFn(torch.autograd.Function):
    """table[ids] (nn.Embedding(940, 16) of model_sat_nwp.py:146-149,252-260) with a dense deterministic gradient.
    forward(table [V,D], ids int32 [B]) -> [B, D].  ``validate_ids``: raise IndexError for ids outside [0, V) as
    nn.Embedding does (one 8-byte device->host read per call; the reference round-trips the ids through the CPU anyway,
    model_sat_nwp.py:257-258).  With False the kernels return zeros / drop the gradient for such ids."""

    validate_ids = True

    @staticmethod
    def forward(ctx, table, ids):
        L = _lib.load()
        _need_cuda(table, "embedding table", torch.float32)
        _need_cuda(ids, "ids", torch.int32)
        V, D = table.shape
        B = ids.shape[0]
        if EmbeddingFn.validate_ids and B > 0:
            lo, hi = (int(v) for v in torch.stack((ids.min(), ids.max())).tolist())
            if lo < 0 or hi >= V:
                raise IndexError(f"embedding: index out of range (ids span [{lo}, {hi}], table has {V} rows)")
        y = torch.empty((B, D), dtype=torch.float32, device=table.device)
        _lib.check(L.pvb200_embedding_fwd_f32(_p(table.contiguous()), _p(ids), _p(y), D, B, V, D, _stream()), "embedding_fwd")
        ctx.save_for_backward(ids)
        ctx.V, ctx.D = V, D
        return y

    @staticmethod
    def backward(ctx, g):
        L = _lib.load()
        (ids,) = ctx.saved_tensors
        g = g.contiguous()
        dtable = torch.empty((ctx.V, ctx.D), dtype=torch.float32, device=g.device)
        _lib.check(L.pvb200_embedding_bwd_f32(_p(g), ctx.D, _p(ids), _p(dtable), ids.shape[0], ctx.V, ctx.D, _stream()), "embedding_bwd")
        return dtable, None


def history_flatten(src: torch.Tensor, nt: int, ns: int) -> torch.Tensor:
    """src[:, :nt, :ns].nan_to_num(0).reshape(B, -1) (model_sat_nwp.py:207-232); data, no gradient."""
    L = _lib.load()
    if not src.is_cuda or src.dtype != torch.float32:
        raise RuntimeError(f"predict_pv_yield_b200: 'history' must be a float32 CUDA tensor (got {src.dtype} on {src.device}); "
                           "there is no CPU fallback")
    if src.dim() != 3 or src.stride(2) != 1 or src.shape[1] < nt or src.shape[2] < ns:
        raise RuntimeError(f"history_flatten: history {tuple(src.shape)} has no [:, :{nt}, :{ns}] block")
    B = src.shape[0]
    out = torch.empty((B, nt * ns), dtype=torch.float32, device=src.device)
    rc = L.pvb200_history_flatten_f32(_p(src), src.stride(0), src.stride(1), _p(out), nt * ns, B, nt, ns, _stream())
    _lib.check(rc, "history_flatten")
    return out


def validation_results(y_hat: torch.Tensor, y: torch.Tensor, capacity: Optional[torch.Tensor]):
    """(out [3,B,FO] = forecast MW, actual MW, capacity; horizon [2,FO] = per-horizon MSE, MAE) in one kernel.
    ``y`` / ``capacity`` are the strided views ``[:, -FO:, 0]`` of the batch tensors (base_model.py:95,222-227)."""
    L = _lib.load()
    _need_cuda(y_hat, "y_hat", torch.float32)
    B, FO = y_hat.shape
    for t, nm in ((y, "y"), (capacity, "capacity")):
        if t is not None and (not t.is_cuda or t.dtype != torch.float32 or tuple(t.shape) != (B, FO)):
            raise RuntimeError(f"validation_results: '{nm}' must be a float32 CUDA view of shape {(B, FO)}")
    out = torch.empty((3, B, FO), dtype=torch.float32, device=y_hat.device)
    horizon = torch.empty((2, FO), dtype=torch.float32, device=y_hat.device)
    rc = L.pvb200_validation_results_f32(_p(y_hat), y.data_ptr(), y.stride(0), y.stride(1),
                                         capacity.data_ptr() if capacity is not None else None,
                                         capacity.stride(0) if capacity is not None else 0,
                                         capacity.stride(1) if capacity is not None else 0, _p(out), _p(horizon), B, FO, _stream())
    _lib.check(rc, "validation_results")
    return out, horizon


class StepLossFn(torch.autograd.Function):
    """forward(y_hat [B,FO], y [B,FO] strided view, weights [FO]) -> losses [4] = nmae, mse, mse_exp, mae_exp.

    Only ``losses[0]`` (the L1 loss the reference returns, base_model.py:99,146) is differentiable."""

    @staticmethod
    def forward(ctx, y_hat, y, weights):
        L = _lib.load()
        _need_cuda(y_hat, "y_hat", torch.float32)
        if not y.is_cuda or y.dtype != torch.float32:
            raise RuntimeError("loss: target must be a CUDA fp32 tensor")
        if tuple(y.shape) != tuple(y_hat.shape):
            # the reference's F.mse_loss raises here too (base_model.py:95-98: batch_size class attr truncates y)
            raise RuntimeError(f"loss: target shape {tuple(y.shape)} != forecast shape {tuple(y_hat.shape)} "
                               "(set model.batch_size to the real batch size)")
        _need_cuda(weights, "loss weights", torch.float32)
        B, FO = y_hat.shape
        losses = torch.empty((4,), dtype=torch.float32, device=y_hat.device)
        rc = L.pvb200_l1_loss_fwd_f32(_p(y_hat), _p(y), y.stride(0), y.stride(1), _p(weights), _p(losses), B, FO, _stream())
        _lib.check(rc, "l1_loss_fwd")
        ctx.save_for_backward(y_hat, y)
        return losses

    @staticmethod
    def backward(ctx, g_losses):
        L = _lib.load()
        y_hat, y = ctx.saved_tensors
        B, FO = y_hat.shape
        g_losses = g_losses.contiguous()  # element 0 is the upstream gradient of nmae (device scalar, no sync)
        g = torch.empty_like(y_hat)
        rc = L.pvb200_l1_loss_bwd_f32(_p(y_hat), _p(y), y.stride(0), y.stride(1), _p(g_losses), _p(g), B, FO, _stream())
        _lib.check(rc, "l1_loss_bwd")
        return g, None, None
