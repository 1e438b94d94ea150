"""Launch each hot kernel ONCE at its full BASELINE shape (layer conv1, batch 32), after one warm-up launch each, for
`ncu --set full` captures:  ncu --set full --clock-control none -k regex:... -o gpurun_out/prof python tools/prof_kernels.py [f32|bf16|tc32]"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from predict_pv_yield_b200 import lib, ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "f32"
L = lib.load()
dev = torch.device("cuda:0")
B, Ci, T, S, Co = 32, 32, 17, 62, 32
g = torch.Generator(device=dev).manual_seed(0)
w = torch.randn(Co, Ci, 3, 3, 3, device=dev, generator=g) / (Ci * 27) ** 0.5
b = torch.randn(Co, device=dev, generator=g)
x = torch.relu(torch.randn(B, Ci, T, S, S, device=dev, generator=g))
gz = torch.randn(B, Co, T - 2, S - 2, S - 2, device=dev, generator=g)
for rep in range(2):  # the second pass is the one to capture (ncu -s <launches of pass 0>)
    if which == "tc32":  # fp32 mode on the tensor cores: 3xTF32 forward / data gradient (CTA pair), bf16x3 weight gradient
        am = torch.zeros(2, device=dev)
        xb4 = ops.to_blocked_f32(x, amax=am[0:1])
        gz4 = ops.to_blocked_f32(gz, pad=2, amax=am[1:2])
        f16 = len(sys.argv) > 2 and sys.argv[2] == "f16"  # two-way fp16 split instead of 3xTF32 / bf16x3
        ops.conv3d_fwd_tf32x3(xb4, w, b, want_blk=True, want_nc=False, amax_in=am[0:1] if f16 else None)
        ops.conv3d_dgrad_tf32x3(gz4, w, xb4, out_pad=2, want_blk=True, want_nc=False, amax_in=am[1:2] if f16 else None)
        ops.conv3d_wgrad_bf16x3(xb4, gz4, Ci, Co, gz_pad=2, amax=(am[0:1], am[1:2]) if f16 else None)
    elif which == "f32":
        ops.conv3d_fwd(x, w, b)
        ops.conv3d_dgrad(gz, w, x, x.shape)
        ops.conv3d_wgrad(x, gz)
    else:
        xb = ops.to_blocked_bf16(x)
        gzp = ops.to_blocked_bf16(gz, pad=2)
        gzw = ops.to_gzw_bf16(gz)
        ops.conv3d_fwd_bf16(xb, w, b)
        ops.conv3d_dgrad_bf16(gzp, w, xb)
        ops.conv3d_wgrad_bf16(xb, gzw, Ci, Co)
        ops.conv3d_wgrad_bf16_rows(xb, gzp, Ci, Co, gz_pad=2)  # round 2: row-step weight gradient (what the encoder runs)
        # fc1 (last activation 32 x 11 x 56 x 56) and the fused Adam + shadow pass
        Cg, Tf, Hf, Wf, F1 = 4, 11, 56, 56, 128
        K1 = Cg * 8 * Tf * Hf * Wf
        act = torch.relu(torch.randn(B, Cg, Tf, Hf, Wf, 8, device=dev, generator=g)).bfloat16()
        w1 = torch.randn(F1, K1, device=dev, generator=g) / K1 ** 0.5
        g1 = torch.randn(B, F1, device=dev, generator=g)
        st = torch.cuda.current_stream().cuda_stream
        shadow = torch.empty(L.pvb200_fc1_bf16_shadow_bytes(Cg, Tf, Hf, Wf), dtype=torch.uint8, device=dev)
        lib.check(L.pvb200_fc1_make_shadow_bf16(w1.data_ptr(), shadow.data_ptr(), F1, Cg, Tf, Hf, Wf, st))
        partial = torch.empty((L.pvb200_fc1_fwd_bf16_splits(), B, F1), device=dev)
        lib.check(L.pvb200_fc1_fwd_bf16(act.data_ptr(), shadow.data_ptr(), partial.data_ptr(), B, F1, Cg, Tf, Hf, Wf, st))
        dw = torch.empty_like(w1)
        lib.check(L.pvb200_fc1_wgrad_bf16(g1.data_ptr(), act.data_ptr(), dw.data_ptr(), B, F1, Cg, Tf, Hf, Wf, st))
        QP = L.pvb200_conv3d_wgrad_bf16_gz_plane(Hf + 2, Wf + 2)
        gz_pad = torch.zeros((B, Cg, Tf + 4, Hf + 4, Wf + 4, 8), dtype=torch.bfloat16, device=dev)
        gzw2 = torch.zeros((B, Cg, Tf, QP, 8), dtype=torch.bfloat16, device=dev)
        lib.check(L.pvb200_fc1_dgrad_bf16(g1.data_ptr(), shadow.data_ptr(), act.data_ptr(), gz_pad.data_ptr(), gzw2.data_ptr(), B, F1,
                                          Cg, Tf, Hf, Wf, st))
        m, v = torch.zeros_like(w1), torch.zeros_like(w1)
        lib.check(L.pvb200_adam_fc1_shadow(w1.data_ptr(), dw.data_ptr(), m.data_ptr(), v.data_ptr(), shadow.data_ptr(), F1, Cg, Tf, Hf,
                                           Wf, 5e-4, 0.9, 0.999, 1e-8, 1, 1.0, st))
        sat = torch.randint(0, 1024, (B, 12, 19, 64, 64), device=dev, dtype=torch.int16)
        ops.sat_normalise_blocked_bf16(sat, torch.ones(12, device=dev), torch.ones(12, device=dev))
    torch.cuda.synchronize()
