"""Per-kernel timings at the full layer shapes of BASELINE config 2/3 (CUDA events, L2 flushed between
iterations).  Diagnostic tool: python tools/bench_kernels.py [--batch 32] [--only bf16|f32]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from predict_pv_yield_b200 import lib, ops  # noqa: E402


def timeit(fn, iters=5, flush=None):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--only", default="")
    ap.add_argument("--size", type=int, default=64)
    args = ap.parse_args()
    lib.load()
    dev = torch.device("cuda:0")
    B = args.batch
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    layers = []
    T, S, C = 19, args.size, 12
    for l in range(4):
        layers.append((C, T, S))
        C, T, S = 32, T - 2, S - 2
    print(f"{'kernel':<34}{'ms':>9}{'TFLOP/s':>10}{'GB/s':>9}")
    if args.only in ("", "f32", "head"):
        K1, F1 = 32 * 11 * 56 * 56, 128
        feats = torch.relu(torch.randn(B, K1, device=dev)).requires_grad_(True)
        P = [torch.randn(F1, K1, device=dev) / K1 ** 0.5, torch.zeros(F1, device=dev), torch.randn(128, F1, device=dev) / 11,
             torch.zeros(128, device=dev), None, None, torch.randn(64, 128, device=dev) / 11, torch.zeros(64, device=dev),
             torch.randn(12, 64, device=dev) / 8, torch.zeros(12, device=dev)]
        P = [p.requires_grad_(True) if p is not None else None for p in P]
        tm = ops.KernelTimer()
        for it in range(4):
            flush.zero_()
            if it == 1:
                ops.set_timer(tm)
            out = ops.HeadFn.apply(feats, None, None, *P)
            out.backward(torch.ones_like(out))
        torch.cuda.synchronize()
        ops.set_timer(None)
        for k, v in tm.summary().items():
            ms = v["ms"] / v["calls"]
            print(f"{k:<34}{ms:9.3f}{v['flops'] / v['calls'] / ms / 1e9:10.1f}{v['bytes'] / v['calls'] / ms / 1e6:9.0f}")
        if args.only == "head":
            return
    for l, (Ci, Ti, Si) in enumerate(layers):
        Co = 32
        npos = B * (Ti - 2) * (Si - 2) ** 2
        flops = 2.0 * 27 * Ci * Co * npos
        w = torch.randn(Co, Ci, 3, 3, 3, device=dev) / (Ci * 27) ** 0.5
        b = torch.randn(Co, device=dev)
        x = torch.randn(B, Ci, Ti, Si, Si, device=dev)
        gz = torch.randn(B, Co, Ti - 2, Si - 2, Si - 2, device=dev)
        if args.only in ("", "f32"):
            ms = timeit(lambda: ops.conv3d_fwd(x, w, b), flush=flush)
            print(f"conv{l} fwd f32 Ci={Ci:<3}{'':<14}{ms:9.3f}{flops / ms / 1e9:10.1f}{4 * (x.numel() + gz.numel()) / ms / 1e6:9.0f}")
            if l > 0:
                ms = timeit(lambda: ops.conv3d_dgrad(gz, w, x, x.shape), flush=flush)
                print(f"conv{l} dgrad f32{'':<19}{ms:9.3f}{flops / ms / 1e9:10.1f}")
            ms = timeit(lambda: ops.conv3d_wgrad(x, gz), flush=flush)
            print(f"conv{l} wgrad f32{'':<19}{ms:9.3f}{flops / ms / 1e9:10.1f}")
        if args.only in ("", "bf16"):
            xb = ops.to_blocked_bf16(x)
            ms = timeit(lambda: ops.conv3d_fwd_bf16(xb, w, b), flush=flush)
            nbytes = 2.0 * (xb.numel() + ops.blocked_groups(Co) * 8 * npos)
            print(f"conv{l} fwd bf16 Ci={Ci:<3}{'':<13}{ms:9.3f}{flops / ms / 1e9:10.1f}{nbytes / ms / 1e6:9.0f}")
            if l > 0:
                gzp = ops.to_blocked_bf16(gz, pad=2)
                ms = timeit(lambda: ops.conv3d_dgrad_bf16(gzp, w, xb), flush=flush)
                print(f"conv{l} dgrad bf16{'':<18}{ms:9.3f}{flops / ms / 1e9:10.1f}")
            gzw = ops.to_gzw_bf16(gz)
            ms = timeit(lambda: ops.conv3d_wgrad_bf16(xb, gzw, Ci, Co), flush=flush)
            print(f"conv{l} wgrad bf16{'':<18}{ms:9.3f}{flops / ms / 1e9:10.1f}")
            ms = timeit(lambda: ops.to_blocked_bf16(x), flush=flush)
            print(f"conv{l} to_blocked{'':<18}{ms:9.3f}{'':>10}{(4 * x.numel() + 2 * xb.numel()) / ms / 1e6:9.0f}")


if __name__ == "__main__":
    main()
