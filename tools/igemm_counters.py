"""Print the igemm kernel's internal cycle counters (MMA warp: total, waiting for input planes, waiting for a free
accumulator, issuing; epilogue warp 2: waiting for accumulators, TMEM load, bias/ReLU/store) for one layer, for the
profiling flags 0 (normal), 1 (no plane copies), 2 (no output stores), 16 (no epilogue work), then the time per call of
the production and the profiling build.  python tools/igemm_counters.py fwd|dgrad [layer]"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from predict_pv_yield_b200 import lib, ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
layer = int(sys.argv[2]) if len(sys.argv) > 2 else 1
L = lib.load()
dev = torch.device("cuda:0")
B = 32
T, S, C = 19, 64, 12
for _ in range(layer):
    C, T, S = 32, T - 2, S - 2
w = torch.randn(32, C, 3, 3, 3, device=dev) / (C * 27) ** 0.5
b = torch.randn(32, device=dev)
x = torch.randn(B, C, T, S, S, device=dev)
gz = torch.randn(B, 32, T - 2, S - 2, S - 2, device=dev)
xb = ops.to_blocked_bf16(x)
gzp = ops.to_blocked_bf16(gz, pad=2)
run = (lambda: ops.conv3d_fwd_bf16(xb, w, b)) if which == "fwd" else (lambda: ops.conv3d_dgrad_bf16(gzp, w, xb))
run()
L.pvb200_debug_set_igemm_counters.argtypes = [ctypes.c_void_p]
names = ["mma_total", "mma_wait_full", "mma_wait_tempty", "mma_issue", "epi_rest(neg)", "epi_wait_tfull", "-", "epi_ld"]
for flags in (0, 1, 2, 16):
    dbg = torch.zeros((148, 8), dtype=torch.int64, device=dev)
    L.pvb200_debug_set_igemm_counters(dbg.data_ptr())
    L.pvb200_debug_set_igemm_flags(flags)
    run()
    torch.cuda.synchronize()
    L.pvb200_debug_set_igemm_counters(None)
    L.pvb200_debug_set_igemm_flags(0)
    d = dbg.double().cpu()
    m = d.mean(0)
    print(which, "layer", layer, "flags", flags, {n: round(float(m[i])) for i, n in enumerate(names)}, "max total", int(d[:, 0].max()))


def timed(n=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    run(); torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


print("production kernel: %.1f us/call (back to back, incl. weight prep)" % timed())
dbg = torch.zeros((148, 8), dtype=torch.int64, device=dev)
L.pvb200_debug_set_igemm_counters(dbg.data_ptr())
print("profiling kernel:  %.1f us/call" % timed())
L.pvb200_debug_set_igemm_counters(None)
