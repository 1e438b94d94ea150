#!/usr/bin/env python
"""bench.py -- train samples/sec of the Conv3d PV-yield step on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] ...
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full train step of the hot path on one synthetic batch: int16 satellite normalise ->
Conv3d+ReLU stack -> FC head -> L1 loss -> backward -> Adam.  Workload at every N (weak scaling):
BASELINE configs[1] -- Conv3d sat-only, fp32, batch 32 PER GPU, 12x19x64x64 int16 cubes, history 30 /
forecast 60 min, 4 conv layers x 32 channels (141.4 M parameters).

One JSON line on rank 0:
  value      whole-job samples/s, inputs already resident in HBM, CUDA-event timed, max over ranks
  e2e        the same step through the public API (Model.training_step / backward / optimizer.step) with
             HOST (pinned) input buffers: H2D copy of every step's inputs and D2H read of the loss inside
             the timed region
  roofline   dominant kernel class, measured live with CUDA events inside the timed region
  cpu_baseline  the oracle port of the reference step (torch CPU, all host threads) timed on this box
``--impl reference`` times that CPU port alone and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "train_samples_per_sec"
UNIT = "samples/s"
MODEL_KW = dict(include_pv_yield=False, include_nwp=False, forecast_minutes=60, history_minutes=30,
                number_of_conv3d_layers=4, conv3d_channels=32, image_size_pixels=64, number_sat_channels=12)
SEED = 518  # configs/experiment/conv3d.yaml:16


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="samples per GPU per step")
    ap.add_argument("--cpu-steps", type=int, default=3, help="timed steps of the CPU baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"],
                    help="fp32 = BASELINE configs[1] (exact fp32 kernels); bf16 = tensor-core convolutions (configs[2] dtype)")
    ap.add_argument("--reserve-sms", type=int, default=-1,
                    help="N > 1: SMs left to NCCL's kernels by the persistent kernels (default 0: measured at N = 8 bf16, "
                         "reserving 16 / 32 SMs speeds the overlapped kernels up but lengthens the step: 3.53 / 3.73 / 4.21 ms)")
    ap.add_argument("--no-shard", action="store_true",
                    help="bf16, N > 1: replicate the fc1 optimiser instead of sharding it by output feature")
    return ap.parse_args()


def workload_config(args, world):
    return {
        "workload": "BASELINE configs[1]: Conv3d sat-only train step (fwd+bwd+Adam), int16 sat 12x19x64x64, "
                    + ("fp32" if args.precision == "fp32" else "bf16 tensor-core convolutions (fp32 accumulate, fp32 master "
                       "weights, fp32 FC head)"),
        "precision": args.precision,
        "batch_per_gpu": args.batch,
        "global_batch": args.batch * world,
        "conv3d_layers": 4,
        "conv3d_channels": 32,
        "params": 141414732,
        "parallelism": f"dp{world}" + ("+fc1-optimizer-sharded" if (world > 1 and args.precision == "bf16" and not args.no_shard) else ""),
        "l2_policy": "working set per step (~0.9 GB activations + 0.57 GB fc1 weights + 4 rotating input "
                     "batches) is far larger than the 126 MB L2; no explicit flush",
    }


# ------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampling (recipe of /opt/skills/guides/B200_PROFILING.md) while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_uuid: str):
        self.rows = []
        self.proc = None
        self.uuid = gpu_uuid

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", self.uuid, f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm_sorted = sorted(sm)
        return {"sm_mhz": sm_sorted[len(sm_sorted) // 2], "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------------
# CPU legs (oracle port of the reference step)
# ------------------------------------------------------------------------------------------------------------
def cpu_step_time(batch_size: int, steps: int, warmup: int):
    """Time the reference train step as restated by the oracle (torch CPU operators, same call order as
    predict_pv_yield/models/base_model.py:78-153,255-257) on all host threads.  Returns (s_per_step, threads)."""
    import torch

    from oracle import conv3d_oracle as O

    threads = os.cpu_count() or 1
    try:
        threads = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(threads)
    torch.manual_seed(SEED)
    m = O.OracleModel(**MODEL_KW)
    m.batch_size = batch_size
    opt = m.configure_optimizers()
    batch = O.make_synthetic_batch(batch_size, seed=SEED, include_legacy_keys=False)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        loss = m.training_step(batch, i)
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return sum(times) / len(times), threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # under torchrun only rank 0 measures the CPU reference
    import torch

    s_per_step, threads = cpu_step_time(args.batch, args.steps, max(args.warmup, 1))
    value = args.batch / s_per_step
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(args.warmup, 1), "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.steps} full train steps of batch {args.batch} (oracle port of the reference "
                                   f"step, torch {torch.__version__} CPU, {threads} threads)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# ours
# ------------------------------------------------------------------------------------------------------------
def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
                    "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "source": "measured"}
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def measure_fp32_fma_peak(torch, lib, dev):
    """TFLOP/s of the FP32 FMA pipe, measured live (best of 5) with the library's probe kernel."""
    import ctypes as C

    L = lib.load()
    sink = torch.zeros(4, device=dev)
    flops = C.c_double(0.0)
    stream = torch.cuda.current_stream().cuda_stream
    best = 0.0
    # scalar FFMA and packed FFMA2 chains: the denominator is whichever form of the instruction is faster on this part
    for probe in (L.pvb200_probe_fp32_fma, L.pvb200_probe_fp32_fma2):
        for _ in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lib.check(probe(sink.data_ptr(), 4096, C.byref(flops), stream), "probe")
            e1.record()
            torch.cuda.synchronize()
            best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def run_ours(args):
    import torch
    import torch.distributed as dist

    from oracle import conv3d_oracle as O  # synthetic-input generator only (not on the measured path)
    from predict_pv_yield_b200 import lib, ops
    from predict_pv_yield_b200.models.conv3d.model import Model

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: predict_pv_yield_b200 has no CPU fallback")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib.load()
    B = args.batch

    torch.manual_seed(SEED)
    model = Model(**MODEL_KW, precision=args.precision).to(dev)
    model.batch_size = B
    opt = model.configure_optimizers()
    exchange = None
    if world > 1:
        from predict_pv_yield_b200.dp import GradientExchange

        exchange = GradientExchange(model, shard_large=(args.precision == "bf16" and not args.no_shard))
        lib.load().pvb200_reserve_sms(max(args.reserve_sms, 0))
        exchange.attach_optimizer(opt)

    # synthetic inputs: 4 rotating batches, pinned host copies + device-resident copies
    NBUF = 4
    host, resident = [], []
    for i in range(NBUF):
        b = O.make_synthetic_batch(B, seed=SEED + 1000 * rank + i, include_legacy_keys=False)
        sat = b["satellite"]["data"].pin_memory()
        yld = b["pv"]["pv_yield"].pin_memory()
        host.append((sat, yld))
        resident.append({"satellite": {"data": sat.to(dev)}, "pv": {"pv_yield": yld.to(dev)}})
    h2d_bytes = host[0][0].numel() * 2 + host[0][1].numel() * 4

    def step(batch, i):
        opt.zero_grad()
        loss = model.training_step(batch, i)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(resident[i % NBUF], i)
    barrier()

    fma_peak = measure_fp32_fma_peak(torch, lib, dev) if rank == 0 else 0.0
    uuid = str(torch.cuda.get_device_properties(dev).uuid)
    uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
    sampler = ClockSampler(uuid)

    # ---- timed region 1: device-resident inputs ---------------------------------------------------------
    timer = ops.KernelTimer()
    barrier()
    if rank == 0:
        sampler.start()
    ops.set_timer(timer)
    lib.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(resident[i % NBUF], i)
    e1.record()
    barrier()
    launches = lib.launch_count()
    ops.set_timer(None)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)
    # the clock record belongs to the timed region above; nvidia-smi polling is stopped before the end-to-end region,
    # whose per-step host synchronisation makes it sensitive to driver-lock hiccups (observed: 22.9 vs 30.2 ms per step)
    clocks = sampler.stop() if rank == 0 else None

    # ---- timed region 2: end to end from pinned host buffers ---------------------------------------------
    e2e = None
    if not args.no_e2e:
        from predict_pv_yield_b200.data import DevicePrefetcher

        def host_batches(n):
            for i in range(n):
                sat, yld = host[i % NBUF]
                yield {"satellite": {"data": sat}, "pv": {"pv_yield": yld}}

        for i, batch in enumerate(DevicePrefetcher(host_batches(2), dev)):  # warm the copy path
            step(batch, i)
        barrier()
        # context for the end-to-end number: raw pinned host -> device bandwidth of this box (the per-step input copy is
        # 60 MB; boxes of this pool were seen between 2 and 20 GB/s, below ~2.8 GB/s the fp32 step becomes copy-bound)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(4):
            resident[0]["satellite"]["data"].copy_(host[0][0], non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        h2d_gbs = 4 * host[0][0].numel() * 2 / (c0.elapsed_time(c1) * 1e-3) / 1e9
        # the public input pipeline: pinned int16 cubes, H2D of batch i+1 on a side stream under the compute of batch i.
        # Every step's loss is read back to the host (4-byte D2H into pinned memory); the host consumes it one step
        # late, the way a training loop logs, so the read does not drain the GPU queue between steps.
        # The region synchronises with the host every step, so a single driver / scheduler hiccup of the box (observed:
        # 25 ms once in 10 steps) moves a K = 10 step measurement by 10 %: it is run E2E_REPEATS times and the MEDIAN
        # repetition is reported; every repetition is listed in "ms_per_step_all".
        E2E_REPEATS = 3
        reps = []
        loss_host = None
        for rep in range(E2E_REPEATS):
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            t0 = time.perf_counter()
            f0.record()
            loss_pinned = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
            loss_ready = [None, None]
            losses_host = []
            for i, batch in enumerate(DevicePrefetcher(host_batches(args.steps), dev)):
                loss = step(batch, i)
                loss_pinned[i & 1].copy_(loss.detach(), non_blocking=True)  # D2H read of the step's result
                ev = torch.cuda.Event()
                ev.record()
                loss_ready[i & 1] = ev
                if i > 0:
                    loss_ready[(i - 1) & 1].synchronize()
                    losses_host.append(float(loss_pinned[(i - 1) & 1]))
            loss_ready[(args.steps - 1) & 1].synchronize()
            losses_host.append(float(loss_pinned[(args.steps - 1) & 1]))
            if rep == 0:
                loss_host = losses_host[-1]  # after the same number of optimiser steps in every run of the bench
            assert len(losses_host) == args.steps
            f1.record()
            barrier()
            wall_ms = (time.perf_counter() - t0) * 1e3
            ms_rep = torch.tensor([max(f0.elapsed_time(f1), wall_ms)], device=dev)
            if world > 1:
                dist.all_reduce(ms_rep, op=dist.ReduceOp.MAX)
            reps.append(float(ms_rep))
        ms2 = sorted(reps)[len(reps) // 2]
        e2e = {"value": B * world * args.steps / (float(ms2) * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
               "d2h_bytes_per_step": 4, "ms_per_step": float(ms2) / args.steps, "last_loss": loss_host,
               "h2d_gbs_measured": h2d_gbs, "repeats": E2E_REPEATS, "reported": "median repetition",
               "ms_per_step_all": [r / args.steps for r in reps]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline from the per-call events of region 1 ----------------------------------------------------
    peaks = load_peaks()
    summ = timer.summary()
    classes = {}
    for name, d in summ.items():
        cls = name.split("[")[0]
        c = classes.setdefault(cls, dict(calls=0, ms=0.0, flops=0.0, bytes=0.0))
        for k in ("calls", "ms", "flops", "bytes"):
            c[k] += d[k]
    per_kernel = {}
    for name, d in sorted(summ.items(), key=lambda kv: -kv[1]["ms"]):
        sec = d["ms"] * 1e-3
        per_kernel[name] = {
            "calls_per_step": d["calls"] / args.steps, "ms_per_call": d["ms"] / d["calls"],
            "share_of_step": d["ms"] / ms_total, "tflops": d["flops"] / sec / 1e12 if sec > 0 else None,
            "gbs": d["bytes"] / sec / 1e9 if sec > 0 else None,
        }
    hbm_bound = {"adam_step_f32", "head_fwd_f32", "head_bwd_f32", "sat_normalise", "sat_normalise_blocked_bf16",
                 "nc_to_blocked_bf16", "blocked_to_nc_f32", "nc_to_gzw_bf16", "adam_fc1_shadow", "adam_fc1_shadow_rows",
                 "fc1_fwd_bf16", "fc1_dgrad_bf16", "fc1_wgrad_bf16", "fc1_make_shadow_bf16"}
    dom = max(classes.items(), key=lambda kv: kv[1]["ms"])
    dname, dd = dom
    if dname in hbm_bound:
        ach = dd["bytes"] / (dd["ms"] * 1e-3) / 1e9
        roof = {"kernel": dname, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"], "peak_source": peaks["source"] + " (MEASURED_PEAKS.json hbm_gbs)"}
    elif dname.endswith("_bf16"):
        ach = dd["flops"] / (dd["ms"] * 1e-3) / 1e12
        roof = {"kernel": dname, "bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": ach / peaks["bf16_tflops_sustained"],
                "peak_source": peaks["source"] + " (MEASURED_PEAKS.json bf16_tflops_sustained: kernel timed inside a long step)"}
    else:
        ach = dd["flops"] / (dd["ms"] * 1e-3) / 1e12
        roof = {"kernel": dname, "bound": "fp32_fma", "achieved": ach, "peak": fma_peak, "unit": "TFLOP/s",
                "frac": ach / fma_peak if fma_peak else None,
                "peak_source": "FP32 FMA pipe measured live by pvb200_probe_fp32_fma (fp32 mode cannot use the bf16 "
                               "tensor peak of MEASURED_PEAKS.json: 1e-5 parity rules out reduced-precision MMA)"}
    # DRAM traffic of the dominant kernel from the committed ncu --set full capture (per launch at the conv1 layer shape;
    # the live figure above is the average over all launches of the class)
    roof["traffic"] = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic_r01c.json"))).get(dname)
        if tr:
            roof["traffic"] = tr["traffic"]
            roof["traffic_note"] = {"algorithmic_bytes_same_launch": tr["algorithmic"], "launch": tr["launch"],
                                    "source": "profiles/traffic_r01c.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"}
    except Exception:
        pass
    roof["share_of_step"] = dd["ms"] / ms_total
    roof["ms_per_step"] = dd["ms"] / args.steps
    roof["by_kernel"] = per_kernel

    line = {
        "metric": METRIC, "value": B * world * args.steps / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
        "config": workload_config(args, world), "clocks": clocks, "e2e": e2e,
        "gpu_launches": int(launches), "gpu_launches_per_step": launches / args.steps, "roofline": roof,
        "peaks": {**peaks, "fp32_fma_tflops_measured": fma_peak},
    }
    if world == 1 and not args.no_cpu_baseline:
        s_per_step, threads = cpu_step_time(B, args.cpu_steps, 1)
        line["cpu_baseline"] = {
            "value": B / s_per_step, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{args.cpu_steps} full train steps of batch {B} after 1 warm-up (oracle port of the reference "
                      f"step, torch {torch.__version__} CPU)", "s_per_step": s_per_step}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
