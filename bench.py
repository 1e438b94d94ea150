#!/usr/bin/env python
"""bench.py -- train samples/sec of the Conv3d PV-yield step on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c1|c3|c4|c5] ...
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full train step of the hot path on one synthetic batch: int16 satellite normalise ->
Conv3d+ReLU stack -> FC head -> L1 loss -> backward -> Adam.  Configurations (BASELINE.json `configs`):

  c1  configs[1]  Conv3d sat-only, fp32, batch 32 PER GPU (weak scaling), 12x19x64x64 int16 cubes, history 30 / forecast
                  60 min, 4 conv layers x 32 channels (141.4 M parameters).  THE DEFAULT: `value`, `e2e`, `roofline`,
                  `cpu_baseline` of the JSON line are this configuration's.
  c3  configs[2]  + NWP (10 x 19 x 2 x 2) + PV history, bf16 tensor-core mode; measured twice: "weak" (128 samples per GPU)
                  and "strong" (global batch 256 split over the ranks; one GPU runs it as two accumulated micro-batches
                  of 128, the largest batch the tensor-core head trains).  The default run appends it as the
                  `c3_bf16` block of the same JSON line (`--no-c3` skips it).
  c4  configs[3]  inference sweep, batch 512-8192 sharded over the ranks without any collective
  c5  configs[4]  deep variant: 8 conv layers, 128x128 crops, bf16, batch 16 per GPU

One JSON line on rank 0:
  value      whole-job samples/s, inputs already resident in HBM, CUDA-event timed, max over ranks
  e2e        the same step through the public API (Model.training_step / backward / optimizer.step) with
             HOST (pinned) input buffers: H2D copy of every step's inputs and D2H read of the loss inside
             the timed region
  roofline   dominant kernel class, measured live with CUDA events inside the timed region
  parity_check  step-0 loss and forecast of the BENCHMARKED batch against the CPU oracle (1e-5 fp32, 2e-2 bf16)
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
SEED = 518  # configs/experiment/conv3d.yaml:16

_SAT = dict(forecast_minutes=60, history_minutes=30, number_of_conv3d_layers=4, conv3d_channels=32, image_size_pixels=64,
            number_sat_channels=12)
CONFIGS = {
    "c1": dict(label="BASELINE configs[1]: Conv3d sat-only train step (fwd+bwd+Adam), int16 sat 12x19x64x64",
               model=dict(include_pv_yield=False, include_nwp=False, **_SAT), precision="fp32", batch=32, mode="train"),
    "c3": dict(label="BASELINE configs[2]: Conv3d + NWP (10x19x2x2) + PV-history train step (fwd+bwd+Adam), int16 sat 12x19x64x64",
               model=dict(include_pv_yield=True, include_nwp=True, **_SAT), precision="bf16", batch=128, mode="train"),
    "c4": dict(label="BASELINE configs[3]: Conv3d inference sweep (no_grad forward), int16 sat 12x19x64x64",
               model=dict(include_pv_yield=False, include_nwp=False, **_SAT), precision="bf16", batch=512, mode="infer"),
    "c5": dict(label="BASELINE configs[4]: deep Conv3d variant (8 layers, 32 channels, 128x128 crops) train step",
               model=dict(include_pv_yield=False, include_nwp=False, **{**_SAT, "number_of_conv3d_layers": 8, "image_size_pixels": 128}),
               precision="bf16", batch=16, mode="train"),
}
MODEL_KW = CONFIGS["c1"]["model"]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c1", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="samples per GPU per step (0 = the configuration's default)")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="strong scaling: global batch split over the ranks (one GPU accumulates micro-batches of <= 128 in bf16)")
    ap.add_argument("--cpu-steps", type=int, default=3, help="timed steps of the CPU baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-c3", action="store_true", help="default run: skip the appended configs[2] (bf16, NWP + PV) block")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-overlap-adam", action="store_true",
                    help="keep the Adam step of fc1.weight on the compute stream (default: side stream, under the next forward)")
    ap.add_argument("--precision", default="", choices=["", "fp32", "bf16"],
                    help="override the configuration's arithmetic: fp32 = fp32-accurate kernels (3xTF32 tensor cores / FMA "
                         "pipe, parity 1e-5); bf16 = bf16 tensor-core mode (parity 2e-2)")
    ap.add_argument("--fp32-fma", action="store_true", help="fp32: keep every convolution on the direct FMA-pipe kernels")
    ap.add_argument("--reserve-sms", type=int, default=-1,
                    help="N > 1: SMs left to NCCL's kernels by the persistent kernels (default 0: measured at N = 8 bf16, "
                         "reserving 16 / 32 SMs speeds the overlapped kernels up but lengthens the step: 3.53 / 3.73 / 4.21 ms)")
    ap.add_argument("--static-tiles", action="store_true",
                    help="N > 1: keep the static work split of the fp32 weight gradient (bit-reproducible sums) instead of chunks "
                         "claimed from an atomic counter (no grid tail behind CTAs displaced by NCCL)")
    ap.add_argument("--no-shard", action="store_true",
                    help="N > 1: replicate the fc1 optimiser (all-reduce) instead of sharding it by output feature")
    ap.add_argument("--infer-batches", default="512,1024,2048,4096,8192", help="c4: global batch sizes of the sweep")
    return ap.parse_args()


def n_params(model_kw):
    L, C, H, ch = model_kw["number_of_conv3d_layers"], model_kw["conv3d_channels"], model_kw["image_size_pixels"], model_kw["number_sat_channels"]
    T = (model_kw["forecast_minutes"] + model_kw["history_minutes"]) // 5 + 1
    feat = C * (H - 2 * L) ** 2 * (T - 2 * L)
    n = (ch * 27 * C + C) + (L - 1) * (C * 27 * C + C) + (feat * 128 + 128) + (128 * 128 + 128)
    fc3_in = 128 + (256 if model_kw["include_pv_yield"] else 0) + (128 if model_kw["include_nwp"] else 0)
    if model_kw["include_nwp"]:
        n += 760 * 128 + 128
    return n + fc3_in * 64 + 64 + 64 * 12 + 12


def workload_config(cfg_name, precision, batch, world, sharded, global_batch=0, micro=1, dynamic=True):
    cfg = CONFIGS[cfg_name]
    arith = {"fp32": "fp32 storage and accumulation; convolutions and fc1 on the tensor cores through split-precision products at fp32 "
                     "accuracy (two-way fp16 split of operands scaled by their tensors' largest magnitudes, 22 significand bits, three "
                     "MMAs per product; 3xTF32 / three-way bf16 split for the shapes it does not take); forecast parity <= 1e-5",
             "bf16": "bf16 tensor-core convolutions and fc1 (fp32 accumulate, fp32 master weights)"}[precision]
    return {
        "workload": f"{cfg['label']}, {arith}",
        "config": cfg_name,
        "precision": precision,
        "batch_per_gpu": batch,
        "global_batch": global_batch or batch * world,
        "micro_batches_per_step": micro,
        "conv3d_layers": cfg["model"]["number_of_conv3d_layers"],
        "conv3d_channels": cfg["model"]["conv3d_channels"],
        "params": n_params(cfg["model"]),
        "parallelism": f"dp{world}" + ("+fc1-optimizer-sharded" if sharded else "") + ("+dynamic-wgrad-chunks" if world > 1 and dynamic else ""),
        "optimizer": "FusedAdam (one launch for the small tensors; fc1.weight: row-sharded under data parallelism)",
        "l2_policy": "working set per step (>= 0.4 GB activations + 0.28-0.57 GB fc1 weights + 4 rotating input "
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
def host_threads():
    threads = os.cpu_count() or 1
    try:
        threads = len(os.sched_getaffinity(0))
    except Exception:
        pass
    return threads


def cpu_step_time(cfg_name: str, batch_size: int, steps: int, warmup: int):
    """Time the reference step as restated by the oracle (torch CPU operators, same call order as
    predict_pv_yield/models/base_model.py:78-153,255-257) on all host threads.  Returns (s_per_step, threads)."""
    import torch

    from oracle import conv3d_oracle as O

    cfg = CONFIGS[cfg_name]
    threads = host_threads()
    torch.set_num_threads(threads)
    torch.manual_seed(SEED)
    m = O.OracleModel(**cfg["model"])
    m.batch_size = batch_size
    opt = m.configure_optimizers()
    batch = O.make_synthetic_batch(batch_size, seed=SEED, image_size_pixels=cfg["model"]["image_size_pixels"])
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        if cfg["mode"] == "infer":
            with torch.no_grad():
                m(batch)
        else:
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

    cfg = CONFIGS[args.config]
    B = args.batch or (32 if args.config in ("c1", "c3", "c4") else cfg["batch"])  # bounded sample: batch 32 steps
    s_per_step, threads = cpu_step_time(args.config, B, args.steps, max(args.warmup, 1))
    value = B / s_per_step
    what = "no_grad forwards" if cfg["mode"] == "infer" else "full train steps"
    line = {
        "impl": "reference", "metric": METRIC if cfg["mode"] == "train" else "inference_samples_per_sec", "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(args.warmup, 1), "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.config, "fp32", B, 1, False),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.steps} {what} of batch {B} (oracle port of the reference "
                                   f"step, torch {torch.__version__} CPU, {threads} threads)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    line["config"]["workload"] = cfg["label"] + ", fp32 (torch CPU: oneDNN / MKL)"
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


HBM_BOUND = {"adam_step_f32", "head_fwd_f32", "head_bwd_f32", "sat_normalise", "sat_normalise_blocked_bf16",
             "sat_normalise_blocked_f32", "nc_to_blocked_f32", "blocked_f32_to_nc",
             "nc_to_blocked_bf16", "blocked_to_nc_f32", "nc_to_gzw_bf16", "adam_fc1_shadow", "adam_fc1_shadow_rows",
             "fc1_fwd_bf16", "fc1_dgrad_bf16", "fc1_wgrad_bf16", "fc1_make_shadow_bf16"}


def roofline_from_timer(timer, steps, ms_total, peaks, fma_peak):
    """Dominant kernel class of the timed region against the roofline that bounds it:
    HBM for the streaming kernels, the bf16 tensor peak for the bf16 convolutions, the FP32 FMA pipe for the fp32 direct
    kernels, and for the split-precision tensor-core convolutions of the fp32 mode the bf16 peak / 3 (two-way fp16 split:
    every fp32-accurate product is three kind::f16 MMAs) or / 6 (3xTF32: three kind::tf32 MMAs at half the bf16 rate;
    three-way bf16 split: six kind::f16 MMAs)."""
    summ = timer.summary()
    classes = {}
    for name, d in summ.items():
        cls = name.split("[")[0]
        c = classes.setdefault(cls, dict(calls=0, ms=0.0, flops=0.0, bytes=0.0))
        for k in ("calls", "ms", "flops", "bytes"):
            c[k] += d[k]
    tf32x3_peak = peaks["bf16_tflops_sustained"] / 6.0
    f16x2_peak = peaks["bf16_tflops_sustained"] / 3.0

    def bound_of(cls):
        if cls in HBM_BOUND:
            return "hbm", peaks["hbm_gbs"]
        if cls.endswith("_f16x2"):
            return "tensor", f16x2_peak
        if cls.endswith("_tf32x3") or cls.endswith("_bf16x3"):
            return "tensor", tf32x3_peak
        if cls.endswith("_bf16"):
            return "tensor", peaks["bf16_tflops_sustained"]
        return "fp32_fma", fma_peak

    per_kernel = {}
    for name, d in sorted(summ.items(), key=lambda kv: -kv[1]["ms"]):
        sec = d["ms"] * 1e-3
        bound, peak = bound_of(name.split("[")[0])
        tf = d["flops"] / sec / 1e12 if sec > 0 else None
        gb = d["bytes"] / sec / 1e9 if sec > 0 else None
        ach = gb if bound == "hbm" else tf
        per_kernel[name] = {
            "calls_per_step": d["calls"] / steps, "ms_per_call": d["ms"] / d["calls"],
            "share_of_step": d["ms"] / ms_total, "tflops": tf, "gbs": gb, "bound": bound,
            "frac": (ach / peak) if (ach and peak) else None,
        }
    dname, dd = max(classes.items(), key=lambda kv: kv[1]["ms"])
    bound, peak = bound_of(dname)
    if bound == "hbm":
        ach = dd["bytes"] / (dd["ms"] * 1e-3) / 1e9
        roof = {"kernel": dname, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "peak_source": peaks["source"] + " (MEASURED_PEAKS.json hbm_gbs)"}
    elif bound == "tensor":
        ach = dd["flops"] / (dd["ms"] * 1e-3) / 1e12
        src = (" (MEASURED_PEAKS.json bf16_tflops_sustained / 6: fp32-accurate products cost three kind::tf32 MMAs at half the "
               "bf16 rate (3xTF32) or six kind::f16 MMAs (three-way bf16 split)") if dname.endswith("x3") else \
            (" (MEASURED_PEAKS.json bf16_tflops_sustained / 3: an fp32-accurate product costs three kind::f16 MMAs in the two-way "
             "fp16 split)") if dname.endswith("_f16x2") else \
            " (MEASURED_PEAKS.json bf16_tflops_sustained: kernel timed inside a long step)"
        roof = {"kernel": dname, "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "peak_source": peaks["source"] + src}
    else:
        ach = dd["flops"] / (dd["ms"] * 1e-3) / 1e12
        roof = {"kernel": dname, "bound": "fp32_fma", "achieved": ach, "peak": fma_peak, "unit": "TFLOP/s",
                "frac": ach / fma_peak if fma_peak else None,
                "peak_source": "FP32 FMA pipe measured live by pvb200_probe_fp32_fma (not in MEASURED_PEAKS.json; theoretical "
                               "148 SM x 128 FMA x 2 x 1.965 GHz = 74.5 TFLOP/s)"}
    # DRAM traffic of the dominant kernel from a committed ncu --set full capture, when one exists for this kernel: the file
    # names the commit it was taken at and the sha1 of the kernel's source then -- compared with the source now
    roof["traffic"] = None
    for fn in ("traffic_r02.json", "traffic_r01c.json"):
        try:
            table = json.load(open(os.path.join(ROOT, "profiles", fn)))
            tr = table.get(dname)
        except Exception:
            table, tr = {}, None
        if tr:
            roof["traffic"] = tr["traffic"]
            note = {"algorithmic_bytes_same_launch": tr["algorithmic"], "launch": tr["launch"],
                    "source": f"profiles/{fn} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum; "
                              "a committed capture, not re-measured by this run)",
                    "capture_commit": table.get("_commit")}
            src = tr.get("source_file")
            if src:
                try:
                    import hashlib
                    now = hashlib.sha1(open(os.path.join(ROOT, src), "rb").read()).hexdigest()
                    note["kernel_source_changed_since_capture"] = now != tr.get("source_sha1")
                except Exception:
                    note["kernel_source_changed_since_capture"] = None
            roof["traffic_note"] = note
            break
    roof["share_of_step"] = dd["ms"] / ms_total
    roof["ms_per_step"] = dd["ms"] / steps
    roof["by_kernel"] = per_kernel
    return roof


class Job:
    """One configuration on this rank: model, optimiser, gradient exchange, synthetic batches."""

    NBUF = 4

    def __init__(self, torch, cfg_name, precision, batch, micro, world, rank, dev, args):
        from oracle import conv3d_oracle as O  # synthetic-input generator (outside the timed regions) and the parity checker
        from predict_pv_yield_b200 import lib
        from predict_pv_yield_b200.models.conv3d.model import Model

        self.torch, self.O = torch, O
        self.cfg_name, self.cfg = cfg_name, CONFIGS[cfg_name]
        self.precision, self.B, self.micro, self.world, self.rank, self.dev = precision, batch, micro, world, rank, dev
        torch.manual_seed(SEED)
        self.model = Model(**self.cfg["model"], precision=precision).to(dev)
        if args.fp32_fma:
            self.model.fp32_tensor_cores = False
        self.model.batch_size = batch
        # fp32 mode, one GPU: the Adam step of fc1.weight (HBM bound) runs on a side stream under the next step's tensor-core
        # bound convolution forward; the work stays inside the timed region (its end synchronises every stream)
        self.model.overlap_optimizer = not args.no_overlap_adam
        self.opt = self.model.configure_optimizers()
        self.exchange = None
        self.sharded = False
        if world > 1 and self.cfg["mode"] == "train":
            from predict_pv_yield_b200.dp import GradientExchange

            self.sharded = not args.no_shard  # fc1 optimiser sharded by rows (bf16: + shadow all-gather; fp32: row all-gather)
            self.exchange = GradientExchange(self.model, shard_large=self.sharded, dynamic_tiles=not args.static_tiles)
            lib.load().pvb200_reserve_sms(max(args.reserve_sms, 0))
            self.exchange.attach_optimizer(self.opt)
        legacy = self.cfg["model"]["include_pv_yield"] or self.cfg["model"]["include_nwp"]
        self.host, self.resident = [], []
        for i in range(self.NBUF * micro):
            b = O.make_synthetic_batch(batch, seed=SEED + 1000 * rank + i, include_legacy_keys=legacy,
                                       image_size_pixels=self.cfg["model"]["image_size_pixels"])
            hb = self._pin(b)
            self.host.append(hb)
            self.resident.append(O.batch_to(hb, dev))
        self.h2d_bytes = micro * sum(t.numel() * t.element_size() for t in self._tensors(self.host[0]))

    def _tensors(self, d):
        for v in d.values():
            if isinstance(v, dict):
                yield from self._tensors(v)
            else:
                yield v

    def _pin(self, d):
        return {k: (self._pin(v) if isinstance(v, dict) else v.pin_memory()) for k, v in d.items()}

    def step(self, batches, i):
        """One optimiser step over `micro` micro-batches (gradient accumulation when micro > 1)."""
        self.opt.zero_grad()
        loss = None
        for mb in batches:
            l = self.model.training_step(mb, i)
            (l / self.micro if self.micro > 1 else l).backward()
            loss = l if loss is None else loss + l
        self.opt.step()
        return loss / self.micro if self.micro > 1 else loss

    def resident_batches(self, i):
        k = (i % self.NBUF) * self.micro
        return self.resident[k: k + self.micro]

    def parity_check(self):
        """Step-0 loss and forecast of the benchmarked batch (rank 0's first resident batch, initial weights) against the
        CPU oracle on the same bits.  Tolerances of BASELINE.json: 1e-5 (fp32), 2e-2 (bf16), normalised max error.
        EVERY rank runs the device half (training_step averages its logged scalars over the ranks: a collective), rank 0
        alone runs the oracle and returns the record."""
        torch, O = self.torch, self.O
        tol = 1e-5 if self.precision == "fp32" else 2e-2
        t0 = time.perf_counter()
        with torch.no_grad():
            y = self.model(self.resident[0]).float().cpu()
            loss = float(self.model.training_step(self.resident[0], 0).detach()) if self.cfg["mode"] == "train" else None
        if self.rank != 0:
            return None
        om = O.OracleModel(**self.cfg["model"])
        om.batch_size = self.B
        om.load_state_dict({k: v.detach().cpu() for k, v in self.model.state_dict().items()})
        with torch.no_grad():
            torch.set_num_threads(host_threads())
            want = om(self.host[0])
            want_loss = float(om.training_step(self.host[0], 0)) if self.cfg["mode"] == "train" else None
        err_y = O.normalised_max_err(y, want)
        out = {"forecast_err": err_y, "tol": tol, "batch": self.B, "oracle_s": time.perf_counter() - t0,
               "what": "normalised max error of the forecast (and relative error of the L1 loss) of the benchmarked batch at "
                       "the initial weights, CUDA path vs oracle/conv3d_oracle.py (torch CPU fp32)"}
        ok = err_y <= tol
        if loss is not None:
            out["loss"] = loss
            out["loss_oracle"] = want_loss
            out["loss_rel_err"] = abs(loss - want_loss) / abs(want_loss)
            ok = ok and out["loss_rel_err"] <= tol
        out["ok"] = bool(ok)
        return out

    def close(self):
        if self.exchange is not None:
            self.exchange.remove()
        self.model = self.opt = self.exchange = None
        self.host = self.resident = None
        from predict_pv_yield_b200 import ops

        ops._persistent.clear()
        self.torch.cuda.empty_cache()


def measure_train(torch, dist, job, args, steps, warmup, with_e2e, with_roofline, sampler=None, fma_peak=0.0):
    """Timed regions of one training configuration.  Returns a dict (rank 0: complete; other ranks: timing only)."""
    from predict_pv_yield_b200 import lib, ops

    world, rank, dev, B = job.world, job.rank, job.dev, job.B

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        job.step(job.resident_batches(i), i)
    barrier()

    # ---- timed region 1: device-resident inputs ---------------------------------------------------------
    timer = ops.KernelTimer()
    barrier()
    if sampler is not None:
        sampler.start()
    if with_roofline:
        ops.set_timer(timer)
    lib.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    trace = os.environ.get("PVB_BENCH_TRACE")  # diagnostics: per-step device and host times of the timed region on stderr
    marks, t_host = [], []
    for i in range(steps):
        job.step(job.resident_batches(i), i)
        if trace:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append(ev)
            t_host.append(time.perf_counter())
    e1.record()
    barrier()
    if trace:
        dev_ms = [round(e0.elapsed_time(m), 2) for m in marks]
        print(f"[trace rank {rank} {job.cfg_name} B={B}x{job.micro}] device ms at step ends: {dev_ms}; host enqueue ms: "
              f"{[round((t - t_host[0]) * 1e3, 1) for t in t_host]}", file=sys.stderr, flush=True)
    launches = lib.launch_count()
    ops.set_timer(None)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)
    # the clock record belongs to the timed region above; nvidia-smi polling is stopped before the end-to-end region,
    # whose per-step host synchronisation makes it sensitive to driver-lock hiccups (observed: 22.9 vs 30.2 ms per step)
    clocks = sampler.stop() if sampler is not None else None
    samples_per_step = B * job.micro * world
    res = {"value": samples_per_step * steps / (ms_total * 1e-3), "ms_per_step": ms_total / steps, "clocks": clocks,
           "gpu_launches": int(launches), "gpu_launches_per_step": launches / steps}

    # ---- timed region 2: end to end from pinned host buffers ---------------------------------------------
    if with_e2e:
        from predict_pv_yield_b200.data import DevicePrefetcher

        def host_batches(n):
            for i in range(n * job.micro):
                yield job.host[i % len(job.host)]

        # allocated ONCE, outside the timed region: the prefetcher's device staging buffers and copy stream, the pinned
        # words the losses are read back into (a cudaHostAlloc / cudaMalloc inside the region synchronises the device)
        prefetcher = DevicePrefetcher(None, dev, depth=job.micro + 2)
        loss_pinned = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]

        def run_steps(n):
            losses_host = []
            loss_ready = [None, None]
            group = []
            i = 0
            prefetcher.batches = host_batches(n)
            for batch in prefetcher:
                group.append(batch)
                if len(group) < job.micro:
                    continue
                loss = job.step(group, i)
                group = []
                loss_pinned[i & 1].copy_(loss.detach(), non_blocking=True)  # D2H read of the step's result
                ev = torch.cuda.Event()
                ev.record()
                loss_ready[i & 1] = ev
                if i > 0:
                    loss_ready[(i - 1) & 1].synchronize()
                    losses_host.append(float(loss_pinned[(i - 1) & 1]))
                i += 1
            loss_ready[(n - 1) & 1].synchronize()
            losses_host.append(float(loss_pinned[(n - 1) & 1]))
            return losses_host

        run_steps(2)  # warm the copy path
        barrier()
        # context for the end-to-end number: raw pinned host -> device bandwidth of this box (boxes of this pool were seen
        # between 2 and 55 GB/s; the per-step input copy is 60 MB at batch 32)
        sat_h, sat_d = job.host[0]["satellite"]["data"], job.resident[0]["satellite"]["data"]
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(4):
            sat_d.copy_(sat_h, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        h2d_gbs = 4 * sat_h.numel() * 2 / (c0.elapsed_time(c1) * 1e-3) / 1e9
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
            losses_host = run_steps(steps)
            if rep == 0:
                loss_host = losses_host[-1]
            assert len(losses_host) == steps
            f1.record()
            barrier()
            wall_ms = (time.perf_counter() - t0) * 1e3
            ms_rep = torch.tensor([max(f0.elapsed_time(f1), wall_ms)], device=dev)
            if world > 1:
                dist.all_reduce(ms_rep, op=dist.ReduceOp.MAX)
            reps.append(float(ms_rep))
        ms2 = sorted(reps)[len(reps) // 2]
        res["e2e"] = {"value": samples_per_step * steps / (float(ms2) * 1e-3), "unit": UNIT, "h2d_bytes_per_step": job.h2d_bytes,
                      "d2h_bytes_per_step": 4, "ms_per_step": float(ms2) / steps, "last_loss": loss_host,
                      "h2d_gbs_measured": h2d_gbs, "repeats": E2E_REPEATS, "reported": "median repetition",
                      "ms_per_step_all": [r / steps for r in reps]}
    if rank == 0 and with_roofline:
        res["roofline"] = roofline_from_timer(timer, steps, ms_total, load_peaks(), fma_peak)
    return res


def measure_infer(torch, dist, job, args, steps, warmup):
    """configs[3]: no_grad forward of a GLOBAL batch sharded over the ranks (no collective on the data path; the forecasts
    stay on the rank that computed them, as a gather of [B, 12] floats is not part of the step).  Returns rank-0 dict."""
    world, rank, dev = job.world, job.rank, job.dev
    out = {}
    sat0 = job.resident[0]["satellite"]["data"]
    for gb in [int(x) for x in args.infer_batches.split(",")]:
        per = gb // world
        reps = (per + sat0.shape[0] - 1) // sat0.shape[0]
        sat = sat0.repeat((reps, 1, 1, 1, 1))[:per].contiguous()
        batch = {"satellite": {"data": sat}}
        with torch.no_grad():
            for _ in range(max(1, warmup // 2)):
                job.model(batch)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = max(2, steps // 4)
            e0.record()
            for _ in range(n):
                y = job.model(batch)
            e1.record()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        out[str(gb)] = {"samples_per_s": per * world * n / (float(ms) * 1e-3), "ms_per_batch": float(ms) / n, "per_gpu_batch": per,
                        "finite": bool(torch.isfinite(y).all())}
        del sat, batch, y
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    from predict_pv_yield_b200 import lib

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
    cfg = CONFIGS[args.config]
    precision = args.precision or cfg["precision"]

    def split_global(gb):
        """global batch -> (per-GPU micro-batch, micro-batches per step): bf16 trains <= 128 per pass through the head"""
        per = gb // world
        cap = 128 if precision == "bf16" else per
        micro = (per + cap - 1) // cap
        return per // micro, micro

    if args.global_batch:
        B, micro = split_global(args.global_batch)
    else:
        B, micro = (args.batch or cfg["batch"]), 1

    uuid = str(torch.cuda.get_device_properties(dev).uuid)
    uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
    job = Job(torch, args.config, precision, B, micro, world, rank, dev, args)
    parity = job.parity_check() if not args.no_parity else None
    if world > 1:
        dist.barrier()

    if cfg["mode"] == "infer":
        sweep = measure_infer(torch, dist, job, args, args.steps, args.warmup)
        if rank == 0:
            top = sweep[max(sweep, key=lambda k: int(k))]
            line = {"metric": "inference_samples_per_sec", "value": top["samples_per_s"], "unit": UNIT, "n_gpus": world,
                    "steps": args.steps, "warmup": args.warmup, "ms_per_step": top["ms_per_batch"], "higher_is_better": True,
                    "scaling": "strong", "vs_baseline": None, "dtype": "f32" if precision == "fp32" else "bf16", "data": "synthetic",
                    "config": workload_config(args.config, precision, top["per_gpu_batch"], world, False), "sweep": sweep,
                    "parity_check": parity, "gpu_launches": int(lib.launch_count())}
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    fma_peak = measure_fp32_fma_peak(torch, lib, dev) if rank == 0 else 0.0
    sampler = ClockSampler(uuid) if rank == 0 else None
    res = measure_train(torch, dist, job, args, args.steps, args.warmup, not args.no_e2e, True, sampler, fma_peak)
    sharded = job.sharded
    job.close()

    c3 = None
    if args.config == "c1" and not args.no_c3 and not args.precision and not args.global_batch and not args.batch:
        # BASELINE configs[2] in the same run: bf16, NWP + PV history; weak (128 per GPU) and strong (global 256)
        c3 = {}
        c3cfg = CONFIGS["c3"]
        for tag in ("weak", "strong"):
            if tag == "weak":
                b3, m3 = c3cfg["batch"], 1
            else:
                per = 256 // world
                m3 = (per + 127) // 128
                b3 = per // m3
            j3 = Job(torch, "c3", "bf16", b3, m3, world, rank, dev, args)
            p3 = j3.parity_check() if (not args.no_parity and tag == "weak") else None
            if world > 1:
                dist.barrier()
            r3 = measure_train(torch, dist, j3, args, args.steps, args.warmup, tag == "weak" and not args.no_e2e, tag == "weak")
            if rank == 0:
                blk = {"value": r3["value"], "unit": UNIT, "ms_per_step": r3["ms_per_step"], "scaling": tag,
                       "config": workload_config("c3", "bf16", b3, world, j3.sharded, global_batch=b3 * m3 * world, micro=m3, dynamic=not args.static_tiles),
                       "gpu_launches": r3["gpu_launches"]}
                if "e2e" in r3:
                    blk["e2e"] = r3["e2e"]
                if "roofline" in r3:
                    blk["roofline"] = r3["roofline"]
                if p3 is not None:
                    blk["parity_check"] = p3
                c3[tag] = blk
            j3.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    line = {
        "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "strong" if args.global_batch else "weak", "vs_baseline": None,
        "dtype": "f32" if precision == "fp32" else "bf16", "data": "synthetic",
        "config": workload_config(args.config, precision, B, world, sharded, global_batch=B * micro * world, micro=micro,
                                  dynamic=not args.static_tiles),
        "clocks": res["clocks"], "e2e": res.get("e2e"),
        "gpu_launches": res["gpu_launches"], "gpu_launches_per_step": res["gpu_launches_per_step"], "roofline": res.get("roofline"),
        "parity_check": parity,
        "peaks": {**peaks, "fp32_fma_tflops_measured": fma_peak, "tf32x3_tflops": peaks["bf16_tflops_sustained"] / 6.0,
                  "f16x2_tflops": peaks["bf16_tflops_sustained"] / 3.0},
    }
    if c3 is not None:
        line["c3_bf16"] = c3
    if world == 1 and not args.no_cpu_baseline:
        cpuB = 32 if args.config != "c5" else B
        s_per_step, threads = cpu_step_time(args.config, cpuB, args.cpu_steps, 1)
        line["cpu_baseline"] = {
            "value": cpuB / s_per_step, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{args.cpu_steps} full train steps of batch {cpuB} after 1 warm-up (oracle port of the reference "
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
