"""FusedAdam -- ``torch.optim.Optimizer`` surface over the multi-tensor Adam kernel.

Replaces ``torch.optim.Adam(self.parameters(), lr=0.0005)`` returned by the reference's
``configure_optimizers`` (``predict_pv_yield/models/base_model.py:255-257``): same defaults
(betas (0.9, 0.999), eps 1e-8, no weight decay, bias-corrected), same state names
(``step``, ``exp_avg``, ``exp_avg_sq``) so optimizer checkpoints keep their layout.  One kernel launch
updates every parameter; under data parallelism ``grad_scale = 1 / world_size`` folds the gradient
averaging into the same pass.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import ops


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, overlap_large: bool = False):
        if lr < 0.0 or eps < 0.0 or not (0.0 <= betas[0] < 1.0) or not (0.0 <= betas[1] < 1.0):
            raise ValueError("FusedAdam: invalid hyper-parameter")
        # weight_decay / amsgrad / maximize are carried at torch.optim.Adam's defaults (the reference's settings,
        # base_model.py:256) so that a FusedAdam checkpoint resumes under torch.optim.Adam and the reverse; step() refuses
        # any other value (e.g. from a checkpoint trained with weight decay) instead of silently ignoring it
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0.0, amsgrad=False, maximize=False))
        self.grad_scale = 1.0
        self.pre_step_hook: Optional[Callable[[], None]] = None  # e.g. wait for the gradient all-reduce
        # The update of a LARGE parameter (fc1.weight: 141 M elements, 4 GB of HBM traffic, ~0.75 ms) is HBM bound and its
        # first reader in the next step is the head at the END of the forward pass, while the Conv3d forward in front of it
        # is tensor-core bound and leaves the memory system idle: with overlap_large the kernel runs on a side stream,
        # under the next step's normalise + convolutions.  Readers of the parameter call dp.wait_ready(p) (Model.forward
        # does; state_dict() does through a hook); the gradient buffer is kept alive until then.  OPT-IN: code that reads
        # the parameter tensor directly between optimizer.step() and the next forward (EMA averaging, ad-hoc checks) must
        # call dp.wait_ready(p) first, which torch.optim.Adam users do not expect -- hence off by default
        # (Model.overlap_optimizer = True before configure_optimizers() switches it on).
        self.overlap_large = bool(overlap_large)
        self.large_numel = 1 << 22
        self.overlap_width_sms = 36  # the side-stream launch is sized as if the device had this many SMs (x 8 CTAs)
        self._side: Optional[torch.cuda.Stream] = None

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self.pre_step_hook is not None:
            self.pre_step_hook()
        for group in self.param_groups:
            if group.get("weight_decay", 0.0) != 0.0 or group.get("amsgrad", False) or group.get("maximize", False):
                raise RuntimeError("FusedAdam implements torch.optim.Adam with weight_decay=0, amsgrad=False, maximize=False "
                                   "(the reference's optimiser, base_model.py:255-257); this parameter group asks for more")
            ps, gs, ms, vs = [], [], [], []
            step = None
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("FusedAdam does not support sparse gradients")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] = int(st["step"]) + 1
                grad = p.grad.data if p.grad.is_contiguous() else p.grad.data.contiguous()
                shadow = getattr(p, "_pvb_shadow", None)
                shard0 = getattr(p, "_pvb_shard", None)
                if shadow is None and shard0 is not None:
                    # fp32 mode under data parallelism: dp.GradientExchange reduce-scattered the gradient rows this rank
                    # owns; Adam on those rows only (same kernel on the contiguous row range), then the updated rows of
                    # every rank are all-gathered in place on the communication stream
                    lo, hi = shard0.rows(p.shape[0])
                    n = p[0].numel()
                    sl = slice(lo * n, hi * n)
                    b1, b2 = group["betas"]
                    ops.adam_step([p.data.view(-1)[sl]], [grad.view(-1)[sl]], [st["exp_avg"].view(-1)[sl]],
                                  [st["exp_avg_sq"].view(-1)[sl]], group["lr"], b1, b2, group["eps"], st["step"], self.grad_scale)
                    shard0.all_gather_rows(p)
                    p._pvb_gen = getattr(p, "_pvb_gen", 0) + 1
                    continue
                if shadow is not None:
                    # fc1.weight in bf16 mode: Adam + refresh of the tensor-core shadow in one pass over the weight
                    b1, b2 = group["betas"]
                    shard = getattr(p, "_pvb_shard", None)
                    if shard is not None:
                        # data parallel, optimiser sharded by output feature (dp.GradientExchange reduce-scattered the
                        # gradient rows): update 1/world of the rows, exchange the bf16 copies on the comm stream
                        if shadow.adam_step_sharded(p, grad, st["exp_avg"], st["exp_avg_sq"], group["lr"], b1, b2, group["eps"],
                                                    st["step"], self.grad_scale, shard.rank, shard.world, shard.group,
                                                    shard.comm_stream(p.device)):
                            continue
                        raise RuntimeError("FusedAdam: the gradient of a sharded parameter was reduce-scattered but its "
                                           "shadow cannot take a sharded step (geometry unknown)")
                    if shadow.adam_step(p, grad, st["exp_avg"], st["exp_avg_sq"], group["lr"], b1, b2, group["eps"],
                                        st["step"], self.grad_scale):
                        continue
                p._pvb_gen = getattr(p, "_pvb_gen", 0) + 1  # updated behind torch's back: derived copies are stale
                if self.overlap_large and p.is_cuda and p.numel() >= self.large_numel and getattr(p, "_pvb_overlap_ok", False):
                    self._step_on_side_stream(p, grad, st, group)
                    continue
                if step is None:
                    step = st["step"]
                elif step != st["step"]:
                    raise RuntimeError("FusedAdam: parameters of one group must share a step count")
                ps.append(p.data)
                gs.append(grad)
                ms.append(st["exp_avg"])
                vs.append(st["exp_avg_sq"])
            if ps:
                b1, b2 = group["betas"]
                ops.adam_step(ps, gs, ms, vs, group["lr"], b1, b2, group["eps"], step, self.grad_scale)
        return loss

    def _step_on_side_stream(self, p, grad, st, group) -> None:
        from .dp import wait_ready

        wait_ready(p)  # a previous update still in flight (two optimizer steps without a forward in between)
        dev = p.device
        if self._side is None:
            self._side = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        ev = torch.cuda.Event()
        ev.record(main)  # the gradient (and, under data parallelism, its reduction) is complete
        self._side.wait_event(ev)
        b1, b2 = group["betas"]
        # narrow grid (~2 CTAs of 256 threads per SM): at full width (8 per SM) the update's CTAs fill the register files and
        # the persistent convolution kernels of the next forward pass (one 320-thread, 124-register CTA per SM) cannot
        # become resident until it has drained -- no overlap at all, and the normalise kernel launched right behind it
        # crawled from 0.1 to 0.7 ms.  Narrow, it runs at ~half of HBM speed beside them and still ends before the head.
        L = ops._lib.load()
        old = int(L.pvb200_reserve_sms(0))
        full = int(L.pvb200_sm_count())
        L.pvb200_reserve_sms(max(full - self.overlap_width_sms, 0))
        try:
            with torch.cuda.stream(self._side):
                ops.adam_step([p.data], [grad], [st["exp_avg"]], [st["exp_avg_sq"]], group["lr"], b1, b2, group["eps"], st["step"],
                              self.grad_scale)
                ready = torch.cuda.Event()
                ready.record(self._side)
        finally:
            L.pvb200_reserve_sms(old)
        p._pvb_ready = ready
        p._pvb_hold = grad  # keeps the gradient's memory from being recycled before the side stream has read it
