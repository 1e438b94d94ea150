"""Data-parallel gradient exchange for the Conv3d PV step: one process per GPU, NCCL over NVLink 5.

The reference has no distributed code of its own; multi-GPU training would come from Lightning's DDP
wrapper (``configs/trainer/all_params.yaml:8-11,30,43``; ``sync_dist=True`` at ``base_model.py:117``).
The step shards by batch (samples are independent: no BatchNorm), so the ONLY exchange is the gradient
sum.  ``fc1.weight`` is 99.9 % of the bytes (565 MB fp32) and is the FIRST large gradient the backward
pass produces, so:

* a post-accumulate-grad hook on every parameter fires as soon as autograd has written ``p.grad``;
* "large" gradients (``fc1.weight``) are all-reduced immediately, in place, on a side stream that waits
  on an event of the compute stream -> the transfer overlaps the whole Conv3d backward;
* the remaining ~0.2 M gradient elements are packed into one flat bucket and all-reduced once, in
  ``finish()``;
* averaging (1 / world_size) is folded into the Adam kernel (``FusedAdam.grad_scale``), so no extra pass.

``finish()`` must run before the optimizer step (``attach_optimizer`` wires it as the pre-step hook).
Works on CPU tensors with the ``gloo`` backend too (no streams), which is how the host logic is tested.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist


class GradientExchange:
    def __init__(self, module: torch.nn.Module, process_group=None, large_numel: int = 1 << 22):
        if not dist.is_available() or not dist.is_initialized():
            raise RuntimeError("GradientExchange needs an initialised torch.distributed process group")
        self.group = process_group
        self.world_size = dist.get_world_size(process_group)
        self.large_numel = large_numel
        self.params: List[torch.nn.Parameter] = [p for p in module.parameters() if p.requires_grad]
        self._pending: List = []  # (work handle | None, event | None)
        self._small: List[torch.nn.Parameter] = []
        self._handles = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self._comm_stream: Optional[torch.cuda.Stream] = None
        self.bytes_reduced_last_step = 0
        self._bytes = 0

    # -- hook ---------------------------------------------------------------------------------------
    def _comm(self, device: torch.device) -> torch.cuda.Stream:
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=device)
        return self._comm_stream

    def _all_reduce_async(self, t: torch.Tensor) -> None:
        self._bytes += t.numel() * t.element_size()
        if t.is_cuda:
            comm = self._comm(t.device)
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(t.device))
            comm.wait_event(ready)  # gradient kernel finished
            with torch.cuda.stream(comm):
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
                done = torch.cuda.Event()
                done.record(comm)
            t.record_stream(comm)
            self._pending.append((None, done))
        else:
            work = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._pending.append((work, None))

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        if self.world_size == 1 or p.grad is None:
            return
        if p.grad.numel() >= self.large_numel:
            self._all_reduce_async(p.grad)
        else:
            self._small.append(p)

    # -- end of backward ------------------------------------------------------------------------------
    def finish(self) -> None:
        """Reduce the small-gradient bucket, then make the compute stream wait for every transfer."""
        if self.world_size > 1 and self._small:
            grads = [p.grad for p in self._small]
            flat = torch.cat([g.reshape(-1) for g in grads])
            self._all_reduce_async(flat)
            self._wait_all()
            off = 0
            for g in grads:
                n = g.numel()
                g.copy_(flat[off: off + n].view_as(g))
                off += n
        else:
            self._wait_all()
        self._small = []
        if self._bytes:
            self.bytes_reduced_last_step, self._bytes = self._bytes, 0

    def _wait_all(self) -> None:
        for work, ev in self._pending:
            if work is not None:
                work.wait()
            if ev is not None:
                torch.cuda.current_stream().wait_event(ev)  # device-side wait, no host sync
        self._pending = []

    def attach_optimizer(self, optimizer) -> None:
        """Wire ``finish`` as the optimizer's pre-step hook and fold the 1/world averaging into Adam."""
        if hasattr(optimizer, "pre_step_hook") and hasattr(optimizer, "grad_scale"):
            optimizer.pre_step_hook = self.finish
            optimizer.grad_scale = 1.0 / self.world_size
        else:  # plain torch optimizer: average explicitly
            def _hook(opt, args, kwargs):
                self.finish()
                for p in self.params:
                    if p.grad is not None:
                        p.grad.div_(self.world_size)
            optimizer.register_step_pre_hook(_hook)

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []


def reduce_logged_scalars(values: Dict[str, torch.Tensor], process_group=None) -> Dict[str, torch.Tensor]:
    """The ``sync_dist=True`` of ``base_model.py:108-119``: mean of the logged scalars over ranks, as ONE
    all-reduce of a packed vector instead of one collective per scalar."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(process_group) == 1:
        return values
    keys = sorted(values)
    flat = torch.stack([values[k].detach().reshape(()) for k in keys])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=process_group)
    flat = flat / dist.get_world_size(process_group)
    return {k: flat[i] for i, k in enumerate(keys)}
