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

* ``shard_large=True``, fp32 mode: the optimiser of the large parameter is sharded by rows.  Its gradient is
  REDUCE-SCATTERED (half the NVLink bytes of an all-reduce), rank r runs Adam on its rows only (1/N of the 4 GB optimiser
  pass) and the updated fp32 rows are all-gathered in place on the communication stream, under the next step's
  convolution forward (``ShardSpec.all_gather_rows``; readers call ``wait_ready``).
* ``shard_large=True`` (bf16 mode, where ``fc1.weight`` has a tensor-core shadow): the optimiser of the large
  parameter is sharded by output feature.  Its gradient is REDUCE-SCATTERED (half the NVLink bytes of an
  all-reduce): rank r receives the summed rows [r*F/N, (r+1)*F/N), runs Adam on those rows only (1/N of the 4 GB
  optimiser pass) and the bf16 copies of the updated rows are all-gathered on the communication stream into every
  rank's shadow, under the next step's convolution forward (fc1 is the last consumer).  The fp32 master rows owned by
  other ranks go stale until ``gather_master_weights()`` (wired as the module's state_dict pre-hook).

``finish()`` must run before the optimizer step (``attach_optimizer`` wires it as the pre-step hook).
Works on CPU tensors with the ``gloo`` backend too (no streams), which is how the host logic is tested.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist


class ShardSpec:
    """Attached to a parameter as ``_pvb_shard``: this rank owns rows [rank*F/world, (rank+1)*F/world)."""

    def __init__(self, exchange: "GradientExchange"):
        self.rank = dist.get_rank(exchange.group)
        self.world = exchange.world_size
        self.group = exchange.group
        self._exchange = exchange

    def comm_stream(self, device: torch.device):
        return self._exchange._comm(device)

    def rows(self, nrows_total: int):
        n = nrows_total // self.world
        return self.rank * n, (self.rank + 1) * n

    @torch.no_grad()
    def all_gather_rows(self, p: torch.Tensor) -> None:
        """fp32 mode (no bf16 shadow): every rank has just updated its own rows of ``p``; all-gather them IN PLACE on the
        communication stream.  ``p._pvb_ready`` is the event a reader of ``p`` waits on -- the head of the NEXT forward
        pass, so the transfer hides under the next step's normalise + convolution forward (fc1 is the last consumer)."""
        lo, hi = self.rows(p.shape[0])
        n = p[0].numel()
        flat = p.data.view(-1)
        mine = flat[lo * n: hi * n]
        if p.is_cuda and self._exchange._reduce_scatter_ok:
            comm = self.comm_stream(p.device)
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(p.device))
            comm.wait_event(done)  # the Adam kernel on this rank's rows has finished
            with torch.cuda.stream(comm):
                dist.all_gather_into_tensor(flat, mine, group=self.group)  # in place: mine == flat[rank * count ...]
                ready = torch.cuda.Event()
                ready.record(comm)
            p._pvb_ready = ready
        else:  # gloo (CPU tests)
            parts = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(parts, mine.clone(), group=self.group)
            flat.copy_(torch.cat(parts))
            p._pvb_ready = None


def wait_ready(p: torch.Tensor) -> None:
    """Make the current stream wait for an in-flight update of ``p`` on another stream: the all-gather of its rows
    (``ShardSpec.all_gather_rows``) or its Adam step on the optimiser's side stream (``FusedAdam.overlap_large``)."""
    ev = getattr(p, "_pvb_ready", None)
    if ev is not None:
        torch.cuda.current_stream(p.device).wait_event(ev)
        p._pvb_ready = None
        p._pvb_hold = None  # (FusedAdam.overlap_large) the current stream is now ordered after the update: safe to recycle


class GradientExchange:
    def __init__(self, module: torch.nn.Module, process_group=None, large_numel: int = 1 << 22, shard_large: bool = False,
                 dynamic_tiles: bool = False):
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
        self.shard_large = shard_large
        self._sharded: List[torch.nn.Parameter] = []
        self._reduce_scatter_ok = dist.get_backend(process_group) == "nccl"  # gloo has no reduce-scatter
        if shard_large:
            self._sd_hook = module.register_state_dict_pre_hook(lambda *a, **k: self.gather_master_weights())
        if dynamic_tiles and self.world_size > 1 and any(p.is_cuda for p in self.params):
            # NCCL's kernels take SMs while the persistent convolution kernels run: the fp32-mode weight gradient then claims its
            # work in chunks instead of splitting it statically (no grid tail behind displaced CTAs: 4.71 -> 4.53 ms per step on
            # two GPUs).  OPT-IN: which CTA sums which chunk depends on timing, so the fp32 sums are no longer bit-reproducible
            # from run to run (within the parity bound; `sharded == replicated` bit for bit holds with the static split only).
            from . import lib as _lib
            _lib.load().pvb200_set_dynamic_tiles(1)

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
            # no t.record_stream(comm): finish() makes the compute stream wait for `done` before the optimizer step, and
            # the gradient is only freed after that (zero_grad of the next step), so stream order already protects the
            # buffer.  record_stream deferred the free of the 565 MB gradient past the next allocation: the caching
            # allocator then rotated through extra blocks and hit cudaMalloc (a ~100 ms host stall) inside training loops.
            self._pending.append((None, done))
        else:
            work = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._pending.append((work, None))

    def _can_shard(self, p: torch.nn.Parameter) -> bool:
        shadow = getattr(p, "_pvb_shadow", None)
        # ``trained_through``: the forward of THIS step read the weight through its bf16 shadow (ops.HeadBf16Fn /
        # Fc1Bf16Fn).  A step that went through the fp32 head (training batch > 128 per GPU) reads the fp32 master,
        # whose rows owned by other ranks are stale under a row-sharded optimiser: such a step must all-reduce.
        rows_ok = (self.shard_large and p.dim() == 2 and p.shape[0] % self.world_size == 0 and p.grad.is_contiguous()
                   and p.is_contiguous())
        if shadow is None:
            # fp32 mode: rows sharded, the updated fp32 rows are all-gathered in place under the next forward pass
            return rows_ok
        return rows_ok and getattr(shadow, "geom", None) is not None and getattr(shadow, "trained_through", False)

    def _reduce_scatter_async(self, p: torch.nn.Parameter) -> None:
        """Sum of the gradient rows this rank owns, in place in ``p.grad`` (the other rows keep the local values)."""
        if getattr(p, "_pvb_shard", None) is None:
            p._pvb_shard = ShardSpec(self)
            self._sharded.append(p)
        g = p.grad
        lo, hi = p._pvb_shard.rows(g.shape[0])
        flat = g.view(-1)
        n = g.shape[1]
        mine = flat[lo * n: hi * n]
        self._bytes += mine.numel() * g.element_size()
        if g.is_cuda and self._reduce_scatter_ok:
            comm = self._comm(g.device)
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(g.device))
            comm.wait_event(ready)
            with torch.cuda.stream(comm):
                dist.reduce_scatter_tensor(mine, flat, op=dist.ReduceOp.SUM, group=self.group)  # in place (NCCL)
                done = torch.cuda.Event()
                done.record(comm)
            self._pending.append((None, done))  # no record_stream: see _all_reduce_async
        else:  # backends without reduce-scatter (gloo, CPU tests): all-reduce, the owned rows are what is used
            self._bytes -= mine.numel() * g.element_size()
            self._all_reduce_async(g)

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        if self.world_size == 1 or p.grad is None:
            return
        if p.grad.numel() >= self.large_numel:
            if self._can_shard(p):
                self._reduce_scatter_async(p)
            else:
                if getattr(p, "_pvb_shard", None) is not None:
                    raise RuntimeError("GradientExchange: a parameter whose optimiser is sharded by rows was used by a step "
                                       "that cannot be sharded (forward through the fp32 master weight, e.g. a training "
                                       "batch > 128 per GPU in bf16 mode): its foreign rows are stale.  Call "
                                       "gather_master_weights() and rebuild the exchange with shard_large=False.")
                self._all_reduce_async(p.grad)
        else:
            self._small.append(p)

    @torch.no_grad()
    def gather_master_weights(self) -> None:
        """All-gather the fp32 master rows of the sharded parameters (each rank only keeps its own rows current).
        Called before ``state_dict()``; a collective: every rank must call it."""
        for p in self._sharded:
            if getattr(p, "_pvb_shadow", None) is None:
                wait_ready(p)  # fp32 mode: the rows are re-gathered after every step, only the transfer may be in flight
                continue
            lo, hi = p._pvb_shard.rows(p.shape[0])
            flat = p.data.view(-1)
            n = p.shape[1]
            if self._reduce_scatter_ok:
                dist.all_gather_into_tensor(flat, flat[lo * n: hi * n].clone(), group=self.group)
            else:
                parts = [torch.empty_like(flat[lo * n: hi * n]) for _ in range(self.world_size)]
                dist.all_gather(parts, flat[lo * n: hi * n].clone(), group=self.group)
                flat.copy_(torch.cat(parts))

    @torch.no_grad()
    def gather_optimizer_state(self, optimizer) -> None:
        """All-gather the Adam moments (``exp_avg`` / ``exp_avg_sq``) of the row-sharded parameters: each rank only keeps
        the moments of its own rows current, so a checkpoint written from one rank would otherwise resume with zero /
        stale moments for (world-1)/world of fc1.weight next to a large ``step``.  Wired as the optimizer's
        state_dict pre-hook by ``attach_optimizer``; a collective: every rank must call ``optimizer.state_dict()``."""
        for p in self._sharded:
            st = optimizer.state.get(p)
            if not st:
                continue
            lo, hi = p._pvb_shard.rows(p.shape[0])
            n = p.shape[1]
            for key in ("exp_avg", "exp_avg_sq"):
                t = st.get(key)
                if t is None:
                    continue
                flat = t.view(-1)
                mine = flat[lo * n: hi * n].clone()
                if self._reduce_scatter_ok:
                    dist.all_gather_into_tensor(flat, mine, group=self.group)
                else:
                    parts = [torch.empty_like(mine) for _ in range(self.world_size)]
                    dist.all_gather(parts, mine, group=self.group)
                    flat.copy_(torch.cat(parts))

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
        if self.shard_large and hasattr(optimizer, "register_state_dict_pre_hook"):
            self._opt_sd_hook = optimizer.register_state_dict_pre_hook(lambda opt: self.gather_optimizer_state(opt))
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
        """Detach from the module: gradient hooks, the state_dict pre-hook and the shard marks (gathers the master rows first)."""
        if self._sharded:
            self.gather_master_weights()
            for p in self._sharded:
                p._pvb_shard = None
            self._sharded = []
        for h in self._handles:
            h.remove()
        self._handles = []
        if getattr(self, "_sd_hook", None) is not None:
            self._sd_hook.remove()
            self._sd_hook = None
        if getattr(self, "_opt_sd_hook", None) is not None:
            self._opt_sd_hook.remove()
            self._opt_sd_hook = None


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
