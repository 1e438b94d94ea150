"""CPU, world_size 2, gloo: host logic of the data-parallel gradient exchange (predict_pv_yield_b200/dp.py).

The exchange is tensor-agnostic (it hooks ``p.grad``), so a small torch model on CPU exercises exactly the code
that runs under NCCL: the immediate in-place all-reduce of "large" gradients, the packed bucket of small ones,
``finish()`` ordering, the optimizer wiring and the packed logged-scalar reduction.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from predict_pv_yield_b200.dp import GradientExchange, reduce_logged_scalars

        torch.manual_seed(0)  # identical replicas
        model = torch.nn.Sequential(torch.nn.Linear(64, 128), torch.nn.ReLU(), torch.nn.Linear(128, 3))
        ex = GradientExchange(model, large_numel=4096)  # first weight (8192 el) is "large", the rest go in the bucket
        opt = torch.optim.SGD(model.parameters(), lr=0.1)
        ex.attach_optimizer(opt)
        g = torch.Generator().manual_seed(100 + rank)  # different shard per rank
        x = torch.randn(8, 64, generator=g)
        y = torch.randn(8, 3, generator=g)
        before = [p.detach().clone() for p in model.parameters()]
        loss = ((model(x) - y) ** 2).mean()
        loss.backward()
        local = [p.grad.detach().clone() for p in model.parameters()]
        ex.finish()
        summed = [p.grad.detach().clone() for p in model.parameters()]
        # reference: explicit all-reduce of the local gradients
        want = []
        for t in local:
            t = t.clone()
            dist.all_reduce(t)
            want.append(t)
        ok_sum = all(torch.allclose(a, b, atol=1e-6) for a, b in zip(summed, want))
        # a second finish() is a no-op; the optimizer hook averages and steps
        opt.step()
        after = [p.detach().clone() for p in model.parameters()]
        ok_step = all(torch.allclose(a, b0 - 0.1 * w / world, atol=1e-6) for a, b0, w in zip(after, before, want))
        # replicas stay identical
        flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
        other = flat.clone()
        dist.broadcast(other, src=0)
        ok_same = torch.allclose(flat, other)
        red = reduce_logged_scalars({"MSE/Train": torch.tensor(float(rank + 1)), "NMAE/Train": torch.tensor(2.0 * rank)})
        ok_log = abs(float(red["MSE/Train"]) - 1.5) < 1e-6 and abs(float(red["NMAE/Train"]) - 1.0) < 1e-6
        q.put((rank, ok_sum, ok_step, ok_same, ok_log, ex.bytes_reduced_last_step))
    finally:
        dist.destroy_process_group()


def test_gradient_exchange_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_sum, ok_step, ok_same, ok_log, nbytes in res:
        assert ok_sum, f"rank {rank}: summed gradients differ from an explicit all-reduce"
        assert ok_step, f"rank {rank}: optimizer step did not use the averaged gradient"
        assert ok_same, f"rank {rank}: replicas diverged"
        assert ok_log, f"rank {rank}: logged-scalar reduction wrong"
        assert nbytes == 4 * (64 * 128 + 128 + 128 * 3 + 3)


class _FakeShadow:
    """Test double for ops.Fc1Shadow: the sharded Adam step in plain torch (CPU), same call signature."""

    def __init__(self):
        self.geom = (1, 1, 1, 1)
        self.trained_through = True  # the forward read the weight through the shadow (ops.HeadBf16Fn sets this)
        self.full_bf16 = None

    def adam_step_sharded(self, w1, grad, m, v, lr, b1, b2, eps, step, grad_scale, rank, world, group, comm_stream):
        n = w1.shape[0] // world
        lo, hi = rank * n, (rank + 1) * n
        g = grad[lo:hi] * grad_scale
        m[lo:hi].mul_(b1).add_(g, alpha=1 - b1)
        v[lo:hi].mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v[lo:hi] / (1 - b2 ** step)).sqrt_().add_(eps)
        w1.data[lo:hi].addcdiv_(m[lo:hi] / (1 - b1 ** step), denom, value=-lr)
        parts = [torch.empty_like(w1.data[lo:hi], dtype=torch.bfloat16) for _ in range(world)]
        dist.all_gather(parts, w1.data[lo:hi].bfloat16(), group=group)
        self.full_bf16 = torch.cat(parts)
        return True


def _worker_sharded(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from predict_pv_yield_b200.dp import GradientExchange

        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(64, 128), torch.nn.ReLU(), torch.nn.Linear(128, 3))
        big = model[0].weight
        big._pvb_shadow = _FakeShadow()
        ref = torch.nn.Sequential(torch.nn.Linear(64, 128), torch.nn.ReLU(), torch.nn.Linear(128, 3))
        ref.load_state_dict(model.state_dict())
        ex = GradientExchange(model, large_numel=4096, shard_large=True)
        m, v = torch.zeros_like(big), torch.zeros_like(big)
        ref_opt = torch.optim.Adam([ref[0].weight], lr=1e-2)
        ok = True
        for step in range(1, 4):
            g = torch.Generator().manual_seed(100 * step + rank)
            x, y = torch.randn(8, 64, generator=g), torch.randn(8, 3, generator=g)
            model.zero_grad()
            ((model(x) - y) ** 2).mean().backward()
            ex.finish()
            spec = big._pvb_shard
            big._pvb_shadow.adam_step_sharded(big, big.grad, m, v, 1e-2, 0.9, 0.999, 1e-8, step, 1.0 / world, spec.rank,
                                              spec.world, spec.group, None)
            # reference: every rank computes the full averaged gradient and a plain torch Adam step on the whole weight
            ref.zero_grad()
            ((ref(x) - y) ** 2).mean().backward()
            gr = ref[0].weight.grad.clone()
            dist.all_reduce(gr)
            ref[0].weight.grad.copy_(gr / world)
            ref_opt.step()
            lo, hi = spec.rows(128)
            ok = ok and torch.allclose(big.data[lo:hi], ref[0].weight.data[lo:hi], atol=1e-6)
            ok = ok and torch.allclose(big._pvb_shadow.full_bf16.float(), ref[0].weight.data, atol=1e-2)
            # the model itself must keep following the reference in the forward pass: copy the gathered rows as the
            # real bf16 shadow would provide them (here: fp32 all-gather through the exchange)
            ex.gather_master_weights()
            ok = ok and torch.allclose(big.data, ref[0].weight.data, atol=1e-6)
        sd = model.state_dict()  # pre-hook gathers (collective)
        ok_sd = torch.allclose(sd["0.weight"], ref[0].weight.data, atol=1e-6)
        # optimizer checkpoints: the moments of the rows other ranks own are gathered too (a checkpoint from
        # rank 0 must not resume with zero moments next to a large step count)
        lo, hi = spec.rows(128)
        stale = not torch.allclose(m, ref_opt.state[ref[0].weight]["exp_avg"], atol=1e-7)

        class _Opt:
            state = {big: {"step": 3, "exp_avg": m, "exp_avg_sq": v}}

        ex.gather_optimizer_state(_Opt())
        ok_opt = stale and torch.allclose(m, ref_opt.state[ref[0].weight]["exp_avg"], atol=1e-7) and \
            torch.allclose(v, ref_opt.state[ref[0].weight]["exp_avg_sq"], atol=1e-9)
        # a step that did NOT read the weight through the shadow (fp32 head: batch > 128) must refuse to shard
        big._pvb_shadow.trained_through = False
        model.zero_grad()
        try:
            ((model(x) - y) ** 2).mean().backward()
            ok_guard = False
        except RuntimeError as e:
            ok_guard = "stale" in str(e)
        # the logged scalars of the stand-in LightningModule are averaged over the ranks (sync_dist=True)
        from predict_pv_yield_b200.models.base_model import _Base

        mod = _Base()
        mod.log_dict({"MSE/Train": torch.tensor(float(rank + 1))}, on_step=True, sync_dist=True)
        ok_log = abs(float(mod.logged_metrics["MSE/Train"]) - 1.5) < 1e-6
        q.put((rank, bool(ok), bool(ok_sd and ok_opt and ok_guard and ok_log), spec.rows(128)))
    finally:
        dist.destroy_process_group()


def test_sharded_large_parameter_two_ranks_gloo():
    """fc1-style optimiser sharding: each rank steps its own rows, bf16 copies are all-gathered, master rows are
    gathered on demand (host logic of dp.GradientExchange(shard_large=True) with a torch test double for the kernels)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_sharded, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rows = set()
    for rank, ok, ok_sd, r in res:
        assert ok, f"rank {rank}: sharded step diverged from the replicated reference"
        assert ok_sd, f"rank {rank}: state_dict() / optimizer-state gathering, the shard guard or the logged-scalar mean failed"
        rows.add(r)
    assert rows == {(0, 64), (64, 128)}


def _worker_fp32_rows(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from predict_pv_yield_b200.dp import GradientExchange

        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(64, 128), torch.nn.ReLU(), torch.nn.Linear(128, 3))
        big = model[0].weight  # no shadow: the fp32 row-sharded path
        ref = torch.nn.Sequential(torch.nn.Linear(64, 128), torch.nn.ReLU(), torch.nn.Linear(128, 3))
        ref.load_state_dict(model.state_dict())
        ex = GradientExchange(model, large_numel=4096, shard_large=True)
        ref_opt = torch.optim.SGD([ref[0].weight], lr=0.1)
        ok = True
        for step in range(3):
            g = torch.Generator().manual_seed(100 * step + rank)
            x, y = torch.randn(8, 64, generator=g), torch.randn(8, 3, generator=g)
            model.zero_grad()
            ((model(x) - y) ** 2).mean().backward()
            ex.finish()
            spec = big._pvb_shard
            lo, hi = spec.rows(128)
            with torch.no_grad():  # this rank's rows only (what FusedAdam does with the Adam kernel), then the row all-gather
                big.data[lo:hi] -= 0.1 * big.grad[lo:hi] / world
            spec.all_gather_rows(big)
            ref.zero_grad()
            ((ref(x) - y) ** 2).mean().backward()
            gr = ref[0].weight.grad.clone()
            dist.all_reduce(gr)
            ref[0].weight.grad.copy_(gr / world)
            ref_opt.step()
            ok = ok and torch.allclose(big.data, ref[0].weight.data, atol=1e-6)  # EVERY row is current after the gather
        q.put((rank, bool(ok), spec.rows(128)))
    finally:
        dist.destroy_process_group()


def test_fp32_row_sharded_parameter_two_ranks_gloo():
    """fp32 mode: rows of the large parameter are updated by their owner and all-gathered in place every step."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_fp32_rows, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert {r for _, _, r in res} == {(0, 64), (64, 128)}


def test_gradient_exchange_requires_process_group():
    from predict_pv_yield_b200.dp import GradientExchange

    if dist.is_initialized():
        pytest.skip("a process group is already initialised")
    with pytest.raises(RuntimeError, match="process group"):
        GradientExchange(torch.nn.Linear(2, 2))
