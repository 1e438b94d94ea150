"""Multi-rank GPU test of the data-parallel step (needs >= 2 visible GPUs; skipped on a one-GPU box).

Two NCCL ranks train the golden NWP + PV-history case for three steps, each on its own half of the batch, in three ways:
fp32 with the in-place all-reduce of fc1.weight.grad, fp32 with the row-sharded optimiser, bf16 with the replicated optimiser and bf16 with the optimiser of
fc1.weight sharded by rows (reduce-scatter -> row-wise fused Adam -> all-gather of the bf16 shadow).  Checked: the replicas
stay bit-identical, sharded == replicated (same losses, same weights after gathering the master rows), the first loss
equals the CPU oracle's on the full batch (mean of the two half-batch losses), the optimizer state gathered for a
checkpoint is complete, and the logged scalars are the mean over the ranks (sync_dist=True, base_model.py:108-119).
"""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


BATCH = 4


def _full_batch_and_state():
    """The golden NWP + PV-history model (16 x 16 crops) on a seeded batch of 4; identical on every rank and in the parent."""
    from oracle import conv3d_oracle as O
    from oracle.golden_cases import CASES

    torch.manual_seed(7)
    om = O.OracleModel(**CASES["nwp_pv_small"]["model"])
    return O.make_synthetic_batch(BATCH, 12, 19, 16, seed=11), {k: v.clone() for k, v in om.state_dict().items()}


def _worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from oracle import conv3d_oracle as O
        from oracle.golden_cases import CASES
        from predict_pv_yield_b200.dp import GradientExchange
        from predict_pv_yield_b200.models.conv3d.model import Model

        case = CASES["nwp_pv_small"]
        full, state = _full_batch_and_state()
        Bf = BATCH
        per = Bf // world
        rows = slice(rank * per, (rank + 1) * per)

        def shard(d):
            return {k: (shard(v) if isinstance(v, dict) else v[rows]) for k, v in d.items()}

        batch = O.batch_to(shard(full), dev)
        out = {}
        for mode, precision, shard_large in (("fp32", "fp32", False), ("fp32_sharded", "fp32", True),
                                             ("bf16_replicated", "bf16", False), ("bf16_sharded", "bf16", True)):
            m = Model(**case["model"], precision=precision).to(dev)
            m.batch_size = per
            m.load_state_dict(state)
            opt = m.configure_optimizers()
            ex = GradientExchange(m, shard_large=shard_large, large_numel=1 << 16)
            ex.attach_optimizer(opt)
            losses, logged = [], []
            for i in range(3):
                opt.zero_grad()
                loss = m.training_step(batch, i)
                loss.backward()
                opt.step()
                losses.append(float(loss.detach()))
                logged.append(float(m.logged_metrics["NMAE/Train"]))
            osd = opt.state_dict()  # collective when sharded: gathers the moments of the rows other ranks own
            sd = {k: v.detach().clone() for k, v in m.state_dict().items()}  # collective when sharded: gathers the master rows
            flat = torch.cat([v.double().reshape(-1) for v in sd.values()])
            other = flat.clone()
            dist.broadcast(other, src=0)
            fc1_idx = [i for i, (k, _) in enumerate(m.named_parameters()) if k == "fc1.weight"][0]
            exp_avg = osd["state"][fc1_idx]["exp_avg"]
            out[mode] = dict(losses=losses, logged=logged, identical=bool(torch.equal(flat, other)),
                             fc1=sd["fc1.weight"].cpu(), conv0=sd["sat_conv0.weight"].cpu(), exp_avg=exp_avg.detach().cpu(),
                             sharded=bool(ex._sharded), bytes=ex.bytes_reduced_last_step)
            ex.remove()
        # the oracle's loss on the FULL batch = mean over the ranks of the half-batch losses
        local = torch.tensor([out["fp32"]["losses"][0]], device=dev, dtype=torch.float64)
        dist.all_reduce(local)
        out["mean_first_loss"] = float(local) / world
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_two_rank_training_sharded_equals_replicated_equals_oracle():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp

    from oracle import conv3d_oracle as O
    from oracle.golden_cases import CASES

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    case = CASES["nwp_pv_small"]
    om = O.OracleModel(**case["model"])
    om.batch_size = BATCH
    full, state = _full_batch_and_state()
    om.load_state_dict(state)
    want = float(om.training_step(full, 0))
    for rank, out in res.items():
        assert abs(out["mean_first_loss"] - want) <= 1e-5 * abs(want), rank
        for mode in ("fp32", "fp32_sharded", "bf16_replicated", "bf16_sharded"):
            assert out[mode]["identical"], (rank, mode)
            # sync_dist: the logged loss is the mean over the ranks, the returned loss is the local one
            assert out[mode]["logged"][0] == pytest.approx(0.5 * (res[0][mode]["losses"][0] + res[1][mode]["losses"][0]), rel=1e-6)
        assert out["bf16_sharded"]["sharded"] and not out["bf16_replicated"]["sharded"]
        # fp32: row-sharded Adam + in-place all-gather of the updated rows == replicated Adam after an all-reduce, bit for bit
        assert out["fp32_sharded"]["sharded"] and not out["fp32"]["sharded"]
        assert out["fp32_sharded"]["bytes"] < out["fp32"]["bytes"]
        assert out["fp32_sharded"]["losses"] == out["fp32"]["losses"], rank
        assert torch.equal(out["fp32_sharded"]["fc1"], out["fp32"]["fc1"])
        assert torch.equal(out["fp32_sharded"]["exp_avg"], out["fp32"]["exp_avg"])
        assert out["bf16_sharded"]["bytes"] < out["bf16_replicated"]["bytes"]
        assert out["bf16_sharded"]["losses"] == out["bf16_replicated"]["losses"], rank
        assert torch.equal(out["bf16_sharded"]["fc1"], out["bf16_replicated"]["fc1"])
        assert torch.equal(out["bf16_sharded"]["conv0"], out["bf16_replicated"]["conv0"])
        assert torch.equal(out["bf16_sharded"]["exp_avg"], out["bf16_replicated"]["exp_avg"])  # gathered for the checkpoint
        assert abs(out["bf16_sharded"]["losses"][0] - out["fp32"]["losses"][0]) <= 2e-2 * abs(out["fp32"]["losses"][0])
