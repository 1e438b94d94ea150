"""Data-parallel check of the two-tower model (run under torchrun, 2+ ranks): a few train steps with the gradient exchange,
with the fc1 / nwp_fc1 optimisers sharded (--shard) or replicated; rank 0 prints the loss trajectory and parameter
checksums, which must agree between the two modes and between ranks.
torchrun --nproc-per-node 2 tools/dp_check_sat_nwp.py --precision bf16 [--shard]"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--shard", action="store_true")
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    from predict_pv_yield_b200.dp import GradientExchange
    from predict_pv_yield_b200.models.conv3d.model_sat_nwp import Model

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B = 8
    kw = dict(forecast_minutes=30, history_minutes=60, image_size_pixels=24, nwp_image_size_pixels=24)
    torch.manual_seed(0)
    m = Model(**kw, precision=args.precision).to(dev)
    m.batch_size = B
    opt = m.configure_optimizers()
    ex = GradientExchange(m, shard_large=args.shard, large_numel=1 << 20)
    ex.attach_optimizer(opt)
    rs = np.random.RandomState(100 + rank)
    batch = {"satellite": {"data": torch.from_numpy(rs.randint(0, 1024, size=(B, 12, 19, 24, 24)).astype(np.int16)).to(dev)},
             "nwp": {"data": torch.from_numpy(rs.randn(B, 10, 2, 24, 24).astype(np.float32)).to(dev)},
             "pv": {"pv_yield": torch.from_numpy(rs.rand(B, 19, 128).astype(np.float32)).to(dev),
                    "pv_system_row_number": torch.from_numpy(rs.randint(0, 940, size=(B, 128)).astype(np.int64)).to(dev)}}
    losses = []
    for i in range(args.steps):
        opt.zero_grad()
        loss = m.training_step(batch, i)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    sd = m.state_dict()  # gathers the sharded master rows (collective)
    sums = torch.stack([sd[k].double().abs().sum() for k in ("fc1.weight", "nwp_fc1.weight", "sat_conv0.weight", "fc4.weight")])
    other = sums.clone()
    dist.broadcast(other, src=0)
    same = bool(torch.allclose(sums, other, rtol=1e-12))
    with torch.no_grad():
        y = m(batch)
    if rank == 0:
        print(f"shard={args.shard} precision={args.precision} losses={['%.7f' % v for v in losses]} "
              f"checksums={[('%.9e' % float(v)) for v in sums]} replicas_identical={same} bytes_reduced={ex.bytes_reduced_last_step}")
    assert same and torch.isfinite(y).all()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
