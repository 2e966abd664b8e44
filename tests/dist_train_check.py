"""Run under torchrun on >= 2 GPUs: data-parallel training step (plots sharded, SyncBatchNorm, one flat
gradient all-reduce) must reproduce the single-GPU full-batch gradients.  Prints DIST_TRAIN_OK."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "stratanet2-vegetation-coverage-maps_b200"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import train_loss  # noqa: E402
from model.point_net2 import PointNet2  # noqa: E402
from model.project_to_2d import project_to_plotwise_coverages  # noqa: E402
from sn2 import parallel  # noqa: E402
from sn2.config import default_args  # noqa: E402
from sn2.synth import randomize_bn_, synth_batch  # noqa: E402


def grads(net, args, batch, dev, local_plots, global_plots, reduce=True):
    bucket = parallel.GradBucket(net)
    bucket.zero()
    cov, proba = net({"xyz": batch["xyz"], "cloud": batch["cloud"]})
    pw = project_to_plotwise_coverages(cov, net.last_cloud_device, args)
    z = batch["xyz"][:, 2, :].reshape(-1, 1).double().to(dev)
    pdf = torch.cat([torch.exp(-z), 0.5 * torch.exp(-0.5 * (z - 1.0) ** 2), 0.1 + 0.05 * z], dim=1)
    loss = train_loss(proba, pw, batch["gt"].to(dev), pdf)
    loss.backward()
    if reduce:
        bucket.allreduce(local_plots, global_plots)
    return bucket.flat.clone(), {k: v.clone() for k, v in net.state_dict().items() if "running" in k}


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B, N = 4, 2048
    args = default_args(subsample_size=N, cuda=local)
    full = synth_batch(5, B, N)
    full["gt"] = torch.rand(B, 4, generator=torch.Generator().manual_seed(2))

    torch.manual_seed(0)
    ref = PointNet2(args)
    randomize_bn_(ref)
    ref.train()
    sd0 = {k: v.clone() for k, v in ref.state_dict().items()}
    want, want_stats = grads(ref, args, full, dev, B, B, reduce=False)  # every rank computes the full-batch reference itself

    net = PointNet2(args)
    net.load_state_dict(sd0)
    net.train()
    net = parallel.convert_sync_batchnorm(net)
    mine = parallel.shard_plots(full, rank, world)
    got, got_stats = grads(net, args, mine, dev, mine["cloud"].shape[0], B)

    scale = want.abs().max().item()
    err = (got - want).abs().max().item()
    ok = err <= 2e-3 * scale
    for k in want_stats:
        ok = ok and torch.allclose(got_stats[k], want_stats[k], rtol=1e-3, atol=1e-5)
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"world={world} max grad err {err:.3e} (scale {scale:.3e})")
        print("DIST_TRAIN_OK" if flag.item() == 1.0 else "DIST_TRAIN_FAIL")
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
