"""ncu target: two eager config-3-shaped training steps (B plots x N points) -- the second one is the capture window."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stratanet2-vegetation-coverage-maps_b200")]
import torch  # noqa: E402

import bench  # noqa: E402
from model.project_to_2d import project_to_plotwise_coverages  # noqa: E402
from sn2 import losses  # noqa: E402
from sn2.optim import FusedAdam  # noqa: E402
from sn2.synth import synth_batch  # noqa: E402

B, N = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda", 0)
args, net = bench.make_model(N, 0)
net.train()
opt = FusedAdam(net.parameters(), lr=args.lr, weight_decay=args.wd)
lut = bench.synthetic_kde_lut(dev)
d = {k: v.to(dev) for k, v in synth_batch(3, B, N).items()}
gt = torch.rand(B, 4, device=dev)
for _ in range(1 if os.environ.get("SN2_NCU_ONE") == "1" else 2):
    opt.zero_grad()
    cov, proba = net(d)
    pw = project_to_plotwise_coverages(cov, net.last_cloud_device, args)
    loss = losses.training_loss(pw, gt, proba, lut.pdf(net.last_cloud_device, args.z_max), args.m, args.e)[0]
    loss.backward()
    opt.step()
    torch.cuda.synchronize()
print("ok", float(loss))
