"""Per-tensor gradient error of the CUDA training step against the CPU oracle in float32 AND float64 (noise floor),
with the reference loss (sn2.losses.training_loss).  Usage: python tools/diag_train_oracle.py [B N seed]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stratanet2-vegetation-coverage-maps_b200")]
import torch  # noqa: E402

from model.point_net2 import PointNet2  # noqa: E402
from model.project_to_2d import project_to_plotwise_coverages  # noqa: E402
from oracle.pointnet2_port import PointNet2Port, project_to_plotwise_coverages_port  # noqa: E402
from sn2 import losses  # noqa: E402
from sn2.config import default_args  # noqa: E402
from sn2.synth import randomize_bn_, synth_batch  # noqa: E402

B, N, seed = (int(a) for a in (sys.argv[1:4] + ["4", "2048", "5"][len(sys.argv) - 1:]))
dev = torch.device("cuda", 0)
args = default_args(subsample_size=N, cuda=0)
full = synth_batch(seed, B, N)
full["gt"] = torch.rand(B, 4, generator=torch.Generator().manual_seed(seed))
z = full["xyz"][:, 2, :].reshape(-1, 1).double()
pdf = torch.cat([torch.exp(-z), 0.5 * torch.exp(-0.5 * (z - 1.0) ** 2), 0.1 + 0.05 * z], dim=1)
torch.manual_seed(0)
net = PointNet2(args)
randomize_bn_(net)
net.train()
sd0 = {k: v.clone().cpu() for k, v in net.state_dict().items()}


def oracle(dtype):
    port = PointNet2Port(default_args(subsample_size=N))
    port.load_state_dict(sd0)
    port = port.to(dtype)
    port.train()
    # the restated index ops (fps / radius / knn) convert positions to fp32 internally: index decisions do not change
    data = {"xyz": full["xyz"].to(dtype), "cloud": full["cloud"].to(dtype)}
    cov, proba = port(data)
    pw = project_to_plotwise_coverages_port(cov, full["cloud"], default_args(subsample_size=N))
    loss = losses.training_loss(pw, full["gt"].to(dtype), proba, pdf, fused=False)[0]
    loss.backward()
    return {k: p.grad.double() for k, p in port.named_parameters()}, float(loss.detach())


cov, proba = net({"xyz": full["xyz"], "cloud": full["cloud"]})
pw = project_to_plotwise_coverages(cov, net.last_cloud_device, args)
loss = losses.training_loss(pw, full["gt"].to(dev), proba, pdf.to(dev))[0]
loss.backward()
g32, l32 = oracle(torch.float32)
try:
    g64, l64 = oracle(torch.float64)
except Exception as e:  # noqa: BLE001
    print("fp64 oracle failed:", e)
    g64, l64 = None, None
print(f"loss cuda {float(loss):.7f} oracle32 {l32:.7f} oracle64 {l64}")
for k, p in net.named_parameters():
    w = g32[k]
    sc = w.abs().max().item() + 1e-30
    e1 = (p.grad.cpu().double() - w).abs().max().item() / sc
    e2 = (g64[k] - w).abs().max().item() / sc if g64 is not None else float("nan")
    e3 = (p.grad.cpu().double() - g64[k]).abs().max().item() / sc if g64 is not None else float("nan")
    print(f"{k:45s} scale {sc:.3e}  cuda-vs-o32 {e1:.2e}  o64-vs-o32 {e2:.2e}  cuda-vs-o64 {e3:.2e}")
