"""Synthetic 10 m-radius plots -- the workload every config of BASELINE.json uses.

Shapes follow the reference input contract (``load_cloud`` output consumed by ``PointNet2.forward``,
/root/reference/data_loader/loader.py:73-87, /root/reference/model/point_net2.py:107-108):
``xyz`` (B,3,N) fp32 metres, ``cloud`` (B,10,N) fp32 normalised rows
(x/10, y/10, z/24.24, 7 feature rows), including the 316 zero-feature fake ground points of
/root/reference/data_loader/loader.py:90-105.
"""
from __future__ import annotations

import numpy as np
import torch


def _fake_ground(diam_meters: int = 20) -> torch.Tensor:
    c = torch.arange(-diam_meters // 2, diam_meters // 2, 1, dtype=torch.float32) + 0.5
    xx, yy = torch.meshgrid(c, c, indexing="xy")
    x, y = xx.flatten(), yy.flatten()
    keep = torch.sqrt(x * x + y * y) < (diam_meters // 2)
    return torch.stack([x[keep], y[keep], torch.zeros(int(keep.sum()))])  # (3, 316)


def synth_plot(seed: int, n: int, variant: str = "plain"):
    """One plot -> (xyz (3,n), cloud (10,n)) fp32.  variant: plain | cm | dup."""
    g = torch.Generator().manual_seed(int(seed))
    n_raw = 6000 if variant == "dup" else n
    u = torch.rand(n_raw, generator=g)
    th = torch.rand(n_raw, generator=g) * (2 * np.pi)
    r = 10.0 * torch.sqrt(u)
    x, y = r * torch.cos(th), r * torch.sin(th)
    sel = torch.rand(n_raw, generator=g)
    z_ground = torch.randn(n_raw, generator=g).abs() * 0.05
    z_low = torch.rand(n_raw, generator=g) * 0.5
    z_med = 0.5 + torch.rand(n_raw, generator=g) * 1.0
    z_high = 1.5 + torch.rand(n_raw, generator=g) * 13.5
    z = torch.where(sel < 0.55, z_ground, torch.where(sel < 0.75, z_low, torch.where(sel < 0.85, z_med, z_high)))
    feats = torch.rand(7, n_raw, generator=g)
    xyz = torch.stack([x, y, z]).to(torch.float32)
    fake = _fake_ground()
    xyz = torch.cat([xyz, fake], dim=1)
    feats = torch.cat([feats, torch.zeros(7, fake.shape[1])], dim=1)
    tot = xyz.shape[1]
    if tot >= n:
        pick = torch.randperm(tot, generator=g)[:n]
    else:  # up-sample with replacement, /root/reference/data_loader/loader.py:239-244
        pick = torch.cat([torch.arange(tot), torch.randint(0, tot, (n - tot,), generator=g)])
    xyz, feats = xyz[:, pick], feats[:, pick]
    if variant == "cm":
        xyz = torch.round(xyz * 100.0) / 100.0
    cloud = torch.cat([xyz[0:1] / 10.0, xyz[1:2] / 10.0, xyz[2:3] / 24.24, feats], dim=0)
    return xyz.to(torch.float32).contiguous(), cloud.to(torch.float32).contiguous()


def synth_batch(config: int, B: int, n: int, variant: str = "plain", first_plot: int = 0):
    """B plots with seeds 1000*config + plot_index -> {"xyz": (B,3,n), "cloud": (B,10,n)}."""
    plots = [synth_plot(1000 * config + first_plot + i, n, variant) for i in range(B)]
    return {
        "xyz": torch.stack([p[0] for p in plots]).contiguous(),
        "cloud": torch.stack([p[1] for p in plots]).contiguous(),
    }


def randomize_bn_(model: torch.nn.Module, seed: int = 1) -> None:
    """Make eval-mode BatchNorm non-trivial: random running stats, a few negative
    gammas, so a wrong Linear/ReLU/BN order or a BN commuted with max cannot pass parity."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            c = m.num_features
            with torch.no_grad():
                m.running_mean.copy_(torch.randn(c, generator=g) * 0.1)
                m.running_var.copy_(0.5 + torch.rand(c, generator=g))
                gamma = 0.5 + torch.rand(c, generator=g)
                gamma[torch.rand(c, generator=g) < 0.2] *= -1.0
                m.weight.copy_(gamma)
                m.bias.copy_(torch.randn(c, generator=g) * 0.1)
