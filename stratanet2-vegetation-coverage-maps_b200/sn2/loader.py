"""Device-side loader transforms: ``augment`` + ``rescale_cloud`` of
/root/reference/data_loader/loader.py:135-214 for a whole batch in one kernel (csrc/loader.cu).

``draw_augmentation`` draws what the reference draws per plot with np.random -- a rotation angle among 360 whole
degrees, two flips, clipped Gaussian noise on xy and on the four colour channels -- with torch's device generator;
``augment_rescale`` applies given draws, so it can be checked draw for draw against the numpy restatement.
Input plots are what ``load_cloud`` has after ``center_cloud`` + ``add_fake_empty_ground_points``: float32 (B,10,N) in
metres / raw feature units (sn2.parcel.extract_plots delivers the eval-time, already rescaled form directly).
"""
from __future__ import annotations

import torch

from . import _lib, ops
from ._lib import check, dptr, stream_ptr


def draw_augmentation(B: int, N: int, device, generator=None):
    """-> (angle float64 [B] radians, flip uint8 [B,2], noise float64 [B,6,N]) as get_xyz_augmentation_params (:217-222)
    and the randn calls of augment (:183-207) would draw them."""
    angle = torch.deg2rad(torch.randint(0, 360, (B,), device=device, generator=generator).to(torch.float64))
    flip = (torch.rand((B, 2), device=device, generator=generator) > 0.5).to(torch.uint8)
    noise = torch.randn((B, 6, N), device=device, dtype=torch.float64, generator=generator)
    return angle, flip, noise


def augment_rescale(raw: torch.Tensor, z_max: float, angle=None, flip=None, noise=None):
    """raw (B,10,N) fp32 device -> {"xyz": (B,3,N), "cloud": (B,10,N)}.  Without draws: rescale only (eval)."""
    lib = _lib.load()
    if raw.dim() != 3 or raw.shape[1] != 10 or raw.dtype != torch.float32:
        raise RuntimeError("sn2 augment_rescale: expected float32 (B, 10, N)")
    raw = raw.contiguous()
    B, _, N = raw.shape
    xyz = torch.empty((B, 3, N), dtype=torch.float32, device=raw.device)
    cloud = torch.empty((B, 10, N), dtype=torch.float32, device=raw.device)
    check(lib.sn2_augment_rescale(dptr(raw), B, N, dptr(angle, torch.float64) if angle is not None else None,
                                  dptr(flip.contiguous(), torch.uint8) if flip is not None else None,
                                  dptr(noise.contiguous(), torch.float64) if noise is not None else None, float(z_max), dptr(xyz),
                                  dptr(cloud), stream_ptr()), "sn2_augment_rescale")
    ops._count(1)
    return {"xyz": xyz, "cloud": cloud}
