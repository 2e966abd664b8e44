"""B200-native ``model.project_to_2d`` -- drop-in for /root/reference/model/project_to_2d.py.

``project_to_plotwise_coverages(pred_pointwise, clouds, args)`` and
``project_to_2d_rasters(cloud, coverages_pointwise, args)`` keep the reference signatures and return
types; both run as one kernel launch per batch (csrc/project.cu).  ``project_to_2d_rasters_batched``
is the batched form the parcel driver should use (one launch for all plots of a batch instead of a
Python loop over plots and pixels).
"""
from __future__ import annotations

import numpy as np
import torch

from sn2 import ops as _ops


def _device_of(args, *tensors):
    for t in tensors:
        if isinstance(t, torch.Tensor) and t.is_cuda:
            return t.device
    if getattr(args, "cuda", None) is None:
        raise RuntimeError("sn2 project_to_2d needs a CUDA device (args.cuda); this build has no CPU path")
    return torch.device("cuda", args.cuda)


def project_to_plotwise_coverages(pred_pointwise, clouds, args):
    """pred_pointwise (B*N,4) device, clouds (B,F,N) host or device -> (B,4) device fp32
    [low, bare = 1 - low, medium, high] (reference :7-55)."""
    dev = _device_of(args, pred_pointwise)
    with torch.cuda.device(dev):
        clouds_d = clouds.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()
        if pred_pointwise.requires_grad and torch.is_grad_enabled():
            from sn2.autograd_ops import ProjectPlotwise

            return ProjectPlotwise.apply(pred_pointwise.to(device=dev, dtype=torch.float32), clouds_d, int(args.diam_pix))
        pred = pred_pointwise.detach().to(device=dev, dtype=torch.float32).contiguous()
        return _ops.project_plotwise(clouds_d, pred, int(args.diam_pix))


def project_to_2d_rasters_batched(clouds, coverages_pointwise, args, layout="point_major"):
    """clouds (B,F,N); coverages (B*N,4) ["point_major", as PointNet2.forward returns them] or (B,4,N)
    ["channel_major", as get_batch_format returns them] -> device float64 (B,3,D,D)."""
    dev = _device_of(args, coverages_pointwise, clouds)
    with torch.cuda.device(dev):
        clouds_d = clouds.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()
        cov = coverages_pointwise.detach().to(device=dev, dtype=torch.float32).contiguous()
        return _ops.project_rasters(clouds_d, cov, layout, int(args.diam_pix), int(args.diam_meters))


def project_to_2d_rasters(cloud: torch.Tensor, coverages_pointwise: torch.Tensor, args) -> np.ndarray:
    """cloud (F>=2,N), coverages_pointwise (4,N) -> numpy float64 (3,D,D), NaN where no point,
    bands low/medium/high, first row = largest y (reference :58-113)."""
    out = project_to_2d_rasters_batched(cloud[None, :2], coverages_pointwise[None], args, layout="channel_major")
    return out[0].cpu().numpy()
