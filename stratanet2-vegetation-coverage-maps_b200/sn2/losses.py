"""Training-step loss of the reference on the device.

Reference: ``loss = get_absolute_loss(pred, gt) + args.m * get_NLL_loss(proba, clouds, args)[0] + args.e *
get_entropy_loss(proba)`` (learning/train.py:58-62, learning/loss_functions.py:9-57), where the NLL evaluates three
KDE densities of z by ``scipy.interpolate.interp1d`` on the CPU every step (learning/kde_mixture.py:64-75) and ships
the float64 result to the GPU.

Here: ``KdeLut`` holds the KDE grid on the device and interpolates it in a kernel; ``pointwise_losses`` computes the
NLL (fp64, as the reference) and the entropy in one pass with a one-pass backward that writes d loss / d proba; the
plot-wise MAE is a handful of [B,3] torch ops.  ``training_loss`` composes them with the reference weights.
The ``get_*`` functions keep the reference names and return values for callers that want them separately.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, ops
from ._lib import check, dptr, stream_ptr

EPS = 0.0001  # learning/loss_functions.py:6


class KdeLut:
    """Device copy of the KDE mixture's evaluation grid: knots X [K] and densities Y [3, K] (ground, medium, high), as
    ``KdeMixture.evaluate_kdes`` produces them (learning/kde_mixture.py:89-100); ``pdf`` is ``KdeMixture.predict`` on
    ``z = cloud[:, 2] * z_max`` for every point of a batch."""

    def __init__(self, X, Y, device):
        X = np.asarray(X, dtype=np.float64).reshape(-1)
        Y = np.asarray(Y, dtype=np.float64).reshape(3, -1)
        if X.shape[0] != Y.shape[1] or X.shape[0] < 2 or np.any(np.diff(X) <= 0):
            raise ValueError("KdeLut: X must be strictly increasing with one column of Y per knot")
        self.X = torch.from_numpy(X).to(device)
        self.Y = torch.from_numpy(np.ascontiguousarray(Y)).to(device)

    @classmethod
    def from_kde_mixture(cls, kde_mixture, device):
        X, y1, y2, y3 = kde_mixture.evaluate_kdes()
        return cls(X, np.stack([y1, y2, y3]), device)

    def pdf(self, cloud_dev: torch.Tensor, z_max: float) -> torch.Tensor:
        """cloud (B,F,N) device fp32 -> pdf (B*N, 3) float64."""
        lib = _lib.load()
        B, F, N = cloud_dev.shape
        out = torch.empty((B * N, 3), dtype=torch.float64, device=cloud_dev.device)
        check(lib.sn2_kde_lut(dptr(cloud_dev, torch.float32), B, F, N, float(z_max), dptr(self.X), dptr(self.Y), self.X.numel(),
                              dptr(out), stream_ptr()), "sn2_kde_lut")
        ops._count(1)
        return out


class _PointwiseLosses(torch.autograd.Function):
    @staticmethod
    def forward(ctx, proba, pdf):
        lib = _lib.load()
        proba = proba.contiguous()
        R = proba.shape[0]
        sums = torch.empty(2, dtype=torch.float64, device=proba.device)
        out = torch.empty(2, dtype=torch.float64, device=proba.device)
        check(lib.sn2_pointwise_loss_fwd(dptr(proba, torch.float32), dptr(pdf, torch.float64), R, dptr(sums), dptr(out), stream_ptr()),
              "sn2_pointwise_loss_fwd")
        ops._count(2)
        ctx.save_for_backward(proba, pdf)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        proba, pdf = ctx.saved_tensors
        dproba = torch.empty_like(proba)
        check(lib.sn2_pointwise_loss_bwd(dptr(proba), dptr(pdf), dptr(g.contiguous(), torch.float64), proba.shape[0], dptr(dproba),
                                         stream_ptr()), "sn2_pointwise_loss_bwd")
        ops._count(1)
        return dproba, None


def pointwise_losses(proba: torch.Tensor, pdf: torch.Tensor) -> torch.Tensor:
    """proba (R,4) fp32, pdf (R,3) float64 -> float64 [2] = (NLL mean, entropy mean), differentiable w.r.t. proba."""
    if pdf.dtype != torch.float64 or pdf.shape != (proba.shape[0], 3):
        raise RuntimeError("sn2 pointwise_losses: pdf must be float64 [R, 3]")
    return _PointwiseLosses.apply(proba, pdf.contiguous())


def get_absolute_loss_by_strata(pred_pl, gt):
    """learning/loss_functions.py:9-11 (columns 0, 2, 3 as slices: no index kernels, CUDA-graph safe)."""
    sel = lambda t: torch.cat([t[:, :1], t[:, 2:4]], dim=1)  # noqa: E731
    return ((sel(pred_pl) - sel(gt)).pow(2) + EPS).pow(0.5).mean(0)


def get_absolute_loss(pred_pl, gt):
    """learning/loss_functions.py:14-16."""
    return get_absolute_loss_by_strata(pred_pl, gt).mean()


def get_entropy_loss(pred_pixels):
    """learning/loss_functions.py:19-24."""
    p = pred_pixels[:, 2:]
    return -(p * torch.log(p + EPS) + (1 - p) * torch.log(1 - p + EPS)).mean()


def get_NLL_loss(pred_pointwise, pdf_all):
    """learning/loss_functions.py:27-57 with the pdf already evaluated (KdeLut.pdf): -> (loss, (p_all, pdf_all))."""
    p_all = torch.stack([pred_pointwise[:, :2].sum(1), pred_pointwise[:, 2], pred_pointwise[:, 3]], dim=1)
    return -torch.log((p_all * pdf_all).sum(1)).mean(), (p_all, pdf_all)


def training_loss(pred_coverages, gt_coverages, proba_pointwise, pdf_all, m: float = 0.10, e: float = 0.2 / 5, fused: bool = True):
    """learning/train.py:58-62.  -> (loss, loss_abs, loss_log, loss_e); float64 like the reference's sum
    (its NLL term is float64).  fused=False evaluates the same formulas with torch ops (the checker of the kernels)."""
    loss_abs = get_absolute_loss(pred_coverages, gt_coverages)
    if fused:
        pl = pointwise_losses(proba_pointwise, pdf_all)
        loss_log, loss_e = pl[0], pl[1]
    else:
        loss_log = get_NLL_loss(proba_pointwise, pdf_all)[0]
        loss_e = get_entropy_loss(proba_pointwise)
    return loss_abs + m * loss_log + e * loss_e, loss_abs, loss_log, loss_e
