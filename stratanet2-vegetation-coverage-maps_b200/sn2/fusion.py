"""Parcel-scale inference helpers (BASELINE config 4): plot-centre grid of the reference tiling and GPU
local-map fusion of per-plot rasters.

Reference: centres -- /root/reference/inference/prepare_utils.py:95-151 (step 2*cos45*10 - diam_m/diam_pix,
start = min + step/4); geotransform of a plot raster -- /root/reference/inference/geotiff_raster.py:46-61;
mosaic placement -- rasterio.merge (un-vendored, rasterio==1.2.6): integer offsets
round((top - plot_top)/res), round((plot_left - left)/res); weights and averaging rule --
geotiff_raster.py:103-118, 294-347.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, ops
from ._lib import check, dptr, stream_ptr


def plot_centers(x_min, x_max, y_min, y_max, diam_meters=20, diam_pix=20, radius=10.0):
    """Grid of plot centres covering [x_min,x_max] x [y_min,y_max] (prepare_utils.py:110-144), (P,2) float64."""
    step = 2 * math.cos(math.pi / 4) * radius - diam_meters / diam_pix
    nx = math.ceil((x_max - x_min) / step) + 1
    ny = math.ceil((y_max - y_min) / step) + 1
    xs = x_min + step / 4 + step * np.arange(nx)
    ys = y_min + step / 4 + step * np.arange(ny)
    gx, gy = np.meshgrid(xs, ys, indexing="ij")
    return np.stack([gx.ravel(), gy.ravel()], axis=1)


def mosaic_frame(centers: np.ndarray, diam_meters=20, diam_pix=20):
    """-> (left, top, H, W, offsets int32 [P,2]) of the merged raster, as rasterio.merge lays it out."""
    res = diam_meters / diam_pix
    half = diam_meters // 2
    left, right = centers[:, 0].min() - half, centers[:, 0].max() + half
    bottom, top = centers[:, 1].min() - half, centers[:, 1].max() + half
    W = int(round((right - left) / res))
    H = int(round((top - bottom) / res))
    roff = np.array([int(round((top - (cy + half)) / res)) for cy in centers[:, 1]], dtype=np.int32)
    coff = np.array([int(round(((cx - half) - left) / res)) for cx in centers[:, 0]], dtype=np.int32)
    return left, top, H, W, np.stack([roff, coff], axis=1)


class MapFusion:
    """Accumulates plot rasters into a parcel grid on the device; ``finalize`` (after an optional NCCL
    all-reduce of the accumulators across plot-sharded ranks) returns the [4,H,W] float64 mosaic."""

    def __init__(self, H: int, W: int, D: int, device):
        self.H, self.W, self.D = H, W, D
        self.acc = torch.zeros((7, H, W), dtype=torch.float64, device=device)  # num[3], den[3], wsum

    def add(self, rasters: torch.Tensor, offsets: torch.Tensor):
        lib = _lib.load()
        P = rasters.shape[0]
        num, den, ws = self.acc[0:3], self.acc[3:6], self.acc[6]
        check(lib.sn2_fuse_accumulate(dptr(rasters.contiguous(), torch.float64), dptr(offsets.contiguous(), torch.int32), P,
                                      self.D, self.H, self.W, dptr(num), dptr(den), dptr(ws), stream_ptr()),
              "sn2_fuse_accumulate")
        ops._count(1)

    def finalize(self, group=None) -> torch.Tensor:
        lib = _lib.load()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.acc, op=dist.ReduceOp.SUM, group=group)  # the one collective of parcel inference
        out = torch.empty((4, self.H, self.W), dtype=torch.float64, device=self.acc.device)
        check(lib.sn2_fuse_finalize(dptr(self.acc[0:3]), dptr(self.acc[3:6]), dptr(self.acc[6]), self.H, self.W, dptr(out),
                                    stream_ptr()), "sn2_fuse_finalize")
        ops._count(1)
        return out
