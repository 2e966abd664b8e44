"""Parcel-scale inference on the device (BASELINE config 4).

The reference runs a parcel as three scripts with files in between: ``prepare.py`` tiles the LAS cloud into
overlapping 10 m plots on the CPU (cKDTree ball query, <= 50-point filter, sklearn radius search + a Python loop for
the local-min-z normalisation) and pickles them; ``predict.py`` reloads them through a DataLoader (centering, fake
ground points, rescaling, sub-sampling per plot in numpy), runs the model, writes one GeoTIFF per plot and merges them
with rasterio.  Here the parcel cloud is uploaded once; tiling + plot preparation is one kernel launch
(csrc/parcel.cu), its output is the model input, rasters are fused on the device (sn2/fusion.py) and the band
finalisation (hard medium-vegetation band, NaN rules) is three small kernels.

Reference: inference/prepare_utils.py:47-81, 95-165; prepare.py:60-98; utils/load_data.py:149-184, 228-249;
data_loader/loader.py:73-158, 233-255; predict.py:92-142; inference/geotiff_raster.py:121-146, 273-291.
Not rebuilt: LAS / shapefile / GeoTIFF IO, the admissibility band (rasterio sieve + polygonise + shapely buffer).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib, ops
from ._lib import check, dptr, stream_ptr

MIN_N_POINTS_FOR_INFERENCE = 50      # inference/prepare_utils.py:66; prepare.py:93 keeps only plots with MORE points
LAS_PARCEL_BUFFER = 20               # inference/prepare_utils.py:146
PLOT_RADIUS_METERS = 10              # inference/prepare_utils.py:117 ("hardcoded, should not change at any time")


def plot_centers_reference(x_min, x_max, y_min, y_max, args) -> np.ndarray:
    """Plot centres exactly as ``divide_parcel_las_and_get_disk_centers`` lists them (prepare_utils.py:116-144), before
    the shape filter: the first centre appears twice (the list is seeded with it and the loop visits it again).
    x_min ... are the float32 extrema of the LAS cloud; the arithmetic is float64 as in the reference."""
    x_min, x_max, y_min, y_max = (float(np.float32(v)) for v in (x_min, x_max, y_min, y_max))
    width = 2 * math.cos(math.pi / 4) * PLOT_RADIUS_METERS
    movement = width - 1 * (2 * PLOT_RADIUS_METERS) / args.diam_pix
    # the reference divides the float32 range by the float64 movement
    nx = math.ceil(float(np.float32(x_max) - np.float32(x_min)) / movement) + 1
    ny = math.ceil(float(np.float32(y_max) - np.float32(y_min)) / movement) + 1
    sx, sy = x_min + movement / 4, y_min + movement / 4
    out = [[sx, sy]]
    for i in range(nx):
        for j in range(ny):
            out.append([sx + i * movement, sy + j * movement])
    return np.asarray(out, dtype=np.float64)


def keep_points_in_shape(centers: np.ndarray, polygon_xy: np.ndarray, inclusion_buffer: float) -> np.ndarray:
    """Mask of the centres inside ``polygon.buffer(inclusion_buffer)`` (prepare_utils.py:146-151, 168-176): inside the
    polygon (even-odd rule) or within the buffer distance of its boundary.  polygon_xy: (V,2) vertices of one ring."""
    P = np.asarray(polygon_xy, dtype=np.float64)
    A, B = P, np.roll(P, -1, axis=0)
    c = np.asarray(centers, dtype=np.float64)
    x, y = c[:, 0:1], c[:, 1:2]
    with np.errstate(divide="ignore", invalid="ignore"):
        cross = ((A[:, 1] > y) != (B[:, 1] > y)) & (x < (B[:, 0] - A[:, 0]) * (y - A[:, 1]) / (B[:, 1] - A[:, 1]) + A[:, 0])
    inside = cross.sum(axis=1) % 2 == 1
    d = B - A
    t = np.clip(((x - A[:, 0]) * d[:, 0] + (y - A[:, 1]) * d[:, 1]) / np.maximum((d ** 2).sum(1), 1e-300), 0.0, 1.0)
    dist = np.sqrt((x - (A[:, 0] + t * d[:, 0])) ** 2 + (y - (A[:, 1] + t * d[:, 1])) ** 2).min(axis=1)
    return inside | (dist <= inclusion_buffer)


class ParcelCloud:
    """A parcel's LAS cloud in HBM: float32 rows (x, y, z, R, G, B, NIR, intensity, return_num, num_returns) as
    ``load_las_file`` returns them (utils/load_data.py:149-184), plus a cell-grouped float4 copy of (x, y, z, index) over a
    uniform xy grid (cells of 1.6 m >= 1.05 x the z-normalisation radius: a point's 3 x 3 cells hold all its neighbours)."""

    def __init__(self, cloud, device, cell: float = 1.6):
        cloud = torch.as_tensor(cloud)
        if cloud.dim() != 2 or cloud.shape[0] != 10 or cloud.dtype != torch.float32:
            raise RuntimeError("sn2 ParcelCloud: expected a float32 [10, P] array (the layout of load_las_file)")
        lib = _lib.load()
        self.device = torch.device(device)
        c = cloud.to(self.device, non_blocking=True).contiguous()
        self.P = int(c.shape[1])
        self.xyz, self.feat = c[:3].contiguous(), c[3:].contiguous()
        mn, mx = self.xyz[:2].min(dim=1).values.cpu(), self.xyz[:2].max(dim=1).values.cpu()  # get_xy_range (prepare_utils.py:39-44)
        self.x_min, self.y_min, self.x_max, self.y_max = float(mn[0]), float(mn[1]), float(mx[0]), float(mx[1])
        self.cell = float(cell)
        self.nx = max(1, int(math.floor((self.x_max - self.x_min) / cell)) + 1)
        self.ny = max(1, int(math.floor((self.y_max - self.y_min) / cell)) + 1)
        i32 = lambda n: torch.empty(n, dtype=torch.int32, device=self.device)  # noqa: E731
        cell_of, count, cursor = i32(self.P), i32(self.nx * self.ny), i32(self.nx * self.ny)
        self.cell_start, self.pos_of = i32(self.nx * self.ny + 1), i32(self.P)
        self.sorted4 = torch.empty((self.P, 4), dtype=torch.float32, device=self.device)
        check(lib.sn2_parcel_grid_build(dptr(self.xyz[0]), dptr(self.xyz[1]), dptr(self.xyz[2]), self.P, self.x_min, self.y_min, self.cell,
                                        self.nx, self.ny, dptr(cell_of), dptr(count), dptr(self.cell_start), dptr(cursor),
                                        dptr(self.sorted4), dptr(self.pos_of), stream_ptr()), "sn2_parcel_grid_build")
        ops._count(3)


def extract_plots(parcel: ParcelCloud, centers, args, seeds=None, want_src: bool = False):
    """All plots around `centers` ((C,2) float64) in one launch -> dict: ``xyz`` (C,3,S), ``cloud`` (C,10,S) fp32 on the
    device (the model input: centred, z-normalised, fake ground points added, rescaled, sub-sampled to
    args.subsample_size), ``n_points`` (C,) int32 = parcel points inside each disk, ``valid`` = n_points > 50.
    Rows of plots that are not valid are undefined.  seeds: per-plot uint32 of the sampling hash (default: the index)."""
    lib = _lib.load()
    dev = parcel.device
    cen = torch.as_tensor(np.ascontiguousarray(np.asarray(centers, dtype=np.float64))).to(dev)
    C, S = int(cen.shape[0]), int(args.subsample_size)
    sd = np.arange(C, dtype=np.uint32) if seeds is None else np.asarray(seeds, dtype=np.uint32)
    sd = torch.from_numpy(sd.view(np.int32).copy()).to(dev)  # uint32 bit patterns
    xyz = torch.empty((C, 3, S), dtype=torch.float32, device=dev)
    cloud = torch.empty((C, 10, S), dtype=torch.float32, device=dev)
    n = torch.empty(C, dtype=torch.int32, device=dev)
    src = torch.empty((C, S), dtype=torch.int32, device=dev) if want_src else None
    check(lib.sn2_extract_plots(dptr(parcel.xyz), dptr(parcel.feat), parcel.P, parcel.x_min, parcel.y_min, parcel.cell, parcel.nx,
                                parcel.ny, dptr(parcel.cell_start), dptr(parcel.sorted4), dptr(parcel.pos_of), dptr(cen), dptr(sd), C, S,
                                float(args.diam_meters // 2), float(args.znorm_radius_in_meters), float(args.z_max),
                                int(args.diam_meters), MIN_N_POINTS_FOR_INFERENCE, dptr(xyz), dptr(cloud), dptr(n), dptr(src),
                                stream_ptr()), "sn2_extract_plots")
    ops._count(1)
    out = {"xyz": xyz, "cloud": cloud, "n_points": n, "valid": n > MIN_N_POINTS_FOR_INFERENCE}
    if want_src:
        out["src"] = src
    return out


def finalize_mosaic(fused: torch.Tensor):
    """[4,H,W] float64 (Vb, Vm, Vh, weights; NaN = no value) -> ([5,H,W] = (Vb, Vm_soft, Vh, Vm_hard, weights), threshold,
    soft coverage): ``finalize_merged_raster`` without the admissibility band (geotiff_raster.py:121-146, 273-291)."""
    lib = _lib.load()
    if fused.dim() != 3 or fused.shape[0] != 4 or fused.dtype != torch.float64:
        raise RuntimeError("sn2 finalize_mosaic: expected a float64 [4, H, W] mosaic")
    fused = fused.contiguous()
    _, H, W = fused.shape
    out = torch.empty((5, H, W), dtype=torch.float64, device=fused.device)
    hist = torch.empty(10002, dtype=torch.int32, device=fused.device)
    scratch = torch.empty(4, dtype=torch.float64, device=fused.device)
    check(lib.sn2_finalize_mosaic(dptr(fused), H, W, dptr(hist), dptr(scratch), dptr(out), stream_ptr()), "sn2_finalize_mosaic")
    ops._count(3)
    return out, scratch[2], scratch[3]


def predict_parcel(model, args, parcel: ParcelCloud, centers, rank: int = 0, world: int = 1, batch: int = 64, depth: int = 3,
                   group=None):
    """predict.py:92-142 for one parcel on the device: this rank's share of the plot centres (contiguous blocks of the
    centre list, so a rank only touches its stripe of the cloud) -> plot extraction -> batched PointNet2 inference +
    rasters (InferencePipeline) -> local-map fusion -> ONE all-reduce of the accumulators -> band finalisation.
    Returns (mosaic [5,H,W] float64 on the device, info dict)."""
    from .fusion import MapFusion, mosaic_frame
    from .parallel import shard_bounds
    from .pipeline import InferencePipeline

    centers = np.asarray(centers, dtype=np.float64)
    D = int(args.diam_pix)
    left, top, H, W, offsets = mosaic_frame(centers, args.diam_meters, D)
    lo, hi = shard_bounds(centers.shape[0], rank, world)
    fus = MapFusion(H, W, D, parcel.device)
    n_valid = 0
    if hi > lo:
        ex = extract_plots(parcel, centers[lo:hi], args, seeds=np.arange(lo, hi, dtype=np.uint32))
        valid = torch.nonzero(ex["valid"]).view(-1)          # the one host synchronisation: how many plots are worth running
        n_valid = int(valid.numel())
        off_dev = torch.from_numpy(offsets[lo:hi]).to(parcel.device)[valid]
        pipe = InferencePipeline(model, args, depth=depth)
        pending = []

        def collect():
            b0, b1, slot = pending.pop(0)
            _pw, rs = pipe.result(slot)
            fus.add(rs, off_dev[b0:b1])

        for b0 in range(0, n_valid, batch):
            b1 = min(b0 + batch, n_valid)
            sel = valid[b0:b1]
            if len(pending) == depth:
                collect()
            pending.append((b0, b1, pipe.submit({"xyz": ex["xyz"][sel], "cloud": ex["cloud"][sel]}, keep_on_device=True)))
        while pending:
            collect()
    fused = fus.finalize(group)
    mosaic, thr, soft = finalize_mosaic(fused)
    return mosaic, {"left": left, "top": top, "H": H, "W": W, "plots": int(hi - lo), "plots_valid": n_valid, "threshold": thr,
                    "soft_medium_coverage": soft}
