"""CPU oracle: restated PointNet2 forward and 2D projections.  TEST INFRASTRUCTURE ONLY.

An independent restatement (not a copy) of the reference glue, written against the restated
third-party ops in oracle/thirdparty_ops.py.  Each function cites the reference lines it follows.
It is validated against the reference's own files run verbatim (oracle/ref_loader.py) in
tests/test_oracle_port.py and pinned by tests/golden/.

Why a port exists next to ref_loader: /root/reference is absent on the GPU box, so the checker that
travels with the repo must be ours; oracle/_ref (git-ignored staging) is used when present.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import thirdparty_ops as tp


# ---------------------------------------------------------------------------------------------
# parameter containers with the reference's state_dict key layout
# ---------------------------------------------------------------------------------------------
def _mlp(channels):
    """/root/reference/model/point_net2.py:45-53 -- per layer Linear -> ReLU -> BatchNorm1d."""
    blocks = []
    for cin, cout in zip(channels[:-1], channels[1:]):
        blocks.append(nn.Sequential(nn.Linear(cin, cout), nn.ReLU(), nn.BatchNorm1d(cout)))
    return nn.Sequential(*blocks)


class _Conv(nn.Module):
    def __init__(self, local_nn):
        super().__init__()
        self.local_nn = local_nn


class _SA(nn.Module):
    def __init__(self, ratio, r, mlp):
        super().__init__()
        self.ratio, self.r = ratio, r
        self.conv = _Conv(mlp)


class _NN(nn.Module):
    def __init__(self, mlp, k=None):
        super().__init__()
        self.k = k
        self.nn = mlp


class PointNet2Port(nn.Module):
    """Same parameters / buffers / key names as the reference PointNet2
    (/root/reference/model/point_net2.py:71-104), forward restated from :106-153."""

    def __init__(self, args):
        super().__init__()
        self.subsample_size = args.subsample_size
        self.drop = args.drop
        f = args.n_input_feats - 2  # x and y are dropped, :77 and :118
        c1 = [f + 3, 16, 16]
        c2 = [c1[-1] + 3, 32]
        c3 = [c2[-1] + 3, 64]
        self.sa1_module = _SA(args.ratio1, args.r1, _mlp(c1))
        self.sa2_module = _SA(args.ratio2, args.r2, _mlp(c2))
        self.sa3_module = _NN(_mlp(c3))
        c3f = [c3[-1] + c2[-1], 64]
        c2f = [c3f[-1] + c1[-1], 34]
        c1f = [c2f[-1] + f, 34]
        self.fp3_module = _NN(_mlp(c3f), k=1)
        self.fp2_module = _NN(_mlp(c2f), k=3)
        self.fp1_module = _NN(_mlp(c1f), k=3)
        self.lin1 = nn.Linear(c1f[-1], 16)
        self.lin2 = nn.Linear(16, args.n_class + 1)
        self.lin2.bias = nn.Parameter(torch.tensor([0.733, 0.266, 0.235, 0.358, 0.500]))  # :97-99
        self.trace = None  # filled by forward(..., trace=True) with the intermediates

    # -- set abstraction, :21-29 -------------------------------------------------------------
    def _sa(self, mod, x, pos, batch, K, trace, tag):
        idx = tp.fps(pos, batch, ratio=mod.ratio)
        row, col = tp.radius(pos, pos[idx], mod.r, batch, batch[idx], max_num_neighbors=K)
        # edge j -> i: source point col, target centroid row (:26)
        rel = pos[col] - pos[idx][row]
        msg = mod.conv.local_nn(torch.cat([x[col], rel], dim=1))  # A3: features first, then rel-pos
        out, arg = tp.scatter_max(msg, row, dim=0, dim_size=idx.numel())
        if trace is not None:
            trace[tag + "_idx"] = idx
            trace[tag + "_row"] = row
            trace[tag + "_col"] = col
            trace[tag + "_x"] = out
            trace[tag + "_arg"] = arg
        return out, pos[idx], batch[idx]

    # -- feature propagation, :62-67 ---------------------------------------------------------
    @staticmethod
    def _fp(mod, x, pos, batch, x_skip, pos_skip, batch_skip):
        y = tp.knn_interpolate(x, pos, pos_skip, batch, batch_skip, k=mod.k)
        if x_skip is not None:
            y = torch.cat([y, x_skip], dim=1)  # [interpolated, skip], :65
        return mod.nn(y)

    def forward(self, cloud_data, max_num_neighbors=2000, trace=False):
        tr = OrderedDict() if trace else None
        xyz, cloud = cloud_data["xyz"], cloud_data["cloud"]
        B, _, N = cloud.shape
        # :107-118 -- (B,f,N) -> (B*N,f), batch vector, drop the x,y feature columns
        pos0 = xyz.permute(0, 2, 1).reshape(B * N, 3)
        x0 = cloud.permute(0, 2, 1).reshape(B * N, -1)[:, 2:]
        batch0 = torch.arange(B, dtype=torch.int64).repeat_interleave(N)
        x1, pos1, batch1 = self._sa(self.sa1_module, x0, pos0, batch0, max_num_neighbors, tr, "sa1")
        x2, pos2, batch2 = self._sa(self.sa2_module, x1, pos1, batch1, max_num_neighbors, tr, "sa2")
        # global SA, :37-42
        g = tp.global_max_pool(self.sa3_module.nn(torch.cat([x2, pos2], dim=1)), batch2)
        posg = pos2.new_zeros((B, 3))
        batchg = torch.arange(B, dtype=torch.int64)
        f3 = self._fp(self.fp3_module, g, posg, batchg, x2, pos2, batch2)
        f2 = self._fp(self.fp2_module, f3, pos2, batch2, x1, pos1, batch1)
        f1 = self._fp(self.fp1_module, f2, pos1, batch1, x0, pos0, batch0)
        # head, :141-151
        h = F.relu(self.lin1(f1))
        h = F.dropout(h, p=self.drop, training=self.training)
        scores = self.lin2(h)
        proba = torch.softmax(scores[:, :4], dim=1)
        density = torch.sigmoid(scores[:, 4:5])
        cov = proba * density
        if tr is not None:
            tr.update(G=g, fp3=f3, fp2=f2, fp1=f1, scores=scores)
            self.trace = tr
        return cov, proba


# ---------------------------------------------------------------------------------------------
# projections
# ---------------------------------------------------------------------------------------------
def plotwise_pixel_ids(clouds: torch.Tensor, diam_pix: int) -> torch.Tensor:
    """/root/reference/model/project_to_2d.py:14-22 -- data-dependent min/max normalisation, fp32,
    op by op: floor((xy - min) / (max - min + 0.0001) * diam_pix).int().   -> int32 (B, 2, N)."""
    xy = clouds[:, :2, :].to(torch.float32)
    mn = xy.min(dim=2, keepdim=True).values
    mx = xy.max(dim=2, keepdim=True).values
    return torch.floor((xy - mn) / (mx - mn + 0.0001) * diam_pix).int()


def project_to_plotwise_coverages_port(pred_pointwise, clouds, args, return_aux=False):
    """/root/reference/model/project_to_2d.py:7-55.  pred (B*N,4), clouds (B,10,N) -> (B,4)
    [low, bare = 1 - low, medium, high]; mean over OCCUPIED pixels; differentiable."""
    B, _, N = clouds.shape
    pix = plotwise_pixel_ids(clouds, args.diam_pix)
    index, group, off = [], [], 0
    for b in range(B):
        key = pix[b, 0].to(torch.int64) * (4 * args.diam_pix + 4) + pix[b, 1].to(torch.int64)
        uniq, inv = torch.unique(key, return_inverse=True)  # sorted == lexicographic (x, y), :24
        index.append(inv + off)
        group.append(torch.full((uniq.numel(),), b, dtype=torch.int64))
        off += uniq.numel()
    index, group = torch.cat(index), torch.cat(group)
    pixel_max, arg = tp.scatter_max(pred_pointwise.transpose(1, 0), index)  # :39
    low, med, high = pixel_max[0], pixel_max[2], pixel_max[3]  # :41-44
    cols = [tp.scatter_mean(c, group, dim_size=B) for c in (low, 1 - low, med, high)]  # :46-49
    out = torch.stack(cols).T
    if return_aux:
        return out, dict(pix=pix, index=index, arg=arg)
    return out


def raster_pixel_ids(cloud: torch.Tensor, diam_pix: int, diam_meters: int) -> torch.Tensor:
    """/root/reference/model/project_to_2d.py:68-78 -- fixed affine map + clip, fp32, op by op:
    clip(floor((xy + 0.0001) * (10 * diam_pix / diam_meters) + diam_meters // 2).int(), 0, diam_pix-1)."""
    scaling = 10 * (diam_pix / diam_meters)
    xy = cloud[..., :2, :].to(torch.float32)
    return torch.clip(torch.floor((xy + 0.0001) * scaling + float(diam_meters // 2)).int(), 0, diam_pix - 1)


def project_to_2d_rasters_port(cloud, coverages_pointwise, args) -> np.ndarray:
    """/root/reference/model/project_to_2d.py:58-113.  cloud (F>=2, N), coverages (4, N) ->
    float64 (3, diam_pix, diam_pix): bands low/medium/high = channels 0/2/3, NaN where no point,
    image[y_pix, x_pix] then flipped along axis 0 (:103-110)."""
    D = args.diam_pix
    pix = raster_pixel_ids(cloud.detach().cpu(), D, args.diam_meters).numpy()
    vals = coverages_pointwise.detach().cpu().to(torch.float32).numpy()
    img = np.full((3, D, D), np.nan, dtype=np.float64)
    lin = pix[1].astype(np.int64) * D + pix[0].astype(np.int64)  # row = y, col = x
    for band, ch in enumerate((0, 2, 3)):
        flat = np.full(D * D, -np.inf, dtype=np.float32)
        np.maximum.at(flat, lin, vals[ch])
        occ = np.zeros(D * D, dtype=bool)
        occ[lin] = True
        img[band] = np.where(occ, flat.astype(np.float64), np.nan).reshape(D, D)
    return np.flip(img, axis=1).copy()


def m_of(n: int, ratio: float) -> int:
    """Sample count rule of A1: ceil(float32(n) * float32(ratio))."""
    return int(math.ceil(float(np.float32(n) * np.float32(ratio))))
