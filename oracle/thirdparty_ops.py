"""CPU oracle: restated third-party ops of the reference hot path.  TEST INFRASTRUCTURE ONLY.

The reference (/root/reference/model/point_net2.py:9, /root/reference/model/project_to_2d.py:4)
imports its arithmetic from wheels that are not in the tree and cannot be installed here
(torch-cluster==1.5.9, torch-geometric==1.7.2, torch-scatter==2.0.7 --
/root/reference/setup_environment/torch_extensions.txt:1-4).  This module restates their
published behaviour, following SURVEY.md Appendix A rule by rule (A1..A7), with the same public
signatures, so that the reference's own model files run on top of it unmodified
(see oracle/ref_loader.py) and so that the CUDA kernels have a checker.

PARITY WITH THE REAL WHEELS IS UNPINNED (no copy of them exists in this environment).  What pins
this oracle: tests/golden/* (reference files run verbatim on these ops, script committed) and the
known-answer micro-cases in tests/test_oracle_ops.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this package; the product (stratanet2-vegetation-coverage-maps_b200/) never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libsn2_oracle.so")
_lib = None


def build_oracle_lib(force: bool = False) -> str:
    """Compile oracle/csrc/sn2_oracle.c with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "csrc", "sn2_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])
    return _LIB_PATH


def _clib():
    global _lib
    if _lib is None:
        build_oracle_lib()
        lib = ctypes.CDLL(_LIB_PATH)
        vp, i64, f32, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_float, ctypes.c_int
        lib.o_fps.argtypes = [vp, vp, vp, vp, i64, vp]
        lib.o_fps.restype = None
        lib.o_radius.argtypes = [vp, vp, vp, vp, i64, f32, i64, i32, vp, vp, vp]
        lib.o_radius.restype = None
        lib.o_knn.argtypes = [vp, vp, vp, vp, i64, i64, vp, vp]
        lib.o_knn.restype = None
        _lib = lib
    return _lib


def _ptr_from_batch(batch: torch.Tensor | None, n: int, B: int | None = None) -> torch.Tensor:
    """Sorted ``batch`` vector -> CSR segment offsets (int64, length B+1)."""
    if batch is None:
        return torch.tensor([0, n], dtype=torch.int64)
    batch = batch.detach().cpu().to(torch.int64)
    if n == 0:
        return torch.zeros((B or 0) + 1, dtype=torch.int64)
    if bool((batch[1:] < batch[:-1]).any()):
        raise RuntimeError("batch vector must be sorted")
    nb = int(batch.max()) + 1 if B is None else B
    deg = torch.bincount(batch, minlength=nb)
    ptr = torch.zeros(nb + 1, dtype=torch.int64)
    ptr[1:] = torch.cumsum(deg, 0)
    return ptr


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().cpu().to(torch.float32).contiguous()


# --------------------------------------------------------------------------------------------
# A1  fps
# --------------------------------------------------------------------------------------------
def fps(x, batch=None, ratio=0.5, random_start=False, start=None):
    """torch_cluster.fps (called at /root/reference/model/point_net2.py:22).

    SURVEY.md A1.  m_b = ceil(float32(n_b) * float32(ratio)); squared fp32 distances; arg-max ties
    -> lowest index.  Upstream defaults to ``random_start=True`` (not reproducible); the canonical
    contract pins start = first point of each segment, so the default here is False.  ``start``
    (local index per segment) exposes an explicit start for tests.
    """
    if random_start:
        raise RuntimeError("oracle fps: random_start=True is not reproducible; pass start= instead")
    pos = _f32c(x)
    n = pos.shape[0]
    ptr = _ptr_from_batch(batch, n)
    B = ptr.numel() - 1
    deg = (ptr[1:] - ptr[:-1]).to(torch.float32)
    m = torch.ceil(deg * torch.tensor(ratio, dtype=torch.float32)).to(torch.int64)
    optr = torch.zeros(B + 1, dtype=torch.int64)
    optr[1:] = torch.cumsum(m, 0)
    out = torch.empty(int(optr[-1]), dtype=torch.int64)
    st = None
    if start is not None:
        st = torch.as_tensor(start, dtype=torch.int64).contiguous()
    _clib().o_fps(pos.data_ptr(), ptr.data_ptr(), optr.data_ptr(),
                  st.data_ptr() if st is not None else None, B, out.data_ptr())
    return out.to(x.device)


def fps_slow(pos: np.ndarray, m: int, start: int = 0) -> np.ndarray:
    """Independent numpy restatement of A1 for one segment (cross-checks the C code)."""
    pos = np.asarray(pos, dtype=np.float32)
    n = pos.shape[0]
    dist = np.full(n, np.inf, dtype=np.float32)
    out = np.empty(m, dtype=np.int64)
    last = start
    out[0] = last
    for it in range(1, m):
        d = pos - pos[last]
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        dist = np.minimum(dist, d2.astype(np.float32))
        last = int(np.argmax(dist))  # numpy argmax: first occurrence
        out[it] = last
    return out


# --------------------------------------------------------------------------------------------
# A2  radius
# --------------------------------------------------------------------------------------------
def radius_threshold(r: float) -> float:
    """``r*r`` evaluated in double then rounded to fp32 (SURVEY.md A2 / §7 hard part 2)."""
    return float(np.float32(float(r) * float(r)))


def radius(x, y, r, batch_x=None, batch_y=None, max_num_neighbors=32, num_workers=1):
    """torch_cluster.radius (called at /root/reference/model/point_net2.py:23-25).

    Returns int64 [2, E]: row 0 = query index into y, row 1 = point index into x; grouped by query
    (ascending), ascending point index inside a query, strict ``d2 < fp32(r*r)``, first K kept.
    """
    xs, ys = _f32c(x), _f32c(y)
    B = None
    if batch_x is not None and batch_y is not None and xs.shape[0] and ys.shape[0]:
        B = int(max(int(batch_x.max()), int(batch_y.max()))) + 1
    ptr_x = _ptr_from_batch(batch_x, xs.shape[0], B)
    ptr_y = _ptr_from_batch(batch_y, ys.shape[0], B)
    nb = ptr_x.numel() - 1
    ny = ys.shape[0]
    cnt = torch.zeros(ny, dtype=torch.int64)
    r2 = radius_threshold(r)
    lib = _clib()
    lib.o_radius(xs.data_ptr(), ptr_x.data_ptr(), ys.data_ptr(), ptr_y.data_ptr(), nb, r2,
                 int(max_num_neighbors), 0, cnt.data_ptr(), None, None)
    rowptr = torch.zeros(ny + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(cnt, 0)
    col = torch.empty(int(rowptr[-1]), dtype=torch.int64)
    lib.o_radius(xs.data_ptr(), ptr_x.data_ptr(), ys.data_ptr(), ptr_y.data_ptr(), nb, r2,
                 int(max_num_neighbors), 1, None, rowptr.data_ptr(), col.data_ptr())
    row = torch.repeat_interleave(torch.arange(ny, dtype=torch.int64), cnt)
    return torch.stack([row, col], dim=0).to(x.device)


def radius_slow(x: np.ndarray, y: np.ndarray, r: float, K: int):
    """Independent numpy restatement of A2 for one segment -> list of neighbour arrays."""
    x = np.asarray(x, np.float32)
    y = np.asarray(y, np.float32)
    r2 = np.float32(radius_threshold(r))
    out = []
    for q in range(y.shape[0]):
        d = x - y[q]
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        out.append(np.nonzero(d2 < r2)[0][:K].astype(np.int64))
    return out


# --------------------------------------------------------------------------------------------
# A5  knn / knn_interpolate
# --------------------------------------------------------------------------------------------
def knn_raw(x, y, k, batch_x=None, batch_y=None):
    """(idx [Ny,k] int64 with -1 padding, d2 [Ny,k] fp32): k nearest x of the same segment per y."""
    xs, ys = _f32c(x), _f32c(y)
    B = None
    if batch_x is not None and batch_y is not None and xs.shape[0] and ys.shape[0]:
        B = int(max(int(batch_x.max()), int(batch_y.max()))) + 1
    ptr_x = _ptr_from_batch(batch_x, xs.shape[0], B)
    ptr_y = _ptr_from_batch(batch_y, ys.shape[0], B)
    ny = ys.shape[0]
    idx = torch.empty((ny, k), dtype=torch.int64)
    d2 = torch.empty((ny, k), dtype=torch.float32)
    _clib().o_knn(xs.data_ptr(), ptr_x.data_ptr(), ys.data_ptr(), ptr_y.data_ptr(),
                  ptr_x.numel() - 1, k, idx.data_ptr(), d2.data_ptr())
    return idx, d2


def knn(x, y, k, batch_x=None, batch_y=None, cosine=False, num_workers=1):
    """torch_cluster.knn: int64 [2, <= k*len(y)], row 0 = y index, row 1 = x index."""
    assert not cosine
    idx, _ = knn_raw(x, y, k, batch_x, batch_y)
    ny = idx.shape[0]
    row = torch.arange(ny, dtype=torch.int64).view(-1, 1).expand(ny, k)
    keep = idx >= 0
    return torch.stack([row[keep], idx[keep]], dim=0).to(x.device)


def knn_interpolate(x, pos_x, pos_y, batch_x=None, batch_y=None, k=3, num_workers=1):
    """torch_geometric.nn.knn_interpolate (called at /root/reference/model/point_net2.py:63).

    SURVEY.md A5: weights 1 / clamp(d2, min=1e-16) computed under no_grad from
    diff = pos_x[src] - pos_y[dst]; y = sum(w * x[src]) / sum(w); gradient flows to x only.
    """
    with torch.no_grad():
        assign_index = knn(pos_x, pos_y, k, batch_x=batch_x, batch_y=batch_y)
        y_idx, x_idx = assign_index[0], assign_index[1]
        diff = pos_x[x_idx] - pos_y[y_idx]
        sq = diff * diff
        squared_distance = ((sq[:, 0] + sq[:, 1]) + sq[:, 2]).unsqueeze(-1)
        weights = 1.0 / torch.clamp(squared_distance, min=1e-16)
    ny = pos_y.size(0)
    num = torch.zeros((ny, x.size(1)), dtype=x.dtype, device=x.device).index_add_(0, y_idx, x[x_idx] * weights)
    den = torch.zeros((ny, 1), dtype=x.dtype, device=x.device).index_add_(0, y_idx, weights)
    return num / den


# --------------------------------------------------------------------------------------------
# A6 / A7  scatter_max, scatter_mean  (+ A4 global_max_pool, A3 PointConv on top)
# --------------------------------------------------------------------------------------------
class _ScatterMaxLast(torch.autograd.Function):
    """scatter max along the LAST dim with first-index arg-max and arg-routed backward (A6)."""

    @staticmethod
    def forward(ctx, src, index, dim_size):
        n = src.shape[-1]
        lead = src.shape[:-1]
        idx = index.view((1,) * len(lead) + (n,)).expand_as(src)
        out = torch.full(lead + (dim_size,), float("-inf"), dtype=src.dtype)
        out = out.scatter_reduce(-1, idx, src, reduce="amax", include_self=True)
        hit = src == out.gather(-1, idx)
        pos = torch.arange(n, dtype=torch.int64).view((1,) * len(lead) + (n,)).expand_as(src)
        cand = torch.where(hit, pos, torch.full_like(pos, n))
        arg = torch.full(lead + (dim_size,), n, dtype=torch.int64)
        arg = arg.scatter_reduce(-1, idx, cand, reduce="amin", include_self=True)
        empty = arg == n
        out = torch.where(empty, torch.zeros_like(out), out)  # torch_scatter: empty slot -> 0
        ctx.save_for_backward(arg)
        ctx.n = n
        ctx.mark_non_differentiable(arg)
        return out, arg

    @staticmethod
    def backward(ctx, gout, _garg):
        (arg,) = ctx.saved_tensors
        n = ctx.n
        gsrc = torch.zeros(arg.shape[:-1] + (n + 1,), dtype=gout.dtype)
        gsrc.scatter_(-1, arg, gout)
        return gsrc[..., :n], None, None


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    """torch_scatter.scatter_max (called at /root/reference/model/project_to_2d.py:39). -> (out, argmax)."""
    assert out is None
    dim = dim % src.dim()
    if index.dim() != 1:
        # broadcast-style index as torch_scatter accepts: reduce to its 1-D generator along dim
        sl = [0] * index.dim()
        sl[dim] = slice(None)
        index = index[tuple(sl)]
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    moved = src.movedim(dim, -1)
    o, a = _ScatterMaxLast.apply(moved, index.to(torch.int64), dim_size)
    return o.movedim(-1, dim), a.movedim(-1, dim)


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    """torch_scatter.scatter_mean (called at /root/reference/model/project_to_2d.py:46-49)."""
    assert out is None
    dim = dim % src.dim()
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    moved = src.movedim(dim, 0)
    tot = torch.zeros((dim_size,) + moved.shape[1:], dtype=src.dtype).index_add_(0, index, moved)
    cnt = torch.zeros(dim_size, dtype=src.dtype).index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
    cnt = cnt.clamp(min=1).view((dim_size,) + (1,) * (moved.dim() - 1))
    return (tot / cnt).movedim(0, dim)


def scatter_add(src, index, dim=-1, out=None, dim_size=None):
    assert out is None
    dim = dim % src.dim()
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    moved = src.movedim(dim, 0)
    tot = torch.zeros((dim_size,) + moved.shape[1:], dtype=src.dtype).index_add_(0, index, moved)
    return tot.movedim(0, dim)


def global_max_pool(x, batch, size=None):
    """torch_geometric.nn.global_max_pool (called at /root/reference/model/point_net2.py:39). A4."""
    size = int(batch.max()) + 1 if size is None else size
    return scatter_max(x, batch, dim=0, dim_size=size)[0]


class PointConv(torch.nn.Module):
    """torch_geometric.nn.PointConv restated (constructed at /root/reference/model/point_net2.py:19,
    called at :27).  SURVEY.md A3: msg = local_nn(cat[x_j, pos_j - pos_i]); max aggregation over the
    edges into each target; arg-max ties -> first edge; no self loops added here (the reference
    passes add_self_loops=False); ``global_nn`` unused by the reference."""

    def __init__(self, local_nn=None, global_nn=None, add_self_loops=True, **kwargs):
        super().__init__()
        self.local_nn = local_nn
        self.global_nn = global_nn
        self.add_self_loops = add_self_loops
        if add_self_loops:
            raise NotImplementedError("oracle PointConv: the reference path uses add_self_loops=False")

    def forward(self, x, pos, edge_index):
        if not isinstance(x, tuple):
            x = (x, None)
        if isinstance(pos, torch.Tensor):
            pos = (pos, pos)
        src, dst = edge_index[0], edge_index[1]
        msg = pos[0][src] - pos[1][dst]
        if x[0] is not None:
            msg = torch.cat([x[0][src], msg], dim=1)
        if self.local_nn is not None:
            msg = self.local_nn(msg)
        out = scatter_max(msg, dst, dim=0, dim_size=pos[1].size(0))[0]
        if self.global_nn is not None:
            out = self.global_nn(out)
        return out
