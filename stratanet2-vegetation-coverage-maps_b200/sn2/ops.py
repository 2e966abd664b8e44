"""Host-side wrappers over the C ABI (include/sn2.h).

Two levels:
  * dense primitives (``ingest``, ``fps_dense``, ``ball_query_dense`` ...) used by the fused forward
    of ``model.point_net2.PointNet2``: B plots of exactly N points, int32 indices, pos4 positions;
  * operator-level functions with the upstream names and argument meaning that the reference calls
    (``fps``, ``radius``, ``knn``, ``global_max_pool``, ``scatter_max`` ... --
    /root/reference/model/point_net2.py:9, /root/reference/model/project_to_2d.py:4), returning
    int64 indices in the upstream layout, so parity tests read like calls into the original wheels.

torch is used for device memory and streams only; every computation is a kernel of libsn2_b200.so.
No CPU path exists: CPU tensors raise RuntimeError.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib
from ._lib import check, dptr, hptr, stream_ptr

LAUNCHES = 0  # kernels of libsn2_b200.so enqueued by this process (bench.py reports the delta)


def _count(n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n


GRID_CELLS = 64 * 64
GRID_HDR = 12
CF_LD = 36


def m_of(n: int, ratio: float) -> int:
    """Sample count of torch_cluster.fps: ceil(float32(n) * float32(ratio))."""
    return int(math.ceil(float(np.float32(n) * np.float32(ratio))))


def r2_of(r: float) -> float:
    """Threshold of torch_cluster.radius: r*r in double, rounded to fp32."""
    return float(np.float32(float(r) * float(r)))


def _f32(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("sn2: CUDA tensors required (this build has no CPU fallback)")
    return t.detach().to(torch.float32).contiguous()


# ------------------------------------------------------------------------------------------------
# dense primitives
# ------------------------------------------------------------------------------------------------
def ingest(xyz: torch.Tensor, cloud: torch.Tensor):
    """(B,3,N) + (B,10,N) device fp32 -> pos4 (B*N,4), feat (B*N,8)."""
    lib = _lib.load()
    B, F, N = cloud.shape
    pos4 = torch.empty((B * N, 4), dtype=torch.float32, device=cloud.device)
    feat = torch.empty((B * N, F - 2), dtype=torch.float32, device=cloud.device)
    check(lib.sn2_ingest(dptr(xyz, torch.float32), dptr(cloud, torch.float32), B, N, F, dptr(pos4), dptr(feat),
                         stream_ptr()), "sn2_ingest")
    _count(1)
    return pos4, feat


def ingest_pos(xyz: torch.Tensor):
    """(B,3,N) device fp32 -> pos4 (B*N,4): the half of `ingest` that FPS / the grids need."""
    lib = _lib.load()
    B, _, N = xyz.shape
    pos4 = torch.empty((B * N, 4), dtype=torch.float32, device=xyz.device)
    check(lib.sn2_ingest(dptr(xyz, torch.float32), None, B, N, 0, dptr(pos4), None, stream_ptr()), "sn2_ingest")
    _count(1)
    return pos4


def ingest_feat(cloud: torch.Tensor):
    """(B,10,N) device fp32 -> feat (B*N,8) (x, y dropped)."""
    lib = _lib.load()
    B, F, N = cloud.shape
    feat = torch.empty((B * N, F - 2), dtype=torch.float32, device=cloud.device)
    check(lib.sn2_ingest(None, dptr(cloud, torch.float32), B, N, F, None, dptr(feat), stream_ptr()), "sn2_ingest")
    _count(1)
    return feat


FPS_AUTO, FPS_BRUTE, FPS_BUCKETED, FPS_BUCKETED_SPEC4, FPS_CLUSTER4 = 0, 1, 2, 3, 4
FPS_BUCKETED_ILP2, FPS_BUCKETED_NW16, FPS_BUCKETED_NW16_ILP2 = 5, 6, 7


def fps_dense(pos4: torch.Tensor, B: int, N: int, M: int, start: torch.Tensor | None = None, algo: int = FPS_AUTO):
    """-> idx int32 [B*M] (global rows), pos4_out [B*M,4]."""
    lib = _lib.load()
    idx = torch.empty(B * M, dtype=torch.int32, device=pos4.device)
    out = torch.empty((B * M, 4), dtype=torch.float32, device=pos4.device)
    check(lib.sn2_fps_algo(dptr(pos4, torch.float32), B, N, M, dptr(start, torch.int32) if start is not None else None,
                           dptr(idx), dptr(out), int(algo), stream_ptr()), "sn2_fps")
    _count(1)
    return idx, out


_PINNED_RING = None
_PINNED_NEXT = 0


def _pinned_slot() -> torch.Tensor:
    """One int32 of page-locked host memory from a ring allocated once (a fresh pinned allocation per call can
    cost a cudaHostAlloc, which synchronises the device)."""
    global _PINNED_RING, _PINNED_NEXT
    if _PINNED_RING is None:
        _PINNED_RING = torch.empty(1024, dtype=torch.int32, pin_memory=True)
    _PINNED_NEXT = (_PINNED_NEXT + 1) % 1024
    return _PINNED_RING[_PINNED_NEXT:_PINNED_NEXT + 1]


def ball_query_begin(pos4: torch.Tensor, qpos4: torch.Tensor, B: int, N: int, M: int, r: float, K: int) -> dict:
    """First half of the CSR ball query: grid, per-query counts, row pointers, and an asynchronous copy of the edge
    total to pinned host memory.  Nothing here waits for the GPU; `ball_query_finish` does (on the recorded event)."""
    lib = _lib.load()
    dev = pos4.device
    hdr = torch.empty(B * GRID_HDR, dtype=torch.float32, device=dev)
    cell_start = torch.empty(B * (GRID_CELLS + 1), dtype=torch.int32, device=dev)
    sorted4 = torch.empty((B * N, 4), dtype=torch.float32, device=dev)
    st = stream_ptr()
    check(lib.sn2_grid_build(dptr(pos4, torch.float32), B, N, float(r), dptr(hdr), dptr(cell_start), dptr(sorted4), st),
          "sn2_grid_build")
    _count(1)
    r2 = r2_of(r)
    cnt = torch.empty(B * M, dtype=torch.int32, device=dev)
    check(lib.sn2_ball_count(dptr(hdr), dptr(cell_start), dptr(sorted4), dptr(qpos4, torch.float32), B, N, M, r2, int(K),
                             dptr(cnt), st), "sn2_ball_count")
    _count(1)
    rowptr = torch.empty(B * M + 1, dtype=torch.int32, device=dev)
    scratch = torch.empty(B, dtype=torch.int32, device=dev)
    check(lib.sn2_rowptr_scan(dptr(cnt), B, M, dptr(rowptr), dptr(scratch), st), "sn2_rowptr_scan")
    _count(2)
    total = _pinned_slot()
    total.copy_(rowptr[-1:], non_blocking=True)
    ready = torch.cuda.Event()
    ready.record()
    return dict(hdr=hdr, cell_start=cell_start, sorted4=sorted4, qpos4=qpos4, rowptr=rowptr, total=total, ready=ready,
                shape=(B, N, M, r2, int(K)))


def ball_query_finish(state: dict):
    """Second half: wait for the edge total (the one host sync of the level, sizes the edge list) and fill `col`.
    Runs on the current stream, which must be the stream of `ball_query_begin` or ordered after it."""
    lib = _lib.load()
    state["ready"].synchronize()
    E = int(state["total"][0])
    B, N, M, r2, K = state["shape"]
    rowptr = state["rowptr"]
    col = torch.empty(max(E, 1), dtype=torch.int32, device=rowptr.device)
    check(lib.sn2_ball_fill(dptr(state["hdr"]), dptr(state["cell_start"]), dptr(state["sorted4"]), dptr(state["qpos4"]), B, N, M,
                            r2, K, dptr(rowptr), dptr(col), stream_ptr()), "sn2_ball_fill")
    _count(1)
    return rowptr, col[:E]


def ball_query_dense(pos4: torch.Tensor, qpos4: torch.Tensor, B: int, N: int, M: int, r: float, K: int):
    """-> rowptr int32 [B*M+1], col int32 [E] (global point rows, ascending inside each query)."""
    return ball_query_finish(ball_query_begin(pos4, qpos4, B, N, M, r, K))


def pointconv_fwd(level: int, pos4, feat, qpos4, rowptr, col, w_host: torch.Tensor):
    lib = _lib.load()
    Q = qpos4.shape[0]
    cout = 16 if level == 1 else 32
    out = torch.empty((Q, cout), dtype=torch.float32, device=pos4.device)
    check(lib.sn2_pointconv_fwd(level, dptr(pos4, torch.float32), dptr(feat, torch.float32), dptr(qpos4, torch.float32),
                                dptr(rowptr, torch.int32), dptr(col, torch.int32), Q, hptr(w_host), w_host.numel(),
                                dptr(out), stream_ptr()), "sn2_pointconv_fwd")
    _count(1)
    return out


def build_grid(pos4, B: int, N: int, r: float):
    """xy cell grid of B plots of N points -> (hdr, cell_start, sorted4)."""
    lib = _lib.load()
    dev = pos4.device
    hdr = torch.empty(B * GRID_HDR, dtype=torch.float32, device=dev)
    cell_start = torch.empty(B * (GRID_CELLS + 1), dtype=torch.int32, device=dev)
    sorted4 = torch.empty((B * N, 4), dtype=torch.float32, device=dev)
    check(lib.sn2_grid_build(dptr(pos4, torch.float32), B, N, float(r), dptr(hdr), dptr(cell_start), dptr(sorted4),
                             stream_ptr()), "sn2_grid_build")
    _count(1)
    return hdr, cell_start, sorted4


def sa_fused_fwd(level: int, pos4, feat, qpos4, B: int, N: int, M: int, r: float, K: int, w_host, want_counts=False,
                 tensor_core: int = 0, grid=None):
    """Fused ball query + PointConv (eval).  -> out [B*M, 16|32] (+ neighbour counts int32 [B*M]).
    grid: optional prebuilt (hdr, cell_start, sorted4) = build_grid(pos4, B, N, r)."""
    lib = _lib.load()
    dev = pos4.device
    hdr, cell_start, sorted4 = grid if grid is not None else build_grid(pos4, B, N, r)
    _, _, qsorted4 = build_grid(qpos4, B, M, r)  # only used as a cell-ordered permutation of the queries
    cout = 16 if level == 1 else 32
    out = torch.empty((B * M, cout), dtype=torch.float32, device=dev)
    u = torch.empty((B * N, cout), dtype=torch.float32, device=dev)
    cnt = torch.empty(B * M, dtype=torch.int32, device=dev) if want_counts else None
    ovf = torch.empty(B * M + 1, dtype=torch.int32, device=dev)
    check(lib.sn2_sa_fused_fwd(level, dptr(hdr), dptr(cell_start), dptr(sorted4), dptr(qsorted4), dptr(pos4, torch.float32),
                               dptr(feat, torch.float32), dptr(u), dptr(ovf), B, N, M, r2_of(r), int(K), hptr(w_host),
                               w_host.numel(), dptr(out), dptr(cnt), int(tensor_core) if level == 1 else 0, stream_ptr()),
          "sn2_sa_fused_fwd")
    _count(4 if K < N else 3)
    return (out, cnt) if want_counts else out


def global_sa_fwd(x2, pos4, B: int, M: int, w_host):
    lib = _lib.load()
    g = torch.empty((B, 64), dtype=torch.float32, device=x2.device)
    check(lib.sn2_global_sa_fwd(dptr(x2, torch.float32), dptr(pos4, torch.float32), B, M, hptr(w_host), w_host.numel(),
                                dptr(g), stream_ptr()), "sn2_global_sa_fwd")
    _count(1)
    return g


def fp3_fwd(g, x2, pos4, B: int, M: int, w_host):
    lib = _lib.load()
    out = torch.empty((B * M, 64), dtype=torch.float32, device=x2.device)
    check(lib.sn2_fp3_fwd(dptr(g, torch.float32), dptr(x2, torch.float32), dptr(pos4, torch.float32), B, M, hptr(w_host),
                          w_host.numel(), dptr(out), stream_ptr()), "sn2_fp3_fwd")
    _count(1)
    return out


KNN_AUTO, KNN_BRUTE, KNN_GRID = 0, 1, 2


def knn3_dense(spos4, qpos4, B: int, Ms: int, Nq: int, algo: int = KNN_AUTO, qsorted4=None):
    """-> nbr int32 [B*Nq,3] (global source rows, ascending distance), w fp32 [B*Nq,3].
    KNN_GRID bins the sources (one extra kernel) and ring-searches; KNN_BRUTE scans all sources.
    qsorted4: optional cell-ordered copy of the queries (sorted4 of build_grid(qpos4, ...)) for coherent warps."""
    lib = _lib.load()
    dev = spos4.device
    nbr = torch.empty((B * Nq, 3), dtype=torch.int32, device=dev)
    w = torch.empty((B * Nq, 3), dtype=torch.float32, device=dev)
    if algo == KNN_AUTO:
        algo = KNN_GRID if Ms >= 256 else KNN_BRUTE
    if algo == KNN_BRUTE:
        check(lib.sn2_knn3(dptr(spos4, torch.float32), dptr(qpos4, torch.float32), B, Ms, Nq, dptr(nbr), dptr(w),
                           stream_ptr()), "sn2_knn3")
        _count(1)
        return nbr, w
    hdr = torch.empty(B * GRID_HDR, dtype=torch.float32, device=dev)
    cell_start = torch.empty(B * (GRID_CELLS + 1), dtype=torch.int32, device=dev)
    sorted4 = torch.empty((B * Ms, 4), dtype=torch.float32, device=dev)
    st = stream_ptr()
    check(lib.sn2_grid_build(dptr(spos4, torch.float32), B, Ms, -3.0, dptr(hdr), dptr(cell_start), dptr(sorted4), st),
          "sn2_grid_build")
    _count(1)
    qsrc = qsorted4 if qsorted4 is not None else qpos4
    check(lib.sn2_knn3_grid(dptr(hdr), dptr(cell_start), dptr(sorted4), dptr(qsrc, torch.float32), B, Ms, Nq, dptr(nbr),
                            dptr(w), int(qsorted4 is not None), st), "sn2_knn3_grid")
    _count(1)
    return nbr, w


def fp2_fwd(f3, nbr, w, x1, w_host):
    lib = _lib.load()
    Q = x1.shape[0]
    out = torch.empty((Q, CF_LD), dtype=torch.float32, device=x1.device)
    check(lib.sn2_fp2_fwd(dptr(f3, torch.float32), dptr(nbr, torch.int32), dptr(w, torch.float32), dptr(x1, torch.float32), Q,
                          hptr(w_host), w_host.numel(), dptr(out), stream_ptr()), "sn2_fp2_fwd")
    _count(1)
    return out


def fp1_head_fwd(f2, nbr, w, feat, w_host, tensor_core: bool = False):
    lib = _lib.load()
    fn, name = (lib.sn2_fp1_head_fwd_tc, "sn2_fp1_head_fwd_tc") if tensor_core else (lib.sn2_fp1_head_fwd, "sn2_fp1_head_fwd")
    Q = feat.shape[0]
    cov = torch.empty((Q, 4), dtype=torch.float32, device=feat.device)
    proba = torch.empty((Q, 4), dtype=torch.float32, device=feat.device)
    check(fn(dptr(f2, torch.float32), dptr(nbr, torch.int32), dptr(w, torch.float32), dptr(feat, torch.float32), Q,
             hptr(w_host), w_host.numel(), dptr(cov), dptr(proba), stream_ptr()), name)
    _count(1)
    return cov, proba


def project_plotwise(cloud_dev, pred, D: int, want_aux: bool = False):
    """cloud (B,F,N) device, pred (B*N,4) -> out (B,4) [+ pix (B*N), pmax, parg (B,3,D+1,D+1),
    all three in the (D+1) x (D+1) frame px*(D+1)+py of the reference's unclamped pixel ids]."""
    lib = _lib.load()
    B, F, N = cloud_dev.shape
    dev = cloud_dev.device
    out = torch.empty((B, 4), dtype=torch.float32, device=dev)
    pix = pmax = parg = None
    if want_aux:
        pix = torch.empty(B * N, dtype=torch.int32, device=dev)
        pmax = torch.empty((B, 3, D + 1, D + 1), dtype=torch.float32, device=dev)
        parg = torch.empty((B, 3, D + 1, D + 1), dtype=torch.int32, device=dev)
    check(lib.sn2_project_plotwise(dptr(cloud_dev, torch.float32), dptr(pred, torch.float32), B, N, F, D, dptr(out), dptr(pix),
                                   dptr(pmax), dptr(parg), stream_ptr()), "sn2_project_plotwise")
    _count(1)
    return (out, pix, pmax, parg) if want_aux else out


def project_rasters(cloud_dev, cov, layout: str, D: int, diam_meters: int, want_pix: bool = False):
    """cloud (B,F,N) device; cov either "point_major" (B*N,4) or "channel_major" (B,4,N).
    -> rasters float64 (B,3,D,D) [+ pix int32 (B*N) = y*D + x before the flip]."""
    lib = _lib.load()
    B, F, N = cloud_dev.shape
    if layout == "point_major":
        sb, sn, sc = N * 4, 4, 1
    elif layout == "channel_major":
        sb, sn, sc = 4 * N, 1, N
    else:
        raise ValueError(layout)
    dev = cloud_dev.device
    rasters = torch.empty((B, 3, D, D), dtype=torch.float64, device=dev)
    pix = torch.empty(B * N, dtype=torch.int32, device=dev) if want_pix else None
    scale = 10 * (D / diam_meters)          # /root/reference/model/project_to_2d.py:68
    shift = float(diam_meters // 2)         # :74
    check(lib.sn2_project_rasters(dptr(cloud_dev, torch.float32), dptr(cov, torch.float32), sb, sn, sc, B, N, F, D,
                                  float(scale), shift, dptr(rasters), dptr(pix), stream_ptr()), "sn2_project_rasters")
    _count(1)
    return (rasters, pix) if want_pix else rasters


# ------------------------------------------------------------------------------------------------
# operator-level API (upstream names / argument meaning; dense batches only)
# ------------------------------------------------------------------------------------------------
def _dense_shape(batch: torch.Tensor | None, n: int) -> tuple[int, int]:
    """Validate that ``batch`` is B equal, sorted, contiguous segments and return (B, n_per_plot).
    The hot path is dense by contract (/root/reference/model/point_net2.py:112-116)."""
    if batch is None:
        return 1, n
    if batch.numel() != n:
        raise RuntimeError("sn2: batch vector length does not match the number of points")
    B = int(batch[-1].item()) + 1 if n else 0
    if B == 0 or n % B:
        raise RuntimeError("sn2: ragged batches are not supported (plots must have equal point counts)")
    per = n // B
    expect = torch.arange(B, device=batch.device, dtype=batch.dtype).repeat_interleave(per)
    if not torch.equal(batch, expect):
        raise RuntimeError("sn2: batch vector must be sorted with equal-sized plots")
    return B, per


def to_pos4(pos: torch.Tensor) -> torch.Tensor:
    pos = _f32(pos)
    if pos.dim() != 2 or pos.shape[1] != 3:
        raise RuntimeError("sn2: positions must have shape [n, 3]")
    return torch.nn.functional.pad(pos, (0, 1)).contiguous()


def fps(x, batch=None, ratio=0.5, random_start=False, start=None, algo: int = FPS_AUTO):
    """Drop-in for torch_cluster/torch_geometric ``fps`` (/root/reference/model/point_net2.py:22).
    Canonical start = first point of each plot; ``start`` (local index per plot) overrides it;
    ``random_start=True`` draws it with torch's generator."""
    B, n = _dense_shape(batch, x.shape[0])
    M = m_of(n, ratio)
    st = None
    if start is not None:
        st = torch.as_tensor(start, dtype=torch.int32, device=x.device).contiguous()
    elif random_start:
        st = torch.randint(0, n, (B,), device=x.device, dtype=torch.int32)
    idx, _ = fps_dense(to_pos4(x), B, n, M, st, algo)
    return idx.to(torch.int64)


def radius(x, y, r, batch_x=None, batch_y=None, max_num_neighbors=32, num_workers=1):
    """Drop-in for ``radius`` (/root/reference/model/point_net2.py:23-25): int64 [2,E], row 0 = index into
    y, row 1 = index into x, grouped by query, ascending x index, strict d2 < fp32(r*r), first K."""
    B, n = _dense_shape(batch_x, x.shape[0])
    By, m = _dense_shape(batch_y, y.shape[0])
    if B != By:
        raise RuntimeError("sn2: batch_x and batch_y describe different numbers of plots")
    rowptr, col = ball_query_dense(to_pos4(x), to_pos4(y), B, n, m, float(r), int(max_num_neighbors))
    cnt = (rowptr[1:] - rowptr[:-1]).to(torch.int64)
    row = torch.repeat_interleave(torch.arange(B * m, device=x.device, dtype=torch.int64), cnt)
    return torch.stack([row, col.to(torch.int64)], dim=0)


def knn(x, y, k, batch_x=None, batch_y=None, cosine=False, num_workers=1):
    """Drop-in for ``knn`` with k == 3 (the only k the path searches; k=1 against a single source is a
    broadcast, see fp3): int64 [2, 3*len(y)], row 0 = y index, row 1 = x index, ascending distance."""
    if k != 3 or cosine:
        raise RuntimeError("sn2: knn supports k=3, euclidean")
    B, ms = _dense_shape(batch_x, x.shape[0])
    By, nq = _dense_shape(batch_y, y.shape[0])
    if B != By:
        raise RuntimeError("sn2: batch_x and batch_y describe different numbers of plots")
    nbr, _ = knn3_dense(to_pos4(x), to_pos4(y), B, ms, nq)
    row = torch.arange(B * nq, device=x.device, dtype=torch.int64).repeat_interleave(3)
    return torch.stack([row, nbr.reshape(-1).to(torch.int64)], dim=0)


def knn_interpolate(x, pos_x, pos_y, batch_x=None, batch_y=None, k=3, num_workers=1):
    """Drop-in for torch_geometric ``knn_interpolate`` (/root/reference/model/point_net2.py:63), differentiable
    w.r.t. ``x``.  k = 3 (general sources) or k = 1 with a single source per plot (the fp3 broadcast)."""
    from .autograd_ops import Interp3, InterpPlot

    B, ms = _dense_shape(batch_x, pos_x.shape[0])
    By, nq = _dense_shape(batch_y, pos_y.shape[0])
    if B != By:
        raise RuntimeError("sn2: batch_x and batch_y describe different numbers of plots")
    if k == 1 and ms == 1:
        if bool((pos_x != 0).any()):
            raise RuntimeError("sn2: k=1 interpolation is implemented for a single source at the origin (GlobalSAModule output)")
        return InterpPlot.apply(x, to_pos4(pos_y), nq)
    if k != 3:
        raise RuntimeError("sn2: knn_interpolate supports k=3, or k=1 from one source per plot")
    with torch.no_grad():
        nbr, w = knn3_dense(to_pos4(pos_x), to_pos4(pos_y), B, ms, nq)
    return Interp3.apply(x, nbr, w)


def global_max_pool(x, batch, size=None):
    """Drop-in for torch_geometric ``global_max_pool`` (/root/reference/model/point_net2.py:39), differentiable."""
    from .autograd_ops import SegmentMax

    B, m = _dense_shape(batch, x.shape[0])
    ptr = torch.arange(B + 1, dtype=torch.int32, device=x.device) * m
    return SegmentMax.apply(x, ptr)[0]


def pointconv(local_nn, x, pos, edge_index):
    """torch_geometric ``PointConv(local_nn, add_self_loops=False)(x, (pos_src, pos_dst), edge_index)`` with
    ``edge_index = [source point ; target centroid]`` grouped by target as ``radius`` returns it
    (/root/reference/model/point_net2.py:26-27).  Differentiable w.r.t. x and the parameters of local_nn."""
    from .autograd_ops import EdgeMsg, SegmentMax

    pos_src, pos_dst = pos if isinstance(pos, (tuple, list)) else (pos, pos)
    src, dst = edge_index[0], edge_index[1]
    if dst.numel() > 1 and bool((dst[1:] < dst[:-1]).any()):
        raise RuntimeError("sn2 PointConv: edges must be grouped by target (the layout radius() returns)")
    Q = pos_dst.shape[0]
    cnt = torch.bincount(dst, minlength=Q)
    rowptr = torch.zeros(Q + 1, dtype=torch.int32, device=dst.device)
    rowptr[1:] = torch.cumsum(cnt, 0).to(torch.int32)
    msg = EdgeMsg.apply(x, to_pos4(pos_src), to_pos4(pos_dst), rowptr, src.to(torch.int32).contiguous())
    return SegmentMax.apply(local_nn(msg), rowptr)[0]
