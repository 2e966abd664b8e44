"""torch.autograd.Function wrappers of the training-path operators (csrc/train_ops.cu).

Each Function calls one forward and one backward kernel of libsn2_b200.so; torch only carries the
autograd graph and owns the memory.  Index tensors (neighbour lists, arg-max, kNN) are int32 and carry no
gradient; positions are inputs and get none either (as in the reference, where they are leaf data).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib, ops
from . import comm as _comm
from ._lib import check, dptr, stream_ptr


def _c(t: torch.Tensor) -> torch.Tensor:
    return t.contiguous()


class EdgeMsg(torch.autograd.Function):
    """msg[e] = [x[col[e]], pos[col[e]] - qpos[row(e)]]  (PointConv.message)."""

    @staticmethod
    def forward(ctx, x, pos4, qpos4, rowptr, col, rows=None):
        """rows: optional device int32 [1] = live edge count when `col` is a fixed-capacity buffer (graph replay)."""
        lib = _lib.load()
        x = _c(x)
        E, C = col.numel(), x.shape[1]
        msg = torch.empty((E, C + 3), dtype=torch.float32, device=x.device)
        check(lib.sn2_edge_msg_fwd(dptr(x, torch.float32), dptr(pos4), dptr(qpos4), dptr(rowptr, torch.int32),
                                   dptr(col, torch.int32), qpos4.shape[0], C, dptr(msg), stream_ptr()), "sn2_edge_msg_fwd")
        ops._count(1)
        ctx.save_for_backward(col)
        ctx.rows = rows
        ctx.shape = (x.shape[0], C)
        return msg

    @staticmethod
    def backward(ctx, dmsg):
        if not ctx.needs_input_grad[0]:
            return None, None, None, None, None, None
        lib = _lib.load()
        (col,) = ctx.saved_tensors
        P, C = ctx.shape
        dx = torch.zeros((P, C), dtype=torch.float32, device=dmsg.device)
        check(lib.sn2_edge_msg_bwd(dptr(_c(dmsg), torch.float32), dptr(col), col.numel(), dptr(ctx.rows), C, dptr(dx),
                                   stream_ptr()), "sn2_edge_msg_bwd")
        ops._count(1)
        return dx, None, None, None, None, None


class SegmentMax(torch.autograd.Function):
    """Per-row max over CSR rows with first-edge arg-max; gradient routed to the arg-max edges (A3, A4, A6)."""

    @staticmethod
    def forward(ctx, vals, rowptr, ss=None):
        """ss: optional [>= 2C] scale | shift of the producing LinReluBN block (apply=False): the max is taken over
        vals * scale + shift and the gradient returned for `vals` is the gradient with respect to that product."""
        lib = _lib.load()
        vals = _c(vals)
        if vals.data_ptr() % 16:  # the kernel reads float4 channel quads
            vals = vals.clone()
        Q, C = rowptr.numel() - 1, vals.shape[1]
        out = torch.empty((Q, C), dtype=torch.float32, device=vals.device)
        arg = torch.empty((Q, C), dtype=torch.int32, device=vals.device)
        check(lib.sn2_segment_max_fwd(dptr(vals, torch.float32), dptr(ss, torch.float32), dptr(rowptr, torch.int32), Q, C,
                                      dptr(out), dptr(arg), stream_ptr()), "sn2_segment_max_fwd")
        ops._count(1)
        ctx.save_for_backward(arg)
        ctx.E = vals.shape[0]
        ctx.mark_non_differentiable(arg)
        return out, arg

    @staticmethod
    def backward(ctx, dout, _darg):
        lib = _lib.load()
        (arg,) = ctx.saved_tensors
        Q, C = arg.shape
        dvals = torch.zeros((ctx.E, C), dtype=torch.float32, device=dout.device)
        check(lib.sn2_segment_max_bwd(dptr(_c(dout), torch.float32), dptr(arg), Q, C, ctx.E, dptr(dvals), stream_ptr()),
              "sn2_segment_max_bwd")
        ops._count(1)
        return dvals, None, None


class Interp3(torch.autograd.Function):
    """knn_interpolate with k = 3 on precomputed neighbours / weights (A5); gradient to the features only."""

    @staticmethod
    def forward(ctx, x, nbr, w):
        lib = _lib.load()
        x = _c(x)
        Q, C = nbr.shape[0], x.shape[1]
        y = torch.empty((Q, C), dtype=torch.float32, device=x.device)
        check(lib.sn2_interp3_fwd(dptr(x, torch.float32), C, dptr(nbr, torch.int32), dptr(w, torch.float32), Q, C, dptr(y),
                                  stream_ptr()), "sn2_interp3_fwd")
        ops._count(1)
        ctx.save_for_backward(nbr, w)
        ctx.S = x.shape[0]
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        nbr, w = ctx.saved_tensors
        Q, C = dy.shape
        dx = torch.zeros((ctx.S, C), dtype=torch.float32, device=dy.device)
        check(lib.sn2_interp3_bwd(dptr(_c(dy), torch.float32), dptr(nbr), dptr(w), Q, C, dptr(dx), stream_ptr()),
              "sn2_interp3_bwd")
        ops._count(1)
        return dx, None, None


class InterpPlot(torch.autograd.Function):
    """knn_interpolate with k = 1 from the single per-plot vector at the origin (fp3): y = (g * w) / w."""

    @staticmethod
    def forward(ctx, g, pos4, M):
        lib = _lib.load()
        g = _c(g)
        B, C = g.shape
        y = torch.empty((B * M, C), dtype=torch.float32, device=g.device)
        check(lib.sn2_interp_plot_fwd(dptr(g, torch.float32), dptr(pos4), B, M, C, dptr(y), stream_ptr()), "sn2_interp_plot_fwd")
        ops._count(1)
        ctx.save_for_backward(pos4)
        ctx.dims = (B, M, C)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        (pos4,) = ctx.saved_tensors
        B, M, C = ctx.dims
        dg = torch.empty((B, C), dtype=torch.float32, device=dy.device)
        check(lib.sn2_interp_plot_bwd(dptr(_c(dy), torch.float32), dptr(pos4), B, M, C, dptr(dg), stream_ptr()),
              "sn2_interp_plot_bwd")
        ops._count(1)
        return dg, None, None


class ProjectPlotwise(torch.autograd.Function):
    """project_to_plotwise_coverages with its arg-routed backward (A6 + A7)."""

    @staticmethod
    def forward(ctx, pred, cloud_dev, D):
        out, _pix, _pmax, parg = ops.project_plotwise(cloud_dev, _c(pred.detach()), D, want_aux=True)
        ctx.save_for_backward(parg)
        ctx.n = pred.shape[0]
        ctx.D = D
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        (parg,) = ctx.saved_tensors
        dpred = torch.zeros((ctx.n, 4), dtype=torch.float32, device=dout.device)
        check(lib.sn2_project_plotwise_bwd(dptr(_c(dout), torch.float32), dptr(parg), parg.shape[0], ctx.D, dptr(dpred),
                                           stream_ptr()), "sn2_project_plotwise_bwd")
        ops._count(1)
        return dpred, None, None


class TallLinear(torch.autograd.Function):
    """y = x W^T + b for a tall-skinny x (millions of edge rows, <= 64 columns).  Forward and dx are plain torch
    GEMMs; dW / db (a reduction over all rows with a tiny output) run in sn2_linear_wgrad: cuBLAS's large-K
    kernel took ~2 ms per edge layer at config 3, the streaming kernel is bandwidth bound."""

    NBLK = 148 * 4

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        return torch.addmm(bias, x, weight.t())

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, weight = ctx.saved_tensors
        dy = _c(dy)
        dx = dy @ weight if ctx.needs_input_grad[0] else None
        Co, Ci = weight.shape
        dW = torch.empty_like(weight)
        db = torch.empty(Co, dtype=torch.float32, device=dy.device)
        partial = torch.empty((TallLinear.NBLK, Co * (Ci + 1)), dtype=torch.float32, device=dy.device)
        check(lib.sn2_linear_wgrad(dptr(dy, torch.float32), dptr(_c(x), torch.float32), dy.shape[0], Co, Ci, dptr(partial),
                                   TallLinear.NBLK, dptr(dW), dptr(db), stream_ptr()), "sn2_linear_wgrad")
        ops._count(2)
        return dx, dW, db


class LinReluBN(torch.autograd.Function):
    """One train-mode block of the reference MLP() -- Linear -> ReLU -> BatchNorm1d with batch statistics
    (model/point_net2.py:45-53) -- in five kernels of csrc/train_mlp.cu instead of torch's nine.  Running statistics,
    `num_batches_tracked` and SyncBatchNorm (statistics all-reduced as raw fp64 sums + counts, forward and backward)
    behave as torch's modules do; `bn` is the block's own BatchNorm1d / SyncBatchNorm module."""

    NBLK = 148 * 3

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, bn, rows=None, in_ss=None, apply=True):
        """in_ss: scale | shift of the block that produced x with apply=False (x is then that block's un-normalised
        output and is read as x * scale + shift).  apply=False: return (y, ss) instead of z = y * scale + shift;
        the consumer (the next LinReluBN through in_ss, or SegmentMax through ss) applies ss on load and hands the
        gradient with respect to z back as the gradient of y."""
        lib = _lib.load()
        x = _c(x)
        if x.data_ptr() % 16:
            x = x.clone()
        rp = dptr(rows, torch.int32)
        R, (Co, Ci) = x.shape[0], weight.shape
        dev = x.device
        y = torch.empty((R, Co), dtype=torch.float32, device=dev)
        z = torch.empty((R, Co), dtype=torch.float32, device=dev) if apply else None
        ip = dptr(in_ss, torch.float32)
        stats = torch.empty(2 * Co + 1, dtype=torch.float64, device=dev)
        ss = torch.empty(4 * Co, dtype=torch.float32, device=dev)
        track = bn.track_running_stats and bn.running_mean is not None
        rm, rv = (dptr(bn.running_mean), dptr(bn.running_var)) if track else (None, None)
        nbt = dptr(bn.num_batches_tracked, torch.int64) if track and bn.num_batches_tracked is not None else None
        wp, bp = dptr(_c(weight), torch.float32), dptr(_c(bias), torch.float32)
        gp, btp = dptr(_c(gamma), torch.float32), dptr(_c(beta), torch.float32)
        st = stream_ptr()
        group = _sync_group(bn)
        peer = _comm.get_comm(group) if group is not None else None
        if group is None:
            check(lib.sn2_lrb_block_fwd(dptr(x, torch.float32), ip, wp, bp, gp, btp, float(bn.eps), float(bn.momentum), rm, rv, nbt,
                                        R, rp, Co, Ci, dptr(y), dptr(stats), dptr(ss), dptr(z), st), "sn2_lrb_block_fwd")
        elif peer is not None:
            # the statistics are summed over the ranks INSIDE the finalize kernel (NVLink peer stores, csrc/comm.cu)
            check(lib.sn2_lrb_fwd(dptr(x, torch.float32), ip, wp, bp, R, rp, Co, Ci, dptr(y), dptr(stats), st), "sn2_lrb_fwd")
            check(lib.sn2_bn_finalize_sync(peer.handle, dptr(stats), gp, btp, float(bn.eps), float(bn.momentum), rm, rv, nbt,
                                           dptr(ss), Co, st), "sn2_bn_finalize_sync")
            if apply:
                check(lib.sn2_bn_apply(dptr(y), dptr(ss), R, rp, Co, dptr(z), st), "sn2_bn_apply")
        else:
            check(lib.sn2_lrb_fwd(dptr(x, torch.float32), ip, wp, bp, R, rp, Co, Ci, dptr(y), dptr(stats), st), "sn2_lrb_fwd")
            torch.distributed.all_reduce(stats, group=group)
            check(lib.sn2_bn_finalize(dptr(stats), gp, btp, float(bn.eps), float(bn.momentum), rm, rv, nbt, dptr(ss), Co, st),
                  "sn2_bn_finalize")
            if apply:
                check(lib.sn2_bn_apply(dptr(y), dptr(ss), R, rp, Co, dptr(z), st), "sn2_bn_apply")
        ops._count(5 if apply else 4)
        ctx.save_for_backward(x, y, weight, ss, stats)
        ctx.group, ctx.rows, ctx.in_ss, ctx.peer = group, rows, in_ss, peer
        if apply:
            return z
        ctx.mark_non_differentiable(ss)
        return y, ss

    @staticmethod
    def backward(ctx, dz, _dss=None):
        lib = _lib.load()
        x, y, weight, ss, stats = ctx.saved_tensors
        ip = dptr(ctx.in_ss, torch.float32)
        dz = _c(dz)
        if dz.data_ptr() % 16:  # a view into a larger gradient buffer: the kernels fetch 16-byte aligned tiles
            dz = dz.clone()
        R, (Co, Ci) = x.shape[0], weight.shape
        dev = x.device
        sums = torch.empty(2 * Co, dtype=torch.float64, device=dev)
        dgb = torch.empty((2, Co), dtype=torch.float32, device=dev)
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dW = torch.empty_like(weight)
        db = torch.empty(Co, dtype=torch.float32, device=dev)
        partial = torch.empty((LinReluBN.NBLK, Co * (Ci + 1)), dtype=torch.float32, device=dev)
        wp, st, rp = dptr(_c(weight), torch.float32), stream_ptr(), dptr(ctx.rows, torch.int32)
        dgp, dbp = ctypes.c_void_p(dgb.data_ptr()), ctypes.c_void_p(dgb.data_ptr() + 4 * Co)
        if ctx.group is None:
            check(lib.sn2_lrb_block_bwd(dptr(dz, torch.float32), dptr(y), dptr(x), ip, wp, dptr(ss), dptr(stats), R, rp, Co, Ci, dptr(sums),
                                        dgp, dbp, dptr(dx), dptr(partial), LinReluBN.NBLK, dptr(dW), dptr(db), st),
                  "sn2_lrb_block_bwd")
        elif ctx.peer is not None:
            check(lib.sn2_lrb_bwd_reduce(dptr(dz, torch.float32), dptr(y), R, rp, Co, dptr(sums), st), "sn2_lrb_bwd_reduce")
            check(lib.sn2_bn_bwd_sync(ctx.peer.handle, dptr(sums), dptr(ss), Co, dgp, dbp, st), "sn2_bn_bwd_sync")
            check(lib.sn2_lrb_bwd(dptr(dz), dptr(y), dptr(x), ip, wp, dptr(ss), dptr(sums), dptr(stats), R, rp, Co, Ci, dptr(dx),
                                  dptr(partial), LinReluBN.NBLK, dptr(dW), dptr(db), st), "sn2_lrb_bwd")
        else:
            check(lib.sn2_lrb_bwd_reduce(dptr(dz, torch.float32), dptr(y), R, rp, Co, dptr(sums), st), "sn2_lrb_bwd_reduce")
            check(lib.sn2_bn_param_grad(dptr(sums), dptr(ss), Co, dgp, dbp, st), "sn2_bn_param_grad")  # this rank's sums
            torch.distributed.all_reduce(sums, group=ctx.group)
            check(lib.sn2_lrb_bwd(dptr(dz), dptr(y), dptr(x), ip, wp, dptr(ss), dptr(sums), dptr(stats), R, rp, Co, Ci, dptr(dx),
                                  dptr(partial), LinReluBN.NBLK, dptr(dW), dptr(db), st), "sn2_lrb_bwd")
        ops._count(5)
        return dx, dW, db, dgb[0], dgb[1], None, None, None, None


def _bn_stats_to_ss(lib, bn, stats, gamma, beta, Co, st):
    """Raw fp64 sums (+ count) -> ss = scale | shift | mean | invstd with running-stat update, as LinReluBN does it:
    plain BatchNorm1d, SyncBatchNorm over the peer communicator (summed inside the kernel), or over NCCL."""
    ss = torch.empty(4 * Co, dtype=torch.float32, device=stats.device)
    track = bn.track_running_stats and bn.running_mean is not None
    rm, rv = (dptr(bn.running_mean), dptr(bn.running_var)) if track else (None, None)
    nbt = dptr(bn.num_batches_tracked, torch.int64) if track and bn.num_batches_tracked is not None else None
    gp, btp = dptr(_c(gamma), torch.float32), dptr(_c(beta), torch.float32)
    group = _sync_group(bn)
    peer = _comm.get_comm(group) if group is not None else None
    if peer is not None:
        check(lib.sn2_bn_finalize_sync(peer.handle, dptr(stats), gp, btp, float(bn.eps), float(bn.momentum), rm, rv, nbt,
                                       dptr(ss), Co, st), "sn2_bn_finalize_sync")
    else:
        if group is not None:
            torch.distributed.all_reduce(stats, group=group)
        check(lib.sn2_bn_finalize(dptr(stats), gp, btp, float(bn.eps), float(bn.momentum), rm, rv, nbt, dptr(ss), Co, st),
              "sn2_bn_finalize")
    return ss, group, peer


def _bn_bwd_sums(lib, sums, ss, Co, group, peer, st):
    """This rank's raw backward sums -> (dgamma, dbeta); afterwards `sums` holds the sums over all ranks."""
    dgb = torch.empty((2, Co), dtype=torch.float32, device=sums.device)
    dgp, dbp = ctypes.c_void_p(dgb.data_ptr()), ctypes.c_void_p(dgb.data_ptr() + 4 * Co)
    if peer is not None:
        check(lib.sn2_bn_bwd_sync(peer.handle, dptr(sums), dptr(ss), Co, dgp, dbp, st), "sn2_bn_bwd_sync")
    else:
        check(lib.sn2_bn_param_grad(dptr(sums), dptr(ss), Co, dgp, dbp, st), "sn2_bn_param_grad")
        if group is not None:
            torch.distributed.all_reduce(sums, group=group)
    return dgb[0], dgb[1]


class SA1Recompute(torch.autograd.Function):
    """The whole train-mode sa1 block -- PointConv message, MLP([11,16,16]) = 2 x (Linear -> ReLU -> BatchNorm1d with
    batch statistics), max aggregation (reference model/point_net2.py:21-29, :45-53) -- without materialising any
    per-edge array: every sweep recomputes the messages from a per-point table (csrc/train_sa.cu).  Same statistics
    protocol as LinReluBN (running statistics, num_batches_tracked, SyncBatchNorm)."""

    @staticmethod
    def supported(seq, feat) -> bool:
        blocks = list(seq)
        if len(blocks) != 2 or feat.shape[1] != 8 or feat.dtype != torch.float32:
            return False
        for blk, shape in zip(blocks, ((16, 11), (16, 16))):
            layers = list(blk)
            if len(layers) != 3 or not isinstance(layers[0], torch.nn.Linear) or not isinstance(layers[1], torch.nn.ReLU):
                return False
            lin, bn = layers[0], layers[2]
            if not isinstance(bn, (torch.nn.BatchNorm1d, torch.nn.SyncBatchNorm)):
                return False
            if not (bn.training and bn.affine and bn.momentum is not None and lin.bias is not None
                    and tuple(lin.weight.shape) == shape):
                return False
        return True

    @staticmethod
    def forward(ctx, feat, pos4, qpos4, rowptr, col, W1, b1, g1, bt1, W2, b2, g2, bt2, bn1, bn2):
        lib = _lib.load()
        feat, W1, b1, W2, b2, g2 = _c(feat), _c(W1), _c(b1), _c(W2), _c(b2), _c(g2)
        P, M, dev, st = feat.shape[0], qpos4.shape[0], feat.device, stream_ptr()
        f32 = torch.float32
        u = torch.empty((P, 16), dtype=f32, device=dev)
        stats1 = torch.empty(33, dtype=torch.float64, device=dev)
        stats2 = torch.empty(33, dtype=torch.float64, device=dev)
        key = torch.empty((M, 16), dtype=f32, device=dev)
        arg = torch.empty((M, 16), dtype=torch.int32, device=dev)
        x1 = torch.empty((M, 16), dtype=f32, device=dev)
        amax = torch.empty((M, 16), dtype=f32, device=dev)
        queue = torch.empty(4, dtype=torch.int32, device=dev)  # work-queue cursors: each sweep resets its own
        rp, cp, qp = dptr(rowptr, torch.int32), dptr(col, torch.int32), dptr(qpos4)
        check(lib.sn2_sa1t_pre(dptr(feat, f32), dptr(pos4), P, dptr(W1, f32), dptr(u), st), "sn2_sa1t_pre")
        check(lib.sn2_sa1t_stats1(dptr(u), qp, rp, cp, M, dptr(W1), dptr(b1, f32), dptr(stats1), dptr(queue), st), "sn2_sa1t_stats1")
        ss1, group1, peer1 = _bn_stats_to_ss(lib, bn1, stats1, g1, bt1, 16, st)
        check(lib.sn2_sa1t_stats2(dptr(u), qp, rp, cp, M, dptr(W1), dptr(b1), dptr(W2, f32), dptr(b2, f32), dptr(ss1),
                                  dptr(g2, f32), dptr(stats2), dptr(key), dptr(arg), dptr(queue), st), "sn2_sa1t_stats2")
        ss2, group2, peer2 = _bn_stats_to_ss(lib, bn2, stats2, g2, bt2, 16, st)
        check(lib.sn2_sa1t_finish(dptr(key), dptr(arg), dptr(g2), dptr(ss2), M, dptr(x1), dptr(amax), st), "sn2_sa1t_finish")
        ops._count(10)  # pre; 2 x (stats init, constant image, sweep, BatchNorm finalize); finish
        ctx.save_for_backward(feat, pos4, qpos4, rowptr, col, W1, b1, W2, b2, g2, u, ss1, ss2, stats1, stats2, arg, amax)
        ctx.sync = (group1, peer1, group2, peer2)
        return x1

    @staticmethod
    def backward(ctx, dx1):
        lib = _lib.load()
        feat, pos4, qpos4, rowptr, col, W1, b1, W2, b2, g2, u, ss1, ss2, stats1, stats2, arg, amax = ctx.saved_tensors
        group1, peer1, group2, peer2 = ctx.sync
        P, M, dev, st = feat.shape[0], qpos4.shape[0], feat.device, stream_ptr()
        f32 = torch.float32
        dx1 = _c(dx1)
        if dx1.data_ptr() % 16:
            dx1 = dx1.clone()
        sums1 = torch.empty(32, dtype=torch.float64, device=dev)
        sums2 = torch.empty(32, dtype=torch.float64, device=dev)
        partial = torch.empty((int(lib.sn2_sa1t_blocks()), int(lib.sn2_sa1t_partials())), dtype=f32, device=dev)
        dW1, db1 = torch.empty_like(W1), torch.empty_like(b1)
        dW2, db2 = torch.empty_like(W2), torch.empty_like(b2)
        du = torch.empty((P, 16), dtype=f32, device=dev)
        dc = torch.empty((M, 16), dtype=f32, device=dev)
        queue = torch.empty(4, dtype=torch.int32, device=dev)
        rp, cp, qp = dptr(rowptr, torch.int32), dptr(col, torch.int32), dptr(qpos4)
        check(lib.sn2_sa1t_bwd_sums(dptr(dx1, f32), dptr(amax), dptr(arg), M, dptr(sums2), st), "sn2_sa1t_bwd_sums")
        dg2, dbt2 = _bn_bwd_sums(lib, sums2, ss2, 16, group2, peer2, st)
        check(lib.sn2_sa1t_bwd_w2(dptr(u), qp, rp, cp, M, dptr(W1), dptr(b1), dptr(W2), dptr(b2), dptr(g2), dptr(ss1), dptr(ss2),
                                  dptr(stats2), dptr(sums2), dptr(dx1), dptr(arg), dptr(partial), dptr(dW2), dptr(db2),
                                  dptr(sums1), dptr(queue), st), "sn2_sa1t_bwd_w2")
        dg1, dbt1 = _bn_bwd_sums(lib, sums1, ss1, 16, group1, peer1, st)
        check(lib.sn2_sa1t_bwd_in(dptr(u), qp, rp, cp, P, M, dptr(W1), dptr(b1), dptr(W2), dptr(b2), dptr(g2), dptr(ss1),
                                  dptr(ss2), dptr(stats1), dptr(stats2), dptr(sums1), dptr(sums2), dptr(dx1), dptr(arg),
                                  dptr(du), dptr(dc), dptr(queue), st), "sn2_sa1t_bwd_in")
        check(lib.sn2_sa1t_bwd_w1(dptr(du), dptr(dc), dptr(feat), dptr(pos4), qp, P, M, dptr(partial), dptr(dW1), dptr(db1), st),
              "sn2_sa1t_bwd_w1")
        ops._count(10)  # sparse sums, 2 x BatchNorm parameter gradients, 2 x (constant image, sweep), W2 finish, W1 + its reduction
        dfeat = du @ W1[:, :8] if ctx.needs_input_grad[0] else None
        return dfeat, None, None, None, None, dW1, db1, dg1, dbt1, dW2, db2, dg2, dbt2, None, None


class SA2Recompute(torch.autograd.Function):
    """The train-mode sa2 block -- PointConv message, MLP([19,32]) = Linear -> ReLU -> BatchNorm1d with batch
    statistics, max aggregation -- without per-edge arrays (csrc/train_sa.cu, lane = channel); see SA1Recompute."""

    @staticmethod
    def supported(seq, x) -> bool:
        blocks = list(seq)
        if len(blocks) != 1 or x.shape[1] != 16 or x.dtype != torch.float32:
            return False
        layers = list(blocks[0])
        if len(layers) != 3 or not isinstance(layers[0], torch.nn.Linear) or not isinstance(layers[1], torch.nn.ReLU):
            return False
        lin, bn = layers[0], layers[2]
        if not isinstance(bn, (torch.nn.BatchNorm1d, torch.nn.SyncBatchNorm)):
            return False
        return bool(bn.training and bn.affine and bn.momentum is not None and lin.bias is not None
                    and tuple(lin.weight.shape) == (32, 19))

    @staticmethod
    def forward(ctx, x, pos4, qpos4, rowptr, col, W, b, g, bt, bn):
        lib = _lib.load()
        x, W, b, g = _c(x), _c(W), _c(b), _c(g)
        P, M, dev, st = x.shape[0], qpos4.shape[0], x.device, stream_ptr()
        f32 = torch.float32
        u = torch.empty((P, 32), dtype=f32, device=dev)
        stats = torch.empty(65, dtype=torch.float64, device=dev)
        key = torch.empty((M, 32), dtype=f32, device=dev)
        arg = torch.empty((M, 32), dtype=torch.int32, device=dev)
        x2 = torch.empty((M, 32), dtype=f32, device=dev)
        amax = torch.empty((M, 32), dtype=f32, device=dev)
        queue = torch.empty(4, dtype=torch.int32, device=dev)
        rp, cp, qp = dptr(rowptr, torch.int32), dptr(col, torch.int32), dptr(qpos4)
        check(lib.sn2_sa2t_pre(dptr(x, f32), dptr(pos4), P, dptr(W, f32), dptr(u), st), "sn2_sa2t_pre")
        check(lib.sn2_sa2t_fwd(dptr(u), qp, rp, cp, M, dptr(W), dptr(b, f32), dptr(g, f32), dptr(stats), dptr(key), dptr(arg),
                               dptr(queue), st), "sn2_sa2t_fwd")
        ss, group, peer = _bn_stats_to_ss(lib, bn, stats, g, bt, 32, st)
        check(lib.sn2_sa2t_finish(dptr(key), dptr(arg), dptr(g), dptr(ss), M, dptr(x2), dptr(amax), st), "sn2_sa2t_finish")
        ops._count(5)
        ctx.save_for_backward(x, pos4, qpos4, rowptr, col, W, b, g, u, ss, stats, arg, amax)
        ctx.sync = (group, peer)
        return x2

    @staticmethod
    def backward(ctx, dx2):
        lib = _lib.load()
        x, pos4, qpos4, rowptr, col, W, b, g, u, ss, stats, arg, amax = ctx.saved_tensors
        group, peer = ctx.sync
        P, M, dev, st = x.shape[0], qpos4.shape[0], x.device, stream_ptr()
        f32 = torch.float32
        dx2 = _c(dx2)
        sums = torch.empty(64, dtype=torch.float64, device=dev)
        partial = torch.empty((int(lib.sn2_sa1t_blocks()), int(lib.sn2_sa1t_partials())), dtype=f32, device=dev)
        dW, db = torch.empty_like(W), torch.empty_like(b)
        du = torch.empty((P, 32), dtype=f32, device=dev)
        dc = torch.empty((M, 32), dtype=f32, device=dev)
        queue = torch.empty(4, dtype=torch.int32, device=dev)
        rp, cp, qp = dptr(rowptr, torch.int32), dptr(col, torch.int32), dptr(qpos4)
        check(lib.sn2_sa2t_bwd_sums(dptr(dx2, f32), dptr(amax), dptr(arg), M, dptr(sums), st), "sn2_sa2t_bwd_sums")
        dg, dbt = _bn_bwd_sums(lib, sums, ss, 32, group, peer, st)
        check(lib.sn2_sa2t_bwd(dptr(u), qp, rp, cp, P, M, dptr(W), dptr(b), dptr(g), dptr(ss), dptr(stats), dptr(sums), dptr(dx2),
                               dptr(arg), dptr(du), dptr(dc), dptr(queue), st), "sn2_sa2t_bwd")
        check(lib.sn2_sa2t_bwd_w(dptr(du), dptr(dc), dptr(x), dptr(pos4), qp, P, M, dptr(partial), dptr(dW), dptr(db), st),
              "sn2_sa2t_bwd_w")
        ops._count(5)
        dx = du @ W[:, :16] if ctx.needs_input_grad[0] else None
        return dx, None, None, None, None, dW, db, dg, dbt, None


class Head(torch.autograd.Function):
    """relu(lin1) -> lin2 -> softmax(4) x sigmoid(1) of reference model/point_net2.py:141-153 (dropout p = 0) in one
    kernel each way; the backward recomputes the head from its input row, nothing else is saved.
    in_ss: scale | shift of the producing LinReluBN block when its BatchNorm transform was deferred."""

    NBLK = 148 * 4

    @staticmethod
    def forward(ctx, f1, W1, b1, W2, b2, in_ss=None):
        lib = _lib.load()
        f1 = _c(f1)
        R = f1.shape[0]
        if f1.shape[1] != 34 or tuple(W1.shape) != (16, 34) or tuple(W2.shape) != (5, 16):
            raise RuntimeError("sn2 Head: built for lin1 [34 -> 16], lin2 [16 -> 5] (the reference head)")
        cov = torch.empty((R, 4), dtype=torch.float32, device=f1.device)
        proba = torch.empty((R, 4), dtype=torch.float32, device=f1.device)
        check(lib.sn2_head_fwd(dptr(f1, torch.float32), dptr(in_ss, torch.float32), dptr(_c(W1), torch.float32), dptr(_c(b1)),
                               dptr(_c(W2)), dptr(_c(b2)), R, dptr(cov), dptr(proba), stream_ptr()), "sn2_head_fwd")
        ops._count(1)
        ctx.save_for_backward(f1, W1, b1, W2, b2)
        ctx.in_ss = in_ss
        return cov, proba

    @staticmethod
    def backward(ctx, dcov, dproba):
        lib = _lib.load()
        f1, W1, b1, W2, b2 = ctx.saved_tensors
        R, dev = f1.shape[0], f1.device
        df1 = torch.empty_like(f1) if ctx.needs_input_grad[0] else None
        grads = torch.empty(16 * 34 + 16 + 5 * 16 + 5, dtype=torch.float32, device=dev)
        dW1, db1, dW2, db2 = grads[:544].view(16, 34), grads[544:560], grads[560:640].view(5, 16), grads[640:645]
        partial = torch.empty((Head.NBLK, int(lib.sn2_head_bwd_partials())), dtype=torch.float32, device=dev)
        dcp = dptr(_c(dcov), torch.float32) if dcov is not None else None
        dpp = dptr(_c(dproba), torch.float32) if dproba is not None else None
        check(lib.sn2_head_bwd(dptr(f1), dptr(ctx.in_ss, torch.float32), dptr(_c(W1)), dptr(_c(b1)), dptr(_c(W2)), dptr(_c(b2)), dcp, dpp, R,
                               dptr(df1), dptr(partial), Head.NBLK, dptr(dW1), ctypes.c_void_p(db1.data_ptr()),
                               ctypes.c_void_p(dW2.data_ptr()), ctypes.c_void_p(db2.data_ptr()), stream_ptr()), "sn2_head_bwd")
        ops._count(2)
        return df1, dW1, db1, dW2, db2, None


def _sync_group(bn):
    """Process group over which a SyncBatchNorm shares its statistics (None for plain BatchNorm1d / single process)."""
    if not isinstance(bn, torch.nn.SyncBatchNorm):
        return None
    if not (torch.distributed.is_available() and torch.distributed.is_initialized()):
        return None
    group = bn.process_group if bn.process_group is not None else torch.distributed.group.WORLD
    return group if torch.distributed.get_world_size(group) > 1 else None


def _fusable_block(lib, block, x) -> bool:
    layers = list(block)
    if len(layers) != 3 or not isinstance(layers[0], torch.nn.Linear) or not isinstance(layers[1], torch.nn.ReLU):
        return False
    lin, bn = layers[0], layers[2]
    if not isinstance(bn, (torch.nn.BatchNorm1d, torch.nn.SyncBatchNorm)):
        return False
    return (bn.training and bn.affine and bn.momentum is not None and lin.bias is not None and x.dtype == torch.float32
            and bool(lib.sn2_lrb_supported(lin.out_features, lin.in_features)))


def tall_linear(lin, x):
    """lin(x) with the streaming weight-gradient kernel when x is tall (the head's lin1 / lin2 over every point)."""
    lib = _lib.load()
    if (x.shape[0] >= 65536 and lin.bias is not None and (x.requires_grad or lin.weight.requires_grad)
            and lib.sn2_linear_wgrad_supported(lin.out_features, lin.in_features)):
        return TallLinear.apply(x, lin.weight, lin.bias)
    return lin(x)


def run_mlp(seq, x, rows=None, defer_last: bool = False):
    """Apply a reference MLP (Sequential of (Linear, ReLU, BatchNorm1d) blocks).  Blocks of a supported shape (all seven of
    the network; SN2_FUSED_MLP_MIN_ROWS sets a row floor) in training run as the fused LinReluBN (SN2_FUSED_MLP=0 -> the torch modules, with TallLinear where it pays); between
    two fused blocks the BatchNorm transform is not materialised (the consumer applies it on load).
    rows: device int32 [1] live row count when x is a fixed-capacity buffer (graph replay); needs the fused blocks.
    defer_last: return (y, ss) instead of z when the last block is fused -- for SegmentMax(y, rowptr, ss); ss is
    None when the last block was not fused (then y is the finished output)."""
    lib = _lib.load()
    import os
    fused = os.environ.get("SN2_FUSED_MLP", "1") == "1"
    tall = os.environ.get("SN2_TALL_LINEAR", "1") == "1"
    defer_ok = os.environ.get("SN2_DEFER_BN", "1") == "1"
    blocks = list(seq)
    min_rows = int(os.environ.get("SN2_FUSED_MLP_MIN_ROWS", "1"))  # every block of the network has a fused kernel
    fusable = [fused and x.shape[0] >= min_rows and _fusable_block(lib, b, x) for b in blocks]
    in_ss = None
    for i, block in enumerate(blocks):
        lin = block[0]
        if fusable[i]:
            bn = block[2]
            last = i + 1 == len(blocks)
            defer = defer_ok and ((not last and fusable[i + 1]) or (last and defer_last))
            res = LinReluBN.apply(x, lin.weight, lin.bias, bn.weight, bn.bias, bn, rows, in_ss, not defer)
            x, in_ss = res if defer else (res, None)
            continue
        if rows is not None:
            raise RuntimeError("sn2: a fixed-capacity edge buffer needs the fused Linear-ReLU-BatchNorm blocks")
        if tall and x.shape[0] >= 65536 and (x.requires_grad or lin.weight.requires_grad) and lib.sn2_linear_wgrad_supported(
                lin.out_features, lin.in_features):
            x = TallLinear.apply(x, lin.weight, lin.bias)
        else:
            x = lin(x)
        for layer in list(block)[1:]:
            x = layer(x)
    return (x, in_ss) if defer_last else x
