"""Data parallelism over plots: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch).

The path shards naturally by plot: every operator is per plot, so inference needs no
data-path collective (results are gathered for map fusion only), and training couples the ranks in two
places only:
  * the gradient sum -- ONE all-reduce of a flat fp32 bucket holding all 14 997 gradients (59 988 bytes,
    latency bound: a single bucket, no overlap machinery needed);
  * BatchNorm batch statistics -- ``convert_sync_batchnorm`` swaps the model's BatchNorm1d modules for
    SyncBatchNorm containers (same state_dict keys).  The fused Linear-ReLU-BatchNorm blocks
    (sn2/autograd_ops.py::LinReluBN) see that type and sum their RAW fp64 statistics over the ranks -- per
    channel (sum y, sum y^2, row count) forward and (sum dz, sum dz*y) backward, one small vector each -- so a
    global batch split over G GPUs normalises exactly like the single-process reference even though the
    number of edge messages differs per rank (counts are reduced, not assumed equal).  The sum runs either as
    an NCCL all-reduce or inside the finalize kernels over NVLink peer memory (sn2/comm.py).  Without the
    conversion training is "local-BN" DP (a stated deviation).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_plots: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous plot range of ``rank`` (sizes differ by at most one plot)."""
    base, rem = divmod(n_plots, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_plots(cloud_data: dict, rank: int, world: int) -> dict:
    """Slice every per-plot tensor of a batch dict ((B, ...) leading dim) to this rank's plots."""
    B = cloud_data["cloud"].shape[0]
    lo, hi = shard_bounds(B, rank, world)
    out = {}
    for k, v in cloud_data.items():
        out[k] = v[lo:hi] if isinstance(v, torch.Tensor) and v.dim() > 0 and v.shape[0] == B else v
    return out


class GradBucket:
    """All gradients of a model in one contiguous fp32 buffer; ``allreduce`` = one collective per step."""

    def __init__(self, model: torch.nn.Module):
        self.params = [p for p in model.parameters() if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:  # gradients become views into the bucket: no copy before / after the collective
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def allreduce(self, local_plots: int, global_plots: int, group=None):
        """Weighted sum so the result equals the gradient of the loss averaged over ALL plots of the global
        batch (every rank's loss is a mean over its own plots / points; plots have equal point counts)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        for p, _ in zip(self.params, range(len(self.params))):
            if p.grad is None or p.grad.data_ptr() < self.flat.data_ptr():
                raise RuntimeError("GradBucket: a gradient was re-allocated; use bucket.zero() instead of set_to_none")
        self.flat.mul_(local_plots / float(global_plots))
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)


def convert_sync_batchnorm(model: torch.nn.Module, process_group=None) -> torch.nn.Module:
    """BatchNorm1d -> SyncBatchNorm in place of the same attribute names (state_dict keys unchanged)."""
    return torch.nn.SyncBatchNorm.convert_sync_batchnorm(model, process_group)


def gather_plot_results(t: torch.Tensor, group=None) -> torch.Tensor | None:
    """Inference: gather per-plot results ((b_rank, ...) on each rank) to rank 0 in plot order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return t
    world = dist.get_world_size(group)
    sizes = torch.zeros(world, dtype=torch.int64, device=t.device)
    sizes[dist.get_rank(group)] = t.shape[0]
    dist.all_reduce(sizes, group=group)
    mx = int(sizes.max())
    pad = torch.zeros((mx,) + t.shape[1:], dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    if dist.get_rank(group) != 0:
        return None
    return torch.cat([o[: int(s)] for o, s in zip(outs, sizes.tolist())], dim=0)
