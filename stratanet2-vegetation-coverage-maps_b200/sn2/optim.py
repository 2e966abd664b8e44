"""FusedAdam: the reference's optimizer step (learning/train.py:180-185: ``Adam(lr=1e-3, weight_decay=wd)`` stepped
by ``StepLR``) as ONE kernel over flat parameter / gradient buckets, with the data-parallel gradient all-reduce
inside it (csrc/comm.cu::adam_sync_kernel).

torch.optim.Adam runs a dozen foreach kernels over 32 small tensors per step; the model has 14 997 parameters, so
the whole update is one CTA's work.  All parameters become views of one contiguous fp32 buffer, all gradients views
of another (they stay views: ``zero_grad`` zeroes the bucket).  With a ``PeerComm`` the kernel first sums
``grad_scale * grad`` over the ranks in rank order through NVLink peer stores -- every rank computes the same bits,
so the replicas stay identical without a broadcast.  The learning rate and the step count live in device memory: a
captured CUDA graph follows a scheduler (``StepLR.step()`` only changes ``param_groups[0]["lr"]``) without
re-capture.

Same arithmetic as ``torch.optim.Adam(amsgrad=False, maximize=False)`` with L2 weight decay, up to fp32 rounding.
"""
from __future__ import annotations

import torch

from . import _lib, ops
from ._lib import check, dptr, stream_ptr


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, comm=None):
        params = [p for p in params if p.requires_grad]
        if not params:
            raise ValueError("FusedAdam: no parameters")
        dev = params[0].device
        if any((not p.is_cuda) or p.device != dev or p.dtype != torch.float32 for p in params):
            raise RuntimeError("sn2 FusedAdam: fp32 CUDA parameters on one device")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) != 1:
            raise RuntimeError("sn2 FusedAdam: one parameter group")
        self.comm = comm
        self.nccl_group = None  # process group of the NCCL path (None = WORLD); False = never all-reduce (local training)
        n = sum(p.numel() for p in params)
        self.flat_param = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in params:
                k = p.numel()
                self.flat_param[off:off + k].copy_(p.reshape(-1))
                p.data = self.flat_param[off:off + k].view_as(p)
                p.grad = self.flat_grad[off:off + k].view_as(p)
                off += k
        self._params = params
        self.exp_avg = torch.zeros_like(self.flat_param)
        self.exp_avg_sq = torch.zeros_like(self.flat_param)
        self.step_dev = torch.zeros((), dtype=torch.int64, device=dev)
        self.lr_dev = torch.full((), float(lr), dtype=torch.float32, device=dev)
        self._lr_host = float(lr)
        # exposed like torch's per-parameter state so that generic code (GraphedTrainStep's roll-back) can snapshot it
        self.state[params[0]] = {"exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "step": self.step_dev}

    def zero_grad(self, set_to_none: bool = False):  # gradients must stay views of the bucket
        self.flat_grad.zero_()

    def sync_hyperparameters(self):
        """Push a changed learning rate (scheduler) to the device scalar the kernel reads.  Never synchronises."""
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_host:
            self.lr_dev.fill_(lr)
            self._lr_host = lr

    def _check_views(self):
        base = self.flat_grad.data_ptr()
        for p in self._params:
            if p.grad is None or not (base <= p.grad.data_ptr() < base + 4 * self.flat_grad.numel()):
                raise RuntimeError("sn2 FusedAdam: a gradient was re-allocated (use optimizer.zero_grad(), not set_to_none)")

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        """grad_scale: weight of this rank's gradient in the all-reduce (plots of this rank / plots of the global
        batch, so that the sum is the gradient of the mean loss over the global batch); 1.0 for a single process."""
        if closure is not None:
            raise RuntimeError("sn2 FusedAdam: closures are not supported")
        if not torch.cuda.is_current_stream_capturing():
            self._check_views()
            self.sync_hyperparameters()
        g = self.param_groups[0]
        b1, b2 = g["betas"]
        comm = self.comm.handle if self.comm is not None else None
        if comm is None and self.nccl_group is not False:
            # NCCL path (SN2_COMM=nccl or no peer mapping): the same weighted sum as a library all-reduce before the kernel
            import torch.distributed as dist

            if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.nccl_group) > 1:
                self.flat_grad.mul_(float(grad_scale))
                dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.nccl_group)
                grad_scale = 1.0
        check(_lib.load().sn2_adam_step(comm, dptr(self.flat_grad), float(grad_scale), dptr(self.flat_param), dptr(self.exp_avg),
                                        dptr(self.exp_avg_sq), self.flat_param.numel(), dptr(self.lr_dev), float(b1), float(b2),
                                        float(g["eps"]), float(g["weight_decay"]), dptr(self.step_dev), stream_ptr()), "sn2_adam_step")
        ops._count(1)
