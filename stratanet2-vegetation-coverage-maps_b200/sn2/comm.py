"""Peer-memory communicator for data-parallel training (csrc/comm.cu).

One process per GPU.  ``torch.distributed`` (NCCL) is the plumbing: rendezvous, the exchange of the CUDA IPC
handles, barriers.  The data path of the training step does not call it: the BatchNorm statistics of every
Linear-ReLU-BatchNorm block and the flat gradient are summed INSIDE our own kernels by stores into the peers'
memory over NVLink (one-shot all-reduce, rank-order summation, CUDA-graph capturable), see csrc/comm.cu.

``get_comm()`` returns the process-wide communicator, or None when the job is a single process, when
``SN2_COMM=nccl`` asks for the NCCL path (torch.distributed.all_reduce between the kernels -- the baseline the fused
path is measured against), or when peer mapping is not possible on this machine (every rank then agrees to fall
back, with a note on stderr).
"""
from __future__ import annotations

import ctypes
import os
import sys

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check, dptr, stream_ptr


class PeerComm:
    def __init__(self, group=None):
        lib = _lib.load()
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if self.world > lib.sn2_comm_max_world():
            raise RuntimeError(f"sn2 PeerComm: at most {lib.sn2_comm_max_world()} ranks (one NVSwitch domain)")
        self.max_bytes = int(lib.sn2_comm_max_bytes())
        self._region, self._peers, self.handle = None, [], None
        region = ctypes.c_void_p()
        handle, err = (ctypes.c_ubyte * 64)(), None
        try:
            check(lib.sn2_comm_region_alloc(ctypes.byref(region)), "sn2_comm_region_alloc")
            self._region = region
            check(lib.sn2_comm_ipc_export(region, handle), "sn2_comm_ipc_export")
        except RuntimeError as e:
            err = e
        handles = [None] * self.world
        dist.all_gather_object(handles, None if err is not None else bytes(handle), group=group)  # reached by every rank
        if err is not None:
            raise err
        if any(h is None for h in handles):
            raise RuntimeError("a peer could not export its region")
        table = (ctypes.c_void_p * self.world)()
        for r, h in enumerate(handles):
            if r == self.rank:
                table[r] = region.value
                continue
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
            p = ctypes.c_void_p()
            check(lib.sn2_comm_ipc_import(buf, ctypes.byref(p)), "sn2_comm_ipc_import")
            self._peers.append(p)
            table[r] = p.value
        comm = ctypes.c_void_p()
        check(lib.sn2_comm_create(self.rank, self.world, table, ctypes.byref(comm)), "sn2_comm_create")
        self.handle = comm

    def status(self):
        """(collectives completed, sticky error word) -- synchronous; error != 0 means a wait timed out."""
        seq, err = ctypes.c_longlong(), ctypes.c_longlong()
        check(_lib.load().sn2_comm_status(self.handle, ctypes.byref(seq), ctypes.byref(err)), "sn2_comm_status")
        return seq.value, err.value

    def check_healthy(self):
        seq, err = self.status()
        if err:
            raise RuntimeError(f"sn2 PeerComm: a peer did not arrive at collective {err} (rank {self.rank}, {seq} completed)")

    def all_reduce_(self, t: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
        """In-place sum over the ranks (fp32: of scale * t) on the current stream."""
        lib = _lib.load()
        if not t.is_contiguous() or t.numel() * t.element_size() > self.max_bytes:
            raise RuntimeError("sn2 PeerComm.all_reduce_: contiguous tensors of at most 64 KiB")
        if t.dtype == torch.float64:
            if scale != 1.0:
                raise RuntimeError("sn2 PeerComm.all_reduce_: scale is fp32 only")
            check(lib.sn2_comm_allreduce_f64(self.handle, dptr(t), t.numel(), stream_ptr()), "sn2_comm_allreduce_f64")
        elif t.dtype == torch.float32:
            check(lib.sn2_comm_allreduce_f32(self.handle, dptr(t), t.numel(), float(scale), stream_ptr()), "sn2_comm_allreduce_f32")
        else:
            raise RuntimeError("sn2 PeerComm.all_reduce_: float32 / float64 only")
        from . import ops

        ops._count(1)
        return t

    def close(self):
        lib = _lib.load(require_cuda=False)
        torch.cuda.synchronize()
        if self.handle is not None:
            lib.sn2_comm_destroy(self.handle)
            self.handle = None
        for p in self._peers:
            lib.sn2_comm_ipc_release(p)
        self._peers = []
        if self._region is not None:
            lib.sn2_comm_region_free(self._region)
            self._region = None


_COMMS: dict = {}


def backend() -> str:
    """'peer' (default) or 'nccl' (SN2_COMM=nccl): how the training step's small all-reduces run."""
    return os.environ.get("SN2_COMM", "peer").lower()


def get_comm(group=None):
    """Process-wide PeerComm of `group` (WORLD by default), created collectively on first use; None = no peer path."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1 or backend() != "peer":
        return None
    key = id(group) if group is not None else 0
    if key not in _COMMS:
        comm, ok = None, 1
        try:
            comm = PeerComm(group)
        except Exception as e:  # noqa: BLE001 -- any failure (IPC not permitted, no peer access ...) means: fall back, everywhere
            ok = 0
            print(f"[sn2.comm] rank {dist.get_rank(group)}: peer mapping failed ({e}); falling back to NCCL all-reduces", file=sys.stderr)
        # one collective that every rank reaches whatever happened above: agreement + "every region is zeroed and mapped
        # before anybody pushes into it"
        flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            if comm is not None:
                comm.close()
            comm = None
        _COMMS[key] = comm
    return _COMMS[key]


def close_all():
    for c in _COMMS.values():
        if c is not None:
            c.close()
    _COMMS.clear()
