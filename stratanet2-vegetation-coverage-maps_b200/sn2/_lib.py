"""ctypes binding of libsn2_b200.so (the C ABI declared in include/sn2.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is usable, every entry
point raises RuntimeError.  Build it in-tree with ``python stratanet2-vegetation-coverage-maps_b200/build.py``
(or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes
import os

import torch

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG, "lib", "libsn2_b200.so")
_lib = None

_vp, _i, _f, _ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_longlong

# name -> argtypes (restype int unless listed in _RESTYPES); mirrors include/sn2.h one to one
SIGNATURES = {
    "sn2_abi_version": [],
    "sn2_error_string": [_i],
    "sn2_last_cuda_error": [],
    "sn2_ingest": [_vp, _vp, _i, _i, _i, _vp, _vp, _vp],
    "sn2_fps_max_points": [],
    "sn2_fps": [_vp, _i, _i, _i, _vp, _vp, _vp, _vp],
    "sn2_fps_algo": [_vp, _i, _i, _i, _vp, _vp, _vp, _i, _vp],
    "sn2_grid_build": [_vp, _i, _i, _f, _vp, _vp, _vp, _vp],
    "sn2_ball_count": [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _i, _vp, _vp],
    "sn2_rowptr_scan": [_vp, _i, _i, _vp, _vp, _vp],
    "sn2_ball_fill": [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _i, _vp, _vp, _vp],
    "sn2_pointconv_fwd": [_i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _vp],
    "sn2_sa_fused_fwd": [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _i, _vp, _i, _vp, _vp, _i, _vp],
    "sn2_global_sa_fwd": [_vp, _vp, _i, _i, _vp, _i, _vp, _vp],
    "sn2_fp3_fwd": [_vp, _vp, _vp, _i, _i, _vp, _i, _vp, _vp],
    "sn2_knn3": [_vp, _vp, _i, _i, _i, _vp, _vp, _vp],
    "sn2_knn3_grid": [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _i, _vp],
    "sn2_fp2_fwd": [_vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _vp],
    "sn2_fp1_head_fwd": [_vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp],
    "sn2_fp1_head_fwd_tc": [_vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp],
    "sn2_edge_msg_fwd": [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp],
    "sn2_edge_msg_bwd": [_vp, _vp, _ll, _vp, _i, _vp, _vp],
    "sn2_segment_max_fwd": [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp],
    "sn2_segment_max_bwd": [_vp, _vp, _ll, _i, _ll, _vp, _vp],
    "sn2_interp3_fwd": [_vp, _i, _vp, _vp, _ll, _i, _vp, _vp],
    "sn2_interp3_bwd": [_vp, _vp, _vp, _ll, _i, _vp, _vp],
    "sn2_interp_plot_fwd": [_vp, _vp, _i, _i, _i, _vp, _vp],
    "sn2_interp_plot_bwd": [_vp, _vp, _i, _i, _i, _vp, _vp],
    "sn2_project_plotwise_bwd": [_vp, _vp, _i, _i, _vp, _vp],
    "sn2_linear_wgrad_supported": [_i, _i],
    "sn2_linear_wgrad": [_vp, _vp, _ll, _i, _i, _vp, _i, _vp, _vp, _vp],
    "sn2_lrb_supported": [_i, _i],
    "sn2_lrb_fwd": [_vp, _vp, _vp, _vp, _ll, _vp, _i, _i, _vp, _vp, _vp],
    "sn2_bn_finalize": [_vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _i, _vp],
    "sn2_bn_param_grad": [_vp, _vp, _i, _vp, _vp, _vp],
    "sn2_lrb_block_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _ll, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "sn2_lrb_block_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp],
    "sn2_bn_apply": [_vp, _vp, _ll, _vp, _i, _vp, _vp],
    "sn2_lrb_bwd_reduce": [_vp, _vp, _ll, _vp, _i, _vp, _vp],
    "sn2_lrb_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _vp, _i, _i, _vp, _vp, _i, _vp, _vp, _vp],
    "sn2_head_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _ll, _vp, _vp, _vp],
    "sn2_head_bwd_partials": [],
    "sn2_sa1t_partials": [],
    "sn2_sa1t_blocks": [],
    "sn2_sa1t_pre": [_vp, _vp, _ll, _vp, _vp, _vp],
    "sn2_sa1t_stats1": [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp],
    "sn2_sa1t_stats2": [_vp, _vp, _vp, _vp, _i] + [_vp] * 11,
    "sn2_sa1t_finish": [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp],
    "sn2_sa1t_bwd_sums": [_vp, _vp, _vp, _i, _vp, _vp],
    "sn2_sa1t_bwd_w2": [_vp, _vp, _vp, _vp, _i] + [_vp] * 17,
    "sn2_sa1t_bwd_in": [_vp, _vp, _vp, _vp, _ll, _i] + [_vp] * 17,
    "sn2_sa2t_pre": [_vp, _vp, _ll, _vp, _vp, _vp],
    "sn2_sa2t_fwd": [_vp, _vp, _vp, _vp, _i] + [_vp] * 8,
    "sn2_sa2t_finish": [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp],
    "sn2_sa2t_bwd_sums": [_vp, _vp, _vp, _i, _vp, _vp],
    "sn2_sa2t_bwd": [_vp, _vp, _vp, _vp, _ll, _i] + [_vp] * 12,
    "sn2_sa2t_bwd_w": [_vp, _vp, _vp, _vp, _vp, _ll, _i, _vp, _vp, _vp, _vp],
    "sn2_sa1t_bwd_w1": [_vp, _vp, _vp, _vp, _vp, _ll, _i, _vp, _vp, _vp, _vp],
    "sn2_head_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp],
    "sn2_pointwise_loss_fwd": [_vp, _vp, _ll, _vp, _vp, _vp],
    "sn2_pointwise_loss_bwd": [_vp, _vp, _vp, _ll, _vp, _vp],
    "sn2_kde_lut": [_vp, _i, _i, _i, _f, _vp, _vp, _i, _vp, _vp],
    "sn2_comm_region_bytes": [],
    "sn2_comm_max_world": [],
    "sn2_comm_max_bytes": [],
    "sn2_comm_region_alloc": [_vp],
    "sn2_comm_region_free": [_vp],
    "sn2_comm_ipc_export": [_vp, _vp],
    "sn2_comm_ipc_import": [_vp, _vp],
    "sn2_comm_ipc_release": [_vp],
    "sn2_comm_create": [_i, _i, _vp, _vp],
    "sn2_comm_destroy": [_vp],
    "sn2_comm_status": [_vp, _vp, _vp],
    "sn2_comm_allreduce_f64": [_vp, _vp, _i, _vp],
    "sn2_comm_allreduce_f32": [_vp, _vp, _i, _f, _vp],
    "sn2_bn_finalize_sync": [_vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _i, _vp],
    "sn2_bn_bwd_sync": [_vp, _vp, _vp, _i, _vp, _vp, _vp],
    "sn2_adam_step": [_vp, _vp, _f, _vp, _vp, _vp, _ll, _vp, _f, _f, _f, _f, _vp, _vp],
    "sn2_augment_rescale": [_vp, _i, _i, _vp, _vp, _vp, _f, _vp, _vp, _vp],
    "sn2_parcel_grid_build": [_vp, _vp, _vp, _ll, _f, _f, _f, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "sn2_plot_capacity": [],
    "sn2_extract_plots": [_vp, _vp, _ll, _f, _f, _f, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _f, _f, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "sn2_sample_hash": [ctypes.c_uint, ctypes.c_uint],
    "sn2_finalize_mosaic": [_vp, _i, _i, _vp, _vp, _vp, _vp],
    "sn2_fuse_accumulate": [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "sn2_fuse_finalize": [_vp, _vp, _vp, _i, _i, _vp, _vp],
    "sn2_project_plotwise": [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "sn2_project_rasters": [_vp, _vp, _ll, _ll, _ll, _i, _i, _i, _i, _f, _f, _vp, _vp, _vp],
}
_RESTYPES = {"sn2_error_string": ctypes.c_char_p, "sn2_last_cuda_error": ctypes.c_char_p,
             "sn2_comm_region_bytes": ctypes.c_size_t, "sn2_sample_hash": ctypes.c_uint}


def load(require_cuda: bool = True):
    """Load the shared library (once).  Raises RuntimeError when it cannot run the product path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"libsn2_b200.so not found at {LIB_PATH}: build it with "
                "`python stratanet2-vegetation-coverage-maps_b200/build.py` -- there is no CPU fallback"
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here = header/library mismatch
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, ctypes.c_int)
        _lib = lib
    if require_cuda and not torch.cuda.is_available():
        raise RuntimeError("sn2: no CUDA device available and this build has no CPU fallback")
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        lib = load(require_cuda=False)
        msg = lib.sn2_error_string(rc).decode()
        if rc == -3:
            msg += ": " + lib.sn2_last_cuda_error().decode()
        raise RuntimeError(f"{what} failed: {msg}")


def dptr(t: torch.Tensor | None, dtype=None):
    """Raw device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("sn2: expected a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError("sn2: expected a contiguous tensor")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"sn2: expected dtype {dtype}, got {t.dtype}")
    return ctypes.c_void_p(t.data_ptr())


def hptr(t: torch.Tensor):
    """Raw pointer of a contiguous CPU fp32 tensor."""
    assert (not t.is_cuda) and t.is_contiguous() and t.dtype == torch.float32
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    """Raw handle of torch's current stream on the current device (the private getter skips ~15 us of Python
    per call in torch.cuda.current_stream(); the training step makes ~45 such calls)."""
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))
