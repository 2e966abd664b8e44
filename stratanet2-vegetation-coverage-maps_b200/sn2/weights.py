"""Packing of PointNet2 parameters into the flat fp32 blobs the kernels take by value.

Layout per MLP block (csrc/mlp.cu ``Layer<CIN,COUT>``): w[k][o] (= Linear.weight transposed, k-major),
bias, s, t with the eval-mode BatchNorm folded to s = gamma / sqrt(running_var + eps),
t = beta - running_mean * s (BN comes AFTER ReLU: /root/reference/model/point_net2.py:45-53).
Plain Linear (csrc ``Lin``): w[k][o], bias.
"""
from __future__ import annotations

import torch


def _layer(block) -> list[torch.Tensor]:
    lin, bn = block[0], block[2]
    s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    t = bn.bias - bn.running_mean * s
    return [lin.weight.t().reshape(-1), lin.bias, s, t]


def _lin(lin) -> list[torch.Tensor]:
    return [lin.weight.t().reshape(-1), lin.bias]


def pack_eval(model) -> dict[str, torch.Tensor]:
    """-> {stage: contiguous CPU fp32 blob}.  One device->host copy for the whole model."""
    with torch.no_grad():
        parts = {
            "sa1": _layer(model.sa1_module.conv.local_nn[0]) + _layer(model.sa1_module.conv.local_nn[1]),
            "sa2": _layer(model.sa2_module.conv.local_nn[0]),
            "sa3": _layer(model.sa3_module.nn[0]),
            "fp3": _layer(model.fp3_module.nn[0]),
            "fp2": _layer(model.fp2_module.nn[0]),
            "fp1": _layer(model.fp1_module.nn[0]) + _lin(model.lin1) + _lin(model.lin2),
        }
        sizes = {k: sum(p.numel() for p in v) for k, v in parts.items()}
        flat = torch.cat([p.reshape(-1).to(torch.float32) for v in parts.values() for p in v]).cpu()
    out, off = {}, 0
    for k, n in sizes.items():
        out[k] = flat[off:off + n].contiguous()
        off += n
    return out


EXPECTED_SIZES = {"sa1": 528, "sa2": 704, "sa3": 2432, "fp3": 6336, "fp2": 2822, "fp1": 2175}


def params_version(model) -> tuple:
    """Cheap change detector for the packed-weights cache."""
    return tuple((p.data_ptr(), p._version) for p in list(model.parameters()) + list(model.buffers()))
