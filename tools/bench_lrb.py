"""Isolated timing of the fused train-mode block kernels (csrc/train_mlp.cu) at config-3 shapes, with the
algorithmic HBM traffic of each kernel (fp32 rows read + written)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stratanet2-vegetation-coverage-maps_b200"))
import torch
from sn2 import _lib
from sn2._lib import dptr
from sn2.autograd_ops import LinReluBN


def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


lib = _lib.load()
dev = torch.device("cuda")
st = None
for Ci, Co, R, need_dx in ((11, 16, 4_000_000, False), (16, 16, 4_000_000, True), (19, 32, 900_000, True), (80, 34, 80_000, True),
                           (42, 34, 320_000, True)):
    x = torch.randn(R, Ci, device=dev); W = torch.randn(Co, Ci, device=dev) * 0.3; b = torch.randn(Co, device=dev)
    gamma = torch.rand(Co, device=dev) + 0.5; beta = torch.randn(Co, device=dev)
    y = torch.empty(R, Co, device=dev); z = torch.empty(R, Co, device=dev); dz = torch.randn(R, Co, device=dev)
    stats = torch.empty(2 * Co + 1, dtype=torch.float64, device=dev); ss = torch.empty(4 * Co, device=dev)
    sums = torch.empty(2 * Co, dtype=torch.float64, device=dev)
    dx = torch.empty_like(x) if need_dx else None
    nblk = LinReluBN.NBLK
    partial = torch.empty(nblk, Co * (Ci + 1), device=dev); dW = torch.empty_like(W); db = torch.empty(Co, device=dev)
    fwd = lambda: lib.sn2_lrb_fwd(dptr(x), None, dptr(W), dptr(b), R, None, Co, Ci, dptr(y), dptr(stats), st)
    fin = lambda: lib.sn2_bn_finalize(dptr(stats), dptr(gamma), dptr(beta), 1e-5, 0.1, None, None, None, dptr(ss), Co, st)
    app = lambda: lib.sn2_bn_apply(dptr(y), dptr(ss), R, None, Co, dptr(z), st)
    red = lambda: lib.sn2_lrb_bwd_reduce(dptr(dz), dptr(y), R, None, Co, dptr(sums), st)
    bwd = lambda: lib.sn2_lrb_bwd(dptr(dz), dptr(y), dptr(x), None, dptr(W), dptr(ss), dptr(sums), dptr(stats), R, None, Co, Ci, dptr(dx),
                                  dptr(partial), nblk, dptr(dW), dptr(db), st)
    assert fwd() == 0 and fin() == 0 and app() == 0 and red() == 0 and bwd() == 0
    rows = (("lrb_fwd", fwd, Ci + Co), ("bn_apply", app, 2 * Co), ("bwd_reduce", red, 2 * Co),
            ("lrb_bwd", bwd, 2 * Co + Ci + (Ci if need_dx else 0)))
    print(f"--- Ci={Ci} Co={Co} R={R} dx={need_dx}")
    for name, fn, floats in rows:
        ms = t(fn)
        print(f"  {name:11s} {ms:7.4f} ms  {R * floats * 4 / ms / 1e6:7.0f} GB/s")
