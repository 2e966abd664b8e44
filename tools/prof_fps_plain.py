"""Per-phase cycle counts of the plain (one sample per round) bucketed FPS kernel, clock64 inside the kernel
(sn2_debug_fps_profile): mean over the warps of plot 0 and the slowest warp, per SAMPLE."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stratanet2-vegetation-coverage-maps_b200")]
import torch  # noqa: E402

from sn2 import _lib, ops  # noqa: E402
from sn2.synth import synth_batch  # noqa: E402

lib = ctypes.CDLL(_lib.LIB_PATH)
vp, i = ctypes.c_void_p, ctypes.c_int
lib.sn2_debug_fps_profile.argtypes = [vp, i, i, i, vp, vp, i, vp]
dev = torch.device("cuda")
names = ["box test + ballot", "bucket updates", "warp arg-max", "slot write + barrier", "block arg-max + coords"]
for B, N in ((8, 16384), (8, 10000)):
    data = synth_batch(2, B, N)
    pos0, _ = ops.ingest(data["xyz"].to(dev), data["cloud"].to(dev))
    M = ops.m_of(N, 0.25)
    for code, label, nw in ((8, "8 warps x 32 slots", 8), (82, "8 warps, two-bucket update", 8), (16, "16 warps x 16 slots", 16), (162, "16 warps, two-bucket update", 16)):
        if nw == 16 and N > 16384:
            continue
        idx = torch.empty(B * M, dtype=torch.int32, device=dev)
        prof = torch.zeros(B * nw * 8, dtype=torch.int64, device=dev)
        rc = lib.sn2_debug_fps_profile(vp(pos0.data_ptr()), B, N, M, vp(idx.data_ptr()), vp(prof.data_ptr()), code, None)
        torch.cuda.synchronize()
        p = prof.view(B, nw, 8).double().cpu() / (M - 1)
        print(f"N={N} M={M} {label} (rc={rc}): cycles per sample, plot 0")
        for k, nme in enumerate(names):
            print(f"   {nme:24s} mean {p[0, :, k].mean():7.1f}   slowest warp {p[0, :, k].max():7.1f}")
        print(f"   {'total':24s} {p[0, 0, :5].sum().item():7.1f}     active buckets per sample: {p[0, :, 6].sum().item():.2f} (max per warp {p[0, :, 6].max().item():.2f})")
