import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stratanet2-vegetation-coverage-maps_b200"))
import torch
from sn2 import ops, _lib
from sn2.synth import synth_batch
lib = ctypes.CDLL(_lib.LIB_PATH)
vp, i = ctypes.c_void_p, ctypes.c_int
lib.sn2_debug_fps_profile.argtypes = [vp, i, i, i, vp, vp, i, vp]
dev = torch.device("cuda")
for B, N in ((8, 16384), (8, 4096)):
    data = synth_batch(2, B, N)
    pos0, _ = ops.ingest(data["xyz"].to(dev), data["cloud"].to(dev))
    M = ops.m_of(N, 0.25)
    for nw in (8,):
        idx = torch.empty(B * M, dtype=torch.int32, device=dev)
        prof = torch.zeros(B * nw * 8, dtype=torch.int64, device=dev)
        rc = lib.sn2_debug_fps_profile(vp(pos0.data_ptr()), B, N, M, vp(idx.data_ptr()), vp(prof.data_ptr()), nw, None)
        torch.cuda.synchronize()
        raw = prof.view(B, nw, 8).double().cpu()
        rounds = raw[0, 0, 5].item()
        print(f"rounds {rounds:.0f} for {M - 1} samples -> {(M - 1) / rounds:.2f} accepted per round")
        p = raw / rounds
        names = ["test+ballot", "updates", "warp top-4", "barrier", "global top-4", "rounds", "active/warp", "-"]
        print(f"N={N} nw={nw} rc={rc}: cycles per ROUND (mean / min / max over the warps of plot 0)")
        for k, nme in enumerate(names):
            if nme != "-":
                print(f"   {nme:14s} mean {p[0,:,k].mean():8.1f}   min {p[0,:,k].min():8.1f}   max {p[0,:,k].max():8.1f}")
        print("   total", p[0, 0, :5].sum().item())
