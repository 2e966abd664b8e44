"""Isolated timing of the fused SA1 kernel pieces (grid build, pre, search-only, full)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stratanet2-vegetation-coverage-maps_b200"))
import numpy as np, torch
from sn2 import ops, _lib, weights
from sn2.synth import synth_batch
from bench import make_model

def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

B, N = 64, 16384
dev = torch.device("cuda")
args, net = make_model(N, 0)
W = weights.pack_eval(net)
data = synth_batch(2, B, N)
pos0, feat0 = ops.ingest(data["xyz"].to(dev), data["cloud"].to(dev))
M1 = ops.m_of(N, 0.25)
_, pos1 = ops.fps_dense(pos0, B, N, M1)
r = float(np.sqrt(2.0))
print("grid build x2 :", t(lambda: (ops.build_grid(pos0, B, N, r), ops.build_grid(pos1, B, M1, r))))
grid0 = ops.build_grid(pos0, B, N, r)
for name, tc in (("SIMT fp32", 0), ("tcgen05 3xTF32", 1), ("tcgen05 TF32", 2), ("tcgen05 BF16-precision", 3)):
    print(f"sa1 fused, {name:24s}:", t(lambda: ops.sa_fused_fwd(1, pos0, feat0, pos1, B, N, M1, r, 2000, W["sa1"], tensor_core=tc, grid=grid0)))
hdr, cs, sorted4 = ops.build_grid(pos0, B, N, r)
_, _, qsorted4 = ops.build_grid(pos1, B, M1, r)
lib = ctypes.CDLL(_lib.LIB_PATH)
vp, i, f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
lib.sn2_debug_sa_search_only.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, f, i, vp, i, vp, vp, vp]
u = torch.zeros((B * N, 16), device=dev); ovf = torch.zeros(B * M1 + 1, dtype=torch.int32, device=dev)
out = torch.empty((B * M1, 16), device=dev); cnt = torch.empty(B * M1, dtype=torch.int32, device=dev)
w = W["sa1"]
def search_only():
    lib.sn2_debug_sa_search_only(vp(hdr.data_ptr()), vp(cs.data_ptr()), vp(sorted4.data_ptr()), vp(qsorted4.data_ptr()), vp(u.data_ptr()),
                                 vp(ovf.data_ptr()), B, N, M1, ops.r2_of(r), 2000, vp(w.data_ptr()), w.numel(), vp(out.data_ptr()), vp(cnt.data_ptr()), None)
print("search only   :", t(search_only))
c = cnt.float()
print("hits/centroid mean %.1f  median %.0f  p10 %.0f  p90 %.0f  max %.0f   total edges %.2fM" % (c.mean(), c.median(), c.quantile(0.1), c.quantile(0.9), c.max(), c.sum() / 1e6))
print("passes of 32 per centroid: mean %.2f (edges/32 = %.2f)" % (torch.ceil(c / 32).mean(), (c / 32).mean()))
g = hdr.view(B, 12)[0]
print("grid gx,gy,gz:", g[4].view(torch.int32).item(), g[5].view(torch.int32).item(), g[8].view(torch.int32).item(), "cs", g[3].item(), "csz", g[9].item())
zc = pos1.view(B, M1, 4)[0, :, 2]
print("centroid z: frac > 1.5 m = %.2f, frac > 0.5 = %.2f" % ((zc > 1.5).float().mean(), (zc > 0.5).float().mean()))
