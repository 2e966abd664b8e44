"""Isolated timing of the FP1 + head kernel: SIMT fp32 vs tcgen05 3xTF32 (config 2 shapes: 64 x 16384 rows)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stratanet2-vegetation-coverage-maps_b200"))
import torch
from sn2 import ops, weights
from bench import make_model


def t(fn, n=20):
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
M1 = ops.m_of(N, 0.25)
Q = B * N
g = torch.Generator(device=dev).manual_seed(0)
f2 = torch.randn((B * M1, 36), device=dev, generator=g)
feat = torch.randn((Q, 8), device=dev, generator=g)
# neighbours: 3 random sources inside the row's own plot (locality of a real 3-NN list is better than this)
plot = torch.arange(Q, device=dev) // N
nbr = (torch.randint(0, M1, (Q, 3), device=dev, generator=g) + plot[:, None] * M1).int().contiguous()
w = torch.rand((Q, 3), device=dev, generator=g) + 0.1
simt = lambda: ops.fp1_head_fwd(f2, nbr, w, feat, W["fp1"])
tc = lambda: ops.fp1_head_fwd(f2, nbr, w, feat, W["fp1"], tensor_core=True)
c0, p0 = simt(); c1, p1 = tc()
print("max |cov_tc - cov_simt| = %.3g, max |proba| diff = %.3g" % ((c0 - c1).abs().max(), (p0 - p1).abs().max()))
bytes_alg = Q * (3 * 4 + 3 * 4 + 8 * 4 + 2 * 16) + f2.numel() * 4
for name, fn in (("simt fp32", simt), ("tcgen05 3xTF32", tc)):
    ms = t(fn)
    print("%-16s %.4f ms   %.0f GB/s algorithmic" % (name, ms, bytes_alg / ms / 1e6))
