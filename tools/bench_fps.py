"""Time the FPS kernels in isolation (CUDA events), per algorithm, for the two levels of a config."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stratanet2-vegetation-coverage-maps_b200"))
import torch
from sn2 import ops
from sn2.synth import synth_batch

def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

for B, N in ((64, 16384), (32, 10000), (1, 10000), (8, 8192), (64, 4096)):
    data = synth_batch(2, B, N)
    dev = torch.device("cuda")
    pos0, _ = ops.ingest(data["xyz"].to(dev), data["cloud"].to(dev))
    M = ops.m_of(N, 0.25)
    ref = None
    for name, algo in (("brute", 1), ("bucket", 2), ("ilp2", 5), ("nw16", 6), ("nw16+ilp2", 7)):
        ms = t(lambda: ops.fps_dense(pos0, B, N, M, None, algo))
        idx, _ = ops.fps_dense(pos0, B, N, M, None, algo)
        same = True if ref is None else bool(torch.equal(idx, ref))
        ref = idx if ref is None else ref
        print(f"B={B} N={N} M={M} {name:9s} {ms:8.3f} ms  {ms*1e6/M:8.1f} ns/iter  same={same}")
