"""BASELINE config 5: dense-cloud stress sweep.  Per-stage CUDA-event times of the eval forward + projections
vs points per plot, max_num_neighbors = 64, SIMT fp32 and tcgen05 (3xTF32) second SA1 layer.  Prints markdown."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stratanet2-vegetation-coverage-maps_b200"))
import torch
from bench import make_model
from sn2 import ops
from sn2.pipeline import StageTimer, forward_eval
from sn2.synth import synth_batch

dev = torch.device("cuda", 0)
rows = []
for N, B in ((4096, 64), (8192, 64), (16384, 64), (32768, 32), (65536, 32)):
    for K in (64, 2000):
        for tc in (0, 1):
            if tc and K == 2000:
                continue
            args, net = make_model(N, 0)
            net.sa1_module.max_num_neighbors = K
            net.sn2_tensor_core = tc
            data = {k: v.to(dev) for k, v in synth_batch(5, B, N).items()}
            def step(timer=None):
                cov, proba, g, cloud_d = forward_eval(net, data["xyz"], data["cloud"], dev, K, None, timer)
                ops.project_plotwise(cloud_d, cov, args.diam_pix)
                ops.project_rasters(cloud_d, cov, "point_major", args.diam_pix, args.diam_meters)
            with torch.no_grad():
                for _ in range(3): step()
                torch.cuda.synchronize()
                timer = StageTimer()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n = 5
                a.record()
                for _ in range(n): step(timer)
                b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b) / n
            st = {k: v / n for k, v in timer.totals_ms().items()}
            rows.append((N, B, K, tc, ms, B / ms * 1e3, st))
            print(f"N={N} B={B} K={K} tc={tc}: {ms:.2f} ms/step  {B / ms * 1e3:.0f} plots/s", file=sys.stderr)
keys = ["fps1", "sa1_fused", "fps2", "sa2_fused", "knn1", "knn2", "fp1_head", "fp2", "fp3", "global_sa", "ingest"]
print("| points/plot | plots | cap K | SA1 layer 2 | ms/step | plots/s | points/s | " + " | ".join(keys) + " |")
print("|---:|---:|---:|---|---:|---:|---:|" + "---:|" * len(keys))
for N, B, K, tc, ms, pps, st in rows:
    print(f"| {N} | {B} | {K} | {'tcgen05 3xTF32' if tc else 'SIMT fp32'} | {ms:.2f} | {pps:.0f} | {pps * N / 1e6:.0f} M | " +
          " | ".join(f"{st.get(k, 0):.3f}" for k in keys) + " |")
