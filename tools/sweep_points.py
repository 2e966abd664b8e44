"""BASELINE config 5: dense-cloud stress sweep.  Per-stage CUDA-event times of the eval forward + projections vs points per
plot, max_num_neighbors = 64 (the cap binds: exact first-K path) and 2000, second SA1 layer as SIMT fp32 or on tcgen05
(3xTF32 / TF32 / BF16-precision operands).  Median of 9 timed steps after 6 warm-up steps (r1's single outlier row -- knn1
3.5 ms at 8192 points -- was an allocator stall in the first timed step of a configuration: a mean over 5 steps kept it).
Prints markdown."""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stratanet2-vegetation-coverage-maps_b200")]
import torch  # noqa: E402

from bench import make_model  # noqa: E402
from sn2 import ops  # noqa: E402
from sn2.pipeline import StageTimer, forward_eval  # noqa: E402
from sn2.synth import synth_batch  # noqa: E402

dev = torch.device("cuda", 0)
MODES = {0: "SIMT fp32", 1: "tcgen05 3xTF32", 2: "tcgen05 TF32", 3: "tcgen05 BF16 operands"}
sizes = ((4096, 64), (8192, 64), (16384, 64), (32768, 32), (65536, 32))
if len(sys.argv) > 1:
    sizes = tuple((int(a), 64 if int(a) <= 16384 else 32) for a in sys.argv[1:])
rows = []
for N, B in sizes:
    data = {k: v.to(dev) for k, v in synth_batch(5, B, N).items()}
    ref = {}
    for K in (64, 2000):
        for tc in (0, 1, 2, 3):
            args, net = make_model(N, 0)
            net.sa1_module.max_num_neighbors = K
            net.sn2_tensor_core = tc

            def step(timer=None):
                cov, proba, g, cloud_d = forward_eval(net, data["xyz"], data["cloud"], dev, K, None, timer)
                ops.project_plotwise(cloud_d, cov, args.diam_pix)
                ops.project_rasters(cloud_d, cov, "point_major", args.diam_pix, args.diam_meters)
                return cov

            with torch.no_grad():
                for _ in range(6):
                    cov = step()
                torch.cuda.synchronize()
                if tc == 0:
                    ref[K] = cov.clone()
                err = float((cov - ref[K]).abs().max())
                per, stages = [], []
                for _ in range(9):
                    timer = StageTimer()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    step(timer)
                    b.record()
                    torch.cuda.synchronize()
                    per.append(a.elapsed_time(b))
                    stages.append(timer.totals_ms())
            ms = statistics.median(per)
            st = {k: statistics.median(s[k] for s in stages) for k in stages[0]}
            rows.append((N, B, K, tc, ms, B / ms * 1e3, st, err))
            print(f"N={N} B={B} K={K} {MODES[tc]}: {ms:.2f} ms/step  {B / ms * 1e3:.0f} plots/s  max |cov - fp32| {err:.2e}", file=sys.stderr)
keys = ["fps1", "sa1_fused", "fps2", "sa2_fused", "knn1", "knn2", "fp1_head", "fp2", "fp3", "global_sa", "ingest"]
print("| points/plot | plots | cap K | SA1 layer 2 | ms/step | plots/s | points/s | max abs diff of coverages vs fp32 | " + " | ".join(keys) + " |")
print("|---:|---:|---:|---|---:|---:|---:|---:|" + "---:|" * len(keys))
for N, B, K, tc, ms, pps, st, err in rows:
    print(f"| {N} | {B} | {K} | {MODES[tc]} | {ms:.2f} | {pps:.0f} | {pps * N / 1e6:.0f} M | {err:.1e} | " +
          " | ".join(f"{st.get(k, 0):.3f}" for k in keys) + " |")
