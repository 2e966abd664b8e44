"""Per-kernel device time of the round-2 config-3 training step (graphed + prefetched loop, torch.profiler CUDA activity):
prints a markdown table (ms per step, launches per step, library vs ours) for profiles/."""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stratanet2-vegetation-coverage-maps_b200")]
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
from model.project_to_2d import project_to_plotwise_coverages  # noqa: E402
from sn2 import losses  # noqa: E402
from sn2.optim import FusedAdam  # noqa: E402
from sn2.pipeline import GraphedTrainStep, StructurePrefetcher  # noqa: E402
from sn2.synth import synth_batch  # noqa: E402

B, N, STEPS = 32, 10000, 8
dev = torch.device("cuda", 0)
args, net = bench.make_model(N, 0)
net.train()
opt = FusedAdam(net.parameters(), lr=args.lr, weight_decay=args.wd)
lut = bench.synthetic_kde_lut(dev)
d = {k: v.to(dev) for k, v in synth_batch(3, B, N).items()}
d["gt"] = torch.rand(B, 4, device=dev)


def step_fn(batch):
    opt.zero_grad()
    cov, proba = net(batch)
    pw = project_to_plotwise_coverages(cov, net.last_cloud_device, args)
    loss = losses.training_loss(pw, batch["gt"], proba, lut.pdf(net.last_cloud_device, args.z_max), args.m, args.e)[0]
    loss.backward()
    opt.step()
    return loss.detach()


gs = GraphedTrainStep(net, step_fn, opt, device=dev)


def loop(n):
    for b in StructurePrefetcher(net, (d for _ in range(n)), dev):
        gs(b)


loop(5)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    loop(STEPS)
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "sn2_trace_r2.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
span = (ev[-1]["ts"] + ev[-1]["dur"] - ev[0]["ts"]) / 1e3 / STEPS
agg = {}
for e in ev:
    k = e["name"][:70]
    a = agg.setdefault(k, [0.0, 0])
    a[0] += e["dur"] / 1e3 / STEPS
    a[1] += 1
ours = lambda n: n.startswith("sn2::") or n.startswith("void sn2::")  # noqa: E731
tot_ours = sum(v[0] for k, v in agg.items() if ours(k))
tot_lib = sum(v[0] for k, v in agg.items() if not ours(k))
print(f"config 3 (32 plots x 10 000 pts), graphed + prefetched loop: {span:.3f} ms / step wall (trace span); "
      f"kernels of libsn2_b200 {tot_ours:.3f} ms / step, everything else (torch elementwise / cat / copies / memsets) {tot_lib:.3f} ms / step "
      f"(streams overlap: the sums exceed the wall time)")
print("\n| kernel | ms / step | launches / step | ours |\n|---|---:|---:|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"| `{k}` | {v[0]:.4f} | {v[1] / STEPS:.1f} | {'yes' if ours(k) else 'no'} |")
big_lib = [(k, v[0]) for k, v in agg.items() if not ours(k) and v[0] > 0.02]
print("\nnon-sn2 kernels above 0.02 ms / step:", [(k[:50], round(t, 3)) for k, t in sorted(big_lib, key=lambda kv: -kv[1])])
