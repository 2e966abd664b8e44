"""torch.profiler summary of one config-3 training step (where the time goes outside libsn2 kernels)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stratanet2-vegetation-coverage-maps_b200"))
import torch
from torch.profiler import ProfilerActivity, profile
from bench import make_model, train_loss
from model.project_to_2d import project_to_plotwise_coverages
from sn2 import parallel
from sn2.synth import synth_batch

B, N = 32, 10000
dev = torch.device("cuda", 0)
args, net = make_model(N, 0)
net.train()
bucket = parallel.GradBucket(net)
opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-3)
d = synth_batch(3, B, N)
d = {k: v.to(dev) for k, v in d.items()}
gt = torch.rand(B, 4, device=dev)
def step():
    bucket.zero()
    cov, proba = net(d)
    pw = project_to_plotwise_coverages(cov, net.last_cloud_device, args)
    z = d["xyz"][:, 2, :].reshape(-1, 1).double()
    pdf = torch.cat([torch.exp(-z), 0.5 * torch.exp(-0.5 * (z - 1.0) ** 2), 0.1 + 0.05 * z], dim=1)
    loss = train_loss(proba, pw, gt, pdf)
    loss.backward()
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=60))

if "--cprofile" in sys.argv:
    # host-side cost of a step (the step is launch bound once the kernels are fused): Python profile of 20 steps
    import cProfile, pstats, time
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"host enqueue time {1e3 * (t1 - t0) / 20:.2f} ms/step, drained after {1e3 * (t2 - t1):.2f} ms more")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(20): step()
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("tottime").print_stats(45)
