"""Where does the host block in the data-parallel training loop?  Run under torchrun (or alone):
per-phase host time of the StructurePrefetcher loop vs the plain loop, rank 0 prints."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stratanet2-vegetation-coverage-maps_b200"))
import torch, torch.distributed as dist
from bench import make_model, train_loss
from model.project_to_2d import project_to_plotwise_coverages
from sn2 import parallel
from sn2.pipeline import StructurePrefetcher, TrainStructure
from sn2.synth import synth_batch

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1: dist.init_process_group("nccl", device_id=dev)
Bg, N = 32, 10000
args, net = make_model(N, local); net.train()
if world > 1: net = parallel.convert_sync_batchnorm(net)
bucket = parallel.GradBucket(net)
opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-3)
full = synth_batch(3, Bg, N); full["gt"] = torch.rand(Bg, 4)
mine = parallel.shard_plots(full, rank, world); Bl = mine["cloud"].shape[0]
d = {k: v.to(dev) for k, v in mine.items()}
z = d["xyz"][:, 2, :].reshape(-1, 1).double()
pdf = torch.cat([torch.exp(-z), 0.5 * torch.exp(-0.5 * (z - 1.0) ** 2), 0.1 + 0.05 * z], dim=1)

def step(b):
    bucket.zero()
    cov, proba = net({k: b[k] for k in ("xyz", "cloud", "sn2_structure") if k in b})
    pw = project_to_plotwise_coverages(cov, net.last_cloud_device, args)
    loss = train_loss(proba, pw, d["gt"], pdf)
    loss.backward()
    bucket.allreduce(Bl, Bg)
    opt.step()

def sync():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()

for _ in range(5): step(d)
sync(); t0 = time.perf_counter()
for _ in range(10): step(d)
sync(); t1 = time.perf_counter()
if rank == 0: print(f"plain loop      {1e3 * (t1 - t0) / 10:.2f} ms/step")

import faulthandler
pf = StructurePrefetcher(net, None, dev)
def loop(n, log):
    cur_s = pf._finish(pf._begin(d))
    for i in range(n):
        if log: faulthandler.dump_traceback_later(0.06, exit=False)   # a step that blocks > 60 ms shows where
        a = time.perf_counter(); nxt = pf._begin(d)
        b = time.perf_counter(); step({**d, "sn2_structure": cur_s})
        c = time.perf_counter(); pf._retire(cur_s); cur_s = pf._finish(nxt)
        e = time.perf_counter()
        if log: faulthandler.cancel_dump_traceback_later()
        if log and rank == 0: print(f"  step {i}: begin {1e3 * (b - a):.2f}  step {1e3 * (c - b):.2f}  finish {1e3 * (e - c):.2f} ms (host)")
loop(5, False)
seg = lambda: torch.cuda.memory_stats()["segment.all.allocated"]
sync(); s0 = seg(); t0 = time.perf_counter()
loop(30, True)
sync(); t1 = time.perf_counter()
if rank == 0: print(f"prefetch loop   {1e3 * (t1 - t0) / 30:.2f} ms/step   cudaMalloc segments during the loop: {seg() - s0}")
if "--trace" in sys.argv:
    # device timeline of two prefetch-loop steps on rank 0: (start, duration, stream, kernel) from the profiler's trace
    import json, tempfile
    from torch.profiler import ProfilerActivity, profile
    sync()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        loop(3, False)
        sync()
    if rank == 0:
        path = os.path.join(tempfile.gettempdir(), "sn2_trace.json")
        prof.export_chrome_trace(path)
        ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
        ev.sort(key=lambda e: e["ts"])
        t0 = ev[0]["ts"]
        streams = sorted({e["args"].get("stream") for e in ev})
        print("streams:", streams)
        for e in ev:
            if e["dur"] >= 15 or "nccl" in e["name"].lower():
                print(f"{(e['ts'] - t0) / 1e3:9.3f} ms  +{e['dur'] / 1e3:7.3f}  s{streams.index(e['args'].get('stream'))}  {e['name'][:70]}")
    sync()
# variant: structural stage on the main stream but split begin / finish around nothing (isolates the side stream)
pf.side = torch.cuda.current_stream(dev)
loop(5, False)
sync(); t0 = time.perf_counter()
loop(10, False)
sync(); t1 = time.perf_counter()
if rank == 0: print(f"same-stream split loop {1e3 * (t1 - t0) / 10:.2f} ms/step")
if world > 1: dist.destroy_process_group()
