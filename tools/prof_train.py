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

if "--graph" in sys.argv:
    # device timeline of the graphed + prefetched loop: busy time per stream, biggest gaps, top kernels
    import json, tempfile
    from sn2.pipeline import GraphedTrainStep, StructurePrefetcher
    for p in net.parameters():
        p.grad = None
    bucket = parallel.GradBucket(net)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-3, capturable=True)
    pdf_s = None
    def gfn(batch):
        bucket.zero()
        cov, proba = net(batch)
        pw = project_to_plotwise_coverages(cov, net.last_cloud_device, args)
        z = batch["xyz"][:, 2, :].reshape(-1, 1).double()
        pdf = torch.cat([torch.exp(-z), 0.5 * torch.exp(-0.5 * (z - 1.0) ** 2), 0.1 + 0.05 * z], dim=1)
        loss = train_loss(proba, pw, batch["gt"], pdf)
        loss.backward()
        opt.step()
        return loss.detach()
    gs = GraphedTrainStep(net, gfn, optimizer=opt, device=dev)
    src = {"xyz": d["xyz"], "cloud": d["cloud"], "gt": gt}
    def loop(n):
        for b in StructurePrefetcher(net, (src for _ in range(n)), dev):
            gs(b)
    loop(5); torch.cuda.synchronize()
    import time
    pf = StructurePrefetcher(net, None, dev)
    cur_s = pf._finish(pf._begin(src))
    acc = [0.0] * 6
    for i in range(20):
        t = [time.perf_counter()]
        nxt = pf._begin(src); t.append(time.perf_counter())
        S = cur_s; cur = torch.cuda.current_stream(dev)
        cur.wait_event(S.done); t.append(time.perf_counter())
        gs.static_struct.load(S); gs._load_batch(src, S); t.append(time.perf_counter())
        gs.graph.replay(); t.append(time.perf_counter())
        pf._retire(cur_s); t.append(time.perf_counter())
        cur_s = pf._finish(nxt); t.append(time.perf_counter())
        for j in range(6): acc[j] += (t[j + 1] - t[j]) * 1e3 / 20
    torch.cuda.synchronize()
    print("host ms/step: begin %.3f  wait_event %.3f  static loads %.3f  replay %.3f  retire %.3f  finish %.3f" % tuple(acc))
    t0 = time.perf_counter(); loop(20); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"graphed loop: host {1e3 * (t1 - t0) / 20:.2f} ms/step, total {1e3 * (t2 - t0) / 20:.2f} ms/step")
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        loop(6); torch.cuda.synchronize()
    path = os.path.join(tempfile.gettempdir(), "sn2_trace_graph.json")
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    ev.sort(key=lambda e: e["ts"])
    streams = sorted({e["args"].get("stream") for e in ev})
    span = (ev[-1]["ts"] + ev[-1]["dur"] - ev[0]["ts"]) / 1e3
    print(f"trace span {span:.2f} ms for 6 steps; streams {streams}")
    for sid in streams:
        es = [e for e in ev if e["args"].get("stream") == sid]
        busy = sum(e["dur"] for e in es) / 1e3
        gaps = sorted(((b["ts"] - (a["ts"] + a["dur"])) / 1e3, a["name"][:40], b["name"][:40]) for a, b in zip(es, es[1:]))[-6:]
        print(f"stream {sid}: {len(es)} ops, busy {busy:.2f} ms; largest gaps (ms, after, before):")
        for g_ in reversed(gaps): print(f"     {g_[0]:.3f}  {g_[1]}  ->  {g_[2]}")
    agg = {}
    for e in ev: agg[e["name"][:60]] = agg.get(e["name"][:60], 0) + e["dur"] / 1e3 / 6
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:28]: print(f"  {v:7.3f} ms/step  {k}")
