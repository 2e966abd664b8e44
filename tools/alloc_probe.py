import sys, time, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/stratanet2-vegetation-coverage-maps_b200")
from bench import make_model
from sn2 import ops
from sn2.pipeline import forward_eval, StageTimer
from sn2.synth import synth_batch
dev = torch.device("cuda", 0)
for B, N in ((1, 10000), (64, 16384)):
    args, net = make_model(N, 0)
    d = {k: v.to(dev) for k, v in synth_batch(1, B, N).items()}
    def step(timer=None):
        with torch.no_grad():
            cov, proba, g, cl = forward_eval(net, d["xyz"], d["cloud"], dev, 2000, None, timer)
            ops.project_plotwise(cl, cov, 20)
    for _ in range(5): step()
    torch.cuda.synchronize()
    for label, use_timer in (("plain", False), ("timer", True)):
        s0 = torch.cuda.memory_stats()
        t0 = time.perf_counter()
        for _ in range(20):
            step(StageTimer() if use_timer else None)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        s1 = torch.cuda.memory_stats()
        print(f"B={B} {label}: cpu enqueue {1e3*(t1-t0)/20:.2f} ms/step, total {1e3*(t2-t0)/20:.2f} ms/step, cudaMalloc calls {s1['num_device_alloc']-s0['num_device_alloc']}, segments {s1['segment.all.current']}, reserved {s1['reserved_bytes.all.current']/1e6:.0f} MB")
