#!/usr/bin/env python
"""bench.py -- throughput of the PointNet2 + 2D-projection hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2]

One "step" = one pass of the hot path over one batch of synthetic plots: PointNet2
eval forward + project_to_plotwise_coverages + project_to_2d_rasters.  Default workload is
BASELINE.json configs[1]: 64 plots x 16 384 points, fp32, one B200.  With N > 1 (torchrun, one process
per GPU) every rank runs its own 64 plots (weak scaling, no data-path collective: plots are
independent); value = plots of all ranks / max-over-ranks time.

Prints ONE JSON line (rank 0).  ``value``: inputs resident in HBM.  ``e2e``: the same step through the
public drop-in API (``model.point_net2.PointNet2.forward`` + ``model.project_to_2d``) from pinned HOST
buffers with the result read back to the host inside the timed region.  ``roofline``: the dominant
kernel, timed live with CUDA events in the same timed region.  ``cpu_baseline``: the CPU oracle (the
reference's own model files on restated third-party ops when staged, else the port) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "stratanet2-vegetation-coverage-maps_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

CONFIGS = {
    1: dict(B=1, N=10000, mode="infer", name="config1: 1 plot x 10k pts, eval fwd + both projections"),
    2: dict(B=64, N=16384, mode="infer",
            name="config2: batched inference 64 plots x 16384 pts, fp32, eval fwd + both projections"),
    4: dict(B=64, N=10000, mode="parcel",
            name="config4: parcel-scale inference, 1 km^2 synthetic parcel + 20 m buffer (34.6 M points) tiled on the GPU into the "
                 "reference's overlapping 10 m plots x 10k pts, plot-sharded across GPUs, local-map fusion + one all-reduce + band finalisation"),
    3: dict(B=32, N=10000, mode="train",
            name="config3: training step (fwd + plot-wise projection + loss + bwd + grad all-reduce + Adam), global batch 32 plots x 10k pts"),
}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class no_gc:
    """Timed regions run with the cyclic garbage collector off, as `timeit` does: a full collection in a process
    that has torch imported is a 50-200 ms host pause, which lands in the middle of a step (seen as single steps of
    40-180 ms in otherwise 4.1 ms serial passes).  SN2_BENCH_GC=1 leaves it on."""

    def __enter__(self):
        import gc
        self.was = gc.isenabled()
        if os.environ.get("SN2_BENCH_GC", "0") != "1":
            gc.collect()
            gc.disable()
        return self

    def __exit__(self, *exc):
        import gc
        if self.was:
            gc.enable()
        return False


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        if os.environ.get("SN2_NO_CLOCK_SAMPLER") == "1":  # debugging aid: is the sampler itself perturbing the run?
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def wait_ready(self, timeout: float = 8.0):
        """Block until nvidia-smi has produced its first sample.  Its start-up (NVML / driver initialisation, a few
        hundred ms) perturbs kernel submission: a timed pass that began right after start() came out 1.3-3x slower
        in about half of the runs.  Everything timed starts after this returns."""
        t0 = time.perf_counter()
        while self.proc is not None and not self.lines and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_model(N, device_index):
    from model.point_net2 import PointNet2
    from sn2.config import default_args
    from sn2.synth import randomize_bn_

    args = default_args(subsample_size=N, cuda=device_index)
    torch.manual_seed(0)
    net = PointNet2(args)
    randomize_bn_(net)
    net.eval()
    return args, net


def cpu_model(N, state_dict=None):
    """-> (kind, forward_fn(data) -> (cov, proba), plotwise_fn, raster_fn, args)."""
    from oracle import thirdparty_ops as tp  # noqa: F401  (the oracle is the CPU arm being timed)
    from oracle.pointnet2_port import PointNet2Port, project_to_2d_rasters_port, project_to_plotwise_coverages_port
    from oracle.ref_loader import load_reference, reference_root
    from sn2.config import default_args
    from sn2.synth import randomize_bn_

    args = default_args(subsample_size=N)
    torch.manual_seed(0)
    if reference_root() is not None:
        pn2, p2d = load_reference()
        net = pn2.PointNet2(args)
        kind = "reference"
        plotwise, raster = p2d.project_to_plotwise_coverages, p2d.project_to_2d_rasters
    else:
        net = PointNet2Port(args)
        kind = "port"
        plotwise, raster = project_to_plotwise_coverages_port, project_to_2d_rasters_port
    randomize_bn_(net)
    if state_dict is not None:
        net.load_state_dict(state_dict)
    net.eval()
    return kind, net, plotwise, raster, args


def cpu_step(net, plotwise, raster, args, data):
    with torch.no_grad():
        cov, proba = net(data)
        pw = plotwise(cov, data["cloud"], args)
        B, _, N = data["cloud"].shape
        cov_b = cov.view(B, N, 4).transpose(1, 2)
        rasters = [raster(data["cloud"][b], cov_b[b], args) for b in range(B)]
    return pw, rasters


def time_cpu(N, config_id, plots, repeats):
    from sn2.synth import synth_batch

    torch.set_num_threads(os.cpu_count())
    kind, net, plotwise, raster, args = cpu_model(N)
    data = synth_batch(config_id, plots, N)
    cpu_step(net, plotwise, raster, args, data)  # warm-up
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        cpu_step(net, plotwise, raster, args, data)
        best = min(best, time.perf_counter() - t0)
    return kind, plots / best, best


def static_config(cfg, world):
    """The part of `config` that names the WORKLOAD (identical for our arm and the reference arm)."""
    if cfg["mode"] != "infer":
        return {"workload": cfg["name"]}
    return {"workload": cfg["name"], "plots_per_gpu_per_step": cfg["B"], "points_per_plot": cfg["N"], "max_num_neighbors": 2000,
            "l2": "inputs larger than L2: steps rotate over 4 distinct input batches (218 MB)" if cfg["B"] > 1 else
                  "inputs rotate over 64 distinct single plots",
            "parallelism": f"plot-sharded x{world}"}


def run_reference(opts, cfg):
    """--impl reference: the reference's CPU implementation of the path on the host cores, on the SAME workload as our
    arm: every step is one full batch of the config (64 plots x 16 384 points at config 2), steps rotate over the same
    4 distinct batches.  The oracle evaluates a batch in chunks of 16 plots (eval mode: plots are independent) to bound
    its memory.  If the first step says K steps would not end within ~5 minutes, the batch is cut to a bounded sample
    and `config` says so (the rate per plot is what the ratio uses)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from sn2.synth import synth_batch

    N, B = cfg["N"], cfg["B"]
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm (rank 0 alone) is entitled to all host
    # cores, so under torchrun the value is put back before the oracle's OpenMP runtime is loaded (torch's own pool
    # is resized explicitly either way)
    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        os.environ["OMP_NUM_THREADS"] = str(ncores)
    torch.set_num_threads(ncores)
    kind, net, plotwise, raster, args = cpu_model(N)
    nrot = 4 if B > 1 else 8
    rot = [synth_batch(opts.config, B, N, first_plot=r * B) for r in range(nrot)]
    CH = 16

    def step(data, plots):
        for b0 in range(0, plots, CH):
            cpu_step(net, plotwise, raster, args, {k: v[b0:min(b0 + CH, plots)].contiguous() for k, v in data.items()})

    t0 = time.perf_counter()
    step(rot[0], min(B, CH))  # warm-up on one chunk (also the probe that sizes the run)
    per_plot = (time.perf_counter() - t0) / min(B, CH)
    plots = B
    budget_s = float(os.environ.get("SN2_REF_BUDGET_S", "300"))
    if per_plot * B * opts.steps > budget_s:
        plots = max(1, min(B, int(budget_s / (per_plot * opts.steps))))
    t0 = time.perf_counter()
    for i in range(opts.steps):
        step(rot[i % nrot], plots)
    dt = time.perf_counter() - t0
    value = plots * opts.steps / dt
    config = static_config(cfg, opts.gpus)
    if plots != B:
        config["plots_per_gpu_per_step"] = plots
        config["workload"] += f" [bounded sample: {plots} of {B} plots per step]"
    sample = (f"{plots} plots x {N} pts per step, {opts.steps} steps rotating over {nrot} distinct batches, evaluated in chunks of {CH} plots; "
              f"{'reference model files verbatim on restated third-party ops' if kind == 'reference' else 'oracle port'}")
    print(json.dumps({
        "impl": "reference", "metric": "plots/sec (PointNet2 eval forward + project_to_2d)", "value": value, "unit": "plots/s",
        "n_gpus": opts.gpus, "steps": opts.steps, "warmup": opts.warmup, "ms_per_step": 1e3 * dt / opts.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config,
        "points_per_s": value * N,
        "cpu_baseline": {"value": value, "unit": "plots/s", "cores": ncores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "plots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def init_dist(dev):
    """NCCL process group, with NCCL's start-up banner ("NCCL version ...", written to fd 1 when the communicator is
    created) sent to stderr: stdout carries exactly one JSON line."""
    import torch.distributed as dist

    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


def train_loss(proba, pw, gt, pdf):
    """Synthetic-pdf variant of the training loss used by round-1 tests / tools (kept for them); the bench uses the
    reference loss itself, sn2.losses.training_loss."""
    sel = lambda t: torch.cat([t[:, :1], t[:, 2:]], dim=1)  # noqa: E731
    mae = torch.sqrt((sel(pw) - sel(gt)) ** 2 + 1e-4).mean()
    nll = -torch.log((sel(proba).double() * pdf).sum(1) + 1e-6).mean().float()
    p = proba[:, 2:]
    ent = -(p * torch.log(p + 1e-6)).sum(1).mean()
    return mae + 0.10 * nll + 0.04 * ent


def synthetic_kde_lut(dev):
    """A smooth positive 3-column density on the reference's grid size (5000 knots, learning/kde_mixture.py:89-91) in
    place of the fitted KDE mixture (KDEpy is not installed)."""
    from sn2.losses import KdeLut

    X = np.linspace(-30.0, 30.0, 5000)
    a = np.abs(X)
    return KdeLut(X, np.stack([np.exp(-a), 0.5 * np.exp(-0.5 * (a - 1.0) ** 2), 0.1 + 0.05 * a]), dev)


_DIST = {}


def dist_ctx():
    """(world, rank, local, device); NCCL process group created once per process."""
    if not _DIST:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        rank = int(os.environ.get("RANK", "0"))
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dev = torch.device("cuda", local)
        if world > 1:
            init_dist(dev)
        _DIST.update(world=world, rank=rank, local=local, dev=dev)
    return _DIST["world"], _DIST["rank"], _DIST["local"], _DIST["dev"]


def run_train(opts, cfg, scaling="strong", quick=False):
    """Config 3: one training step = train-mode forward (BatchNorm batch statistics over the GLOBAL batch) + plot-wise
    projection + the reference loss (MAE + 0.1 NLL under a KDE look-up + 0.04 entropy) + backward + gradient
    all-reduce + Adam, data-parallel by plot.  scaling = "strong": the global batch of 32 plots is split over the
    ranks; "weak": 32 plots per rank.  Returns the result dict (rank 0 prints it)."""
    import torch.distributed as dist
    from model.project_to_2d import project_to_plotwise_coverages
    from sn2 import comm as sn2_comm
    from sn2 import losses, ops, parallel
    from sn2.optim import FusedAdam
    from sn2.pipeline import GraphedTrainStep, StageTimer, StructurePrefetcher, sa_recompute_allowed
    from sn2.synth import synth_batch

    world, rank, local, dev = dist_ctx()
    N = cfg["N"]
    Bg = cfg["B"] * (world if scaling == "weak" else 1)
    W = max(opts.warmup, 3)
    steps = opts.steps if not quick else min(opts.steps, 20)
    args, net = make_model(N, local)
    net.train()
    comm = None
    if world > 1:
        net = parallel.convert_sync_batchnorm(net)  # reference-exact BatchNorm over the GLOBAL batch
        comm = sn2_comm.get_comm()
    use_graph = not opts.no_graph
    opt = FusedAdam(net.parameters(), lr=args.lr, weight_decay=args.wd, comm=comm)  # learning/train.py:180-185
    lut = synthetic_kde_lut(dev)
    if scaling == "weak":
        mine = synth_batch(3, cfg["B"], N, first_plot=rank * cfg["B"])
        g = torch.Generator().manual_seed(9 + rank)
        mine["gt"] = torch.rand(cfg["B"], 4, generator=g)
    else:
        full = synth_batch(3, Bg, N)
        g = torch.Generator().manual_seed(9)
        full["gt"] = torch.rand(Bg, 4, generator=g)
        mine = parallel.shard_plots(full, rank, world)
    Bl = mine["cloud"].shape[0]
    host = {k: v.pin_memory() for k, v in mine.items()}
    dev_in = {k: v.to(dev) for k, v in mine.items()}
    loss_host = torch.empty((), dtype=torch.float64).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def step_fn(batch):
        """learning/train.py:52-66 (no host synchronisation inside: graph-safe); returns the loss tensor."""
        opt.zero_grad()
        cov, proba = net(batch)
        pw = project_to_plotwise_coverages(cov, net.last_cloud_device, args)
        pdf = lut.pdf(net.last_cloud_device, args.z_max)
        loss = losses.training_loss(pw, batch["gt"], proba, pdf, args.m, args.e)[0]
        loss.backward()
        opt.step(grad_scale=Bl / Bg)
        return loss.detach()

    def eager_step(inp, timer=None, read_loss=False):
        b = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in inp.items()}
        if timer is not None:  # per-stage events of the structural stage (fps1 ...) come from the model's own timer hooks
            opt.zero_grad()
            cov, proba = net({k: b[k] for k in ("xyz", "cloud")}, timer=timer)
            pw = project_to_plotwise_coverages(cov, net.last_cloud_device, args)
            loss = losses.training_loss(pw, b["gt"], proba, lut.pdf(net.last_cloud_device, args.z_max), args.m, args.e)[0]
            loss.backward()
            opt.step(grad_scale=Bl / Bg)
        else:
            loss = step_fn(b)
        if read_loss:
            loss_host.copy_(loss.detach(), non_blocking=True)

    gstep = GraphedTrainStep(net, step_fn, opt, device=dev) if use_graph else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        evs = []
        barrier()
        for _ in range(n):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        barrier()
        return sum(a.elapsed_time(b) for a, b in evs)

    def timed_prefetch(src, n, read_loss):
        """The training loop as a user writes it with StructurePrefetcher: the structural stage (FPS, ball query,
        kNN) of batch i+1 runs on a side stream under step i.  One event pair around the whole loop (L2 flush
        included); the first batch's structural stage is inside the region, none is computed beyond the last."""
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with no_gc():
            a.record()
            for batch in StructurePrefetcher(net, (src for _ in range(n)), dev):
                flush.fill_(1)
                if gstep is not None:  # whole step = input copies into the static buffers + one graph launch
                    loss = gstep(batch)
                    if read_loss:
                        loss_host.copy_(loss, non_blocking=True)
                else:
                    eager_step(batch, None, read_loss)
            b.record()
            barrier()
        return a.elapsed_time(b)

    # W untimed steps of EACH mode right before it is timed: graph capture empties the caching allocator
    # (torch.cuda.graph calls empty_cache), so the first eager steps after it pay their cudaMallocs again
    timed_prefetch(dev_in, W, False)   # captures the graph (every rank, together)
    sampler = ClockSampler(local)
    sampler.start()
    sampler.wait_ready()
    timer = StageTimer()
    ms_serial = ms_serial_e2e = None
    if not quick:
        for _ in range(W):
            eager_step(dev_in)
        ms_serial = timed(lambda: eager_step(dev_in, timer), steps)
        for _ in range(W):
            eager_step(host, read_loss=True)
        ms_serial_e2e = timed(lambda: eager_step(host, None, True), steps)
    timed_prefetch(dev_in, W, False)
    l0, r0 = ops.LAUNCHES, (gstep.replays if gstep is not None else 0)
    ms_res = timed_prefetch(dev_in, steps, False)
    launches = ops.LAUNCHES - l0  # structural stage (launched live) ...
    if gstep is not None:        # ... + our kernels inside each graph replay (counted at capture)
        launches += (gstep.replays - r0) * gstep.launches_per_replay
    timed_prefetch(host, W, True)
    ms_e2e = timed_prefetch(host, steps, True)
    clocks = sampler.stop()
    t = torch.tensor([ms_res, ms_e2e, ms_serial or 0.0, ms_serial_e2e or 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_res, ms_e2e = float(t[0]), float(t[1])
    comm_err = comm.status()[1] if comm is not None else 0
    value = Bg * steps / (ms_res / 1e3)
    mode = "peer" if comm is not None else ("nccl" if world > 1 else "single process")
    per_block = 2  # collectives per Linear-ReLU-BatchNorm block and step: statistics forward, sums backward
    out = {
        "metric": "plots/sec (PointNet2 training step: fwd + projection + loss + bwd + all-reduce + Adam)", "value": value,
        "unit": "plots/s", "n_gpus": world, "steps": steps, "warmup": W, "ms_per_step": ms_res / steps,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["name"], "global_batch": Bg, "plots_per_gpu": Bl, "points_per_plot": N,
                   "batchnorm": "SyncBatchNorm over the global batch (raw fp64 sums + counts)" if world > 1 else "single process",
                   "loss": "reference loss: MAE + 0.10 NLL (fp64, KDE look-up on the device, 5000 knots) + 0.04 entropy",
                   "optimizer": "FusedAdam (lr 1e-3, wd 1e-3): gradient all-reduce + Adam in one kernel over flat buckets",
                   "l2": "flushed between timed steps (256 MiB write, inside the timed region)",
                   "overlap": "StructurePrefetcher: FPS / ball query / kNN of batch i+1 on a side stream under step i",
                   "cuda_graph": ("GraphedTrainStep: fwd + loss + bwd + all-reduces + Adam replayed as ONE CUDA graph per rank (edge lists at "
                                  "fixed capacity, live counts on the device)") if use_graph else "off",
                   "sa_blocks": ("recompute sweeps (csrc/train_sa.cu: no per-edge array is written)" if sa_recompute_allowed(net.sa1_module.conv.local_nn)
                                 else "materialised messages (edge_msg + fused Linear-ReLU-BatchNorm + segment_max): data-parallel jobs keep "
                                      "them, the recompute sweeps are not yet verified on > 1 GPU (SN2_SA_RECOMPUTE_DP=1 opts in)"),
                   "parallelism": f"dp{world} by plot"},
        "points_per_s": value * N,
        "collectives": {"backend": mode,
                        "how": ("one-shot all-reduces over NVLink peer stores INSIDE the BatchNorm finalize / backward-coefficient / Adam "
                                "kernels (csrc/comm.cu): no NCCL call and no extra launch on the step's data path") if comm is not None else
                               ("torch.distributed (NCCL) all-reduce calls between the kernels" if world > 1 else "none"),
                        "allreduce_calls_per_step": 0 if comm is not None or world == 1 else 7 * per_block + 1,
                        "fused_exchanges_per_step": 7 * per_block + 1 if comm is not None else 0,
                        "bytes_per_step_to_each_peer": 8 * (2 * 260 + 7) + 8 * 2 * 260 + 4 * 14997,
                        "peer_error_word": comm_err},
        "e2e": {"value": Bg * steps / (ms_e2e / 1e3), "unit": "plots/s", "ms_per_step": ms_e2e / steps,
                "h2d_bytes_per_step": int(sum(v.numel() * v.element_size() for v in host.values())), "d2h_bytes_per_step": 8,
                "api": ("StructurePrefetcher + GraphedTrainStep(" if use_graph else "StructurePrefetcher + (") +
                       "PointNet2.forward (train) + project_to_plotwise_coverages + sn2.losses.training_loss + backward + FusedAdam.step)"},
        "gpu_launches": launches, "graph_eager_fallback_steps": gstep.eager_steps if gstep is not None else None, "clocks": clocks,
    }
    if not quick:
        ms_serial, ms_serial_e2e = float(t[2]), float(t[3])
        per_step = {k: v / steps for k, v in timer.totals_ms().items()}
        M1 = ops.m_of(N, args.ratio1)
        hbm_peak, peak_src = peaks()
        out["serial"] = {"value": Bg * steps / (ms_serial / 1e3), "ms_per_step": ms_serial / steps,
                         "e2e": Bg * steps / (ms_serial_e2e / 1e3), "unit": "plots/s",
                         "note": "plain eager loop, no prefetch, no graph; per-step event pairs, L2 flush outside them"}
        out["stage_ms_per_step"] = {k: round(v, 4) for k, v in sorted(per_step.items(), key=lambda kv: -kv[1])}
        if "fps1" in per_step:
            ach = (12 * N + 4 * M1) * Bl / (per_step["fps1"] / 1e3) / 1e9
            out["roofline"] = {"kernel": "fps1", "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                               "traffic": None, "peak_source": peak_src, "ms_per_launch": per_step["fps1"],
                               "note": "largest single custom kernel of the step (a serial chain on one SM per plot, latency bound; it "
                                       "runs on the side stream under the previous step); stage times are from the serial pass"}
    del gstep
    return out


def synthetic_parcel(dev, side=1040.0, density=32.0, seed=4):
    """1 km x 1 km parcel + 20 m buffer at 32 points / m^2 (~34.6 M points, ~10 k per 314 m^2 disk) in
    the layout of load_las_file (utils/load_data.py:149-184): float32 [10, P], cm-quantised coordinates.  Generated on the
    device with a fixed Philox seed: every rank builds the same cloud."""
    P = int(side * side * density)
    g = torch.Generator(device=dev).manual_seed(seed)
    r = lambda: torch.rand(P, generator=g, device=dev)  # noqa: E731
    cloud = torch.empty((10, P), dtype=torch.float32, device=dev)
    cloud[0] = torch.round(r() * side * 100.0) / 100.0
    cloud[1] = torch.round(r() * side * 100.0) / 100.0
    sel, h = r(), r()
    z = torch.where(sel < 0.55, 0.05 * h, torch.where(sel < 0.75, 0.5 * h, torch.where(sel < 0.85, 0.5 + h, 1.5 + 13.5 * h)))
    cloud[2] = torch.round((50.0 + 0.01 * cloud[0] + z) * 100.0) / 100.0   # gentle slope + vegetation heights
    for k in range(3, 7):
        cloud[k] = torch.floor(r() * 65536.0)
    cloud[7] = torch.floor(r() * 32768.0)
    cloud[8] = 1.0 + torch.floor(r() * 5.0)
    cloud[9] = 1.0 + torch.floor(r() * 5.0)
    return cloud


def run_parcel(opts, cfg):
    """Config 4: a 1 km^2 parcel (+ 20 m buffer), ~34.6 M points, tiled on the device into the reference's overlapping 10 m
    plots (sn2.parcel: ball query, <= 50-point filter, local-min-z, fake ground points, rescale, sub-sampling), PointNet2
    inference + rasters per batch of 64 plots, local-map fusion, ONE all-reduce of the accumulators, band finalisation.
    Strong scaling: the plot centres are split in contiguous blocks (x stripes) over the ranks; a rank only holds its
    stripe of the cloud."""
    import torch.distributed as dist
    from sn2 import ops
    from sn2.parallel import shard_bounds
    from sn2.parcel import LAS_PARCEL_BUFFER, ParcelCloud, keep_points_in_shape, plot_centers_reference, predict_parcel

    world, rank, local, dev = dist_ctx()
    B, N = cfg["B"], cfg["N"]
    args, net = make_model(N, local)
    full = synthetic_parcel(dev)
    Ptot = full.shape[1]
    x_min, x_max = float(full[0].min()), float(full[0].max())
    y_min, y_max = float(full[1].min()), float(full[1].max())
    centers = plot_centers_reference(x_min, x_max, y_min, y_max, args)
    shape = np.array([[20.0, 20.0], [1020.0, 20.0], [1020.0, 1020.0], [20.0, 1020.0]])  # the parcel inside its 20 m LAS buffer
    centers = centers[keep_points_in_shape(centers, shape, LAS_PARCEL_BUFFER + args.diam_meters // 2)]  # prepare_utils.py:146-151
    C = centers.shape[0]
    lo, hi = shard_bounds(C, rank, world)
    r = args.diam_meters // 2
    if world > 1:
        keep = (full[0] >= float(centers[lo:hi, 0].min()) - r - 0.5) & (full[0] <= float(centers[lo:hi, 0].max()) + r + 0.5)
        mine = full[:, keep].contiguous()
    else:
        mine = full
    del full
    Pm = mine.shape[1]
    mine_host = torch.empty(mine.shape, dtype=torch.float32).pin_memory()
    mine_host.copy_(mine)
    depth = max(opts.pipeline, 1)
    mosaic_host = None
    info = {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_pass(src, read_back):
        nonlocal mosaic_host
        parcel = ParcelCloud(src, dev)                       # H2D when src is the pinned host copy; xy grid over the stripe
        mosaic, inf = predict_parcel(net, args, parcel, centers, rank, world, batch=B, depth=depth)
        info.update(inf)
        if read_back and rank == 0:
            if mosaic_host is None:
                mosaic_host = torch.empty(mosaic.shape, dtype=torch.float64).pin_memory()
            mosaic_host.copy_(mosaic, non_blocking=True)
        return mosaic

    def timed(src, read_back, reps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with no_gc():
            a.record()
            for _ in range(reps):
                one_pass(src, read_back)
            b.record()
            barrier()
        return a.elapsed_time(b) / reps

    with torch.no_grad():
        one_pass(mine, False)
        one_pass(mine_host, True)
        sampler = ClockSampler(local)
        sampler.start()
        sampler.wait_ready()
        l0 = ops.LAUNCHES
        reps = max(2, opts.steps // 10)
        ms_res = timed(mine, False, reps)
        launches = (ops.LAUNCHES - l0)
        ms_e2e = timed(mine_host, True, reps)
        clocks = sampler.stop()
        last = one_pass(mine, False)
    # the one collective of the pass, timed on its own: NCCL all-reduce of the [7,H,W] float64 fusion accumulators
    ar_ms = None
    if world > 1:
        acc = torch.zeros((7, info["H"], info["W"]), dtype=torch.float64, device=dev)
        for _ in range(3):
            dist.all_reduce(acc)
        barrier()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        for _ in range(10):
            dist.all_reduce(acc)
        eb.record()
        torch.cuda.synchronize()
        ar_ms = ea.elapsed_time(eb) / 10
    t = torch.tensor([ms_res, ms_e2e, ar_ms or 0.0], dtype=torch.float64, device=dev)
    nvalid = torch.tensor([info["plots_valid"]], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(nvalid)
        ar_ms = float(t[2])
    ms_res, ms_e2e = float(t[0]), float(t[1])
    covered = float((~torch.isnan(last[4])).float().mean())
    out = {
        "metric": "plots/sec (parcel inference: tiling + plot preparation + PointNet2 eval forward + rasters + local-map fusion + band finalisation)",
        "value": C / (ms_res / 1e3), "unit": "plots/s", "n_gpus": world, "steps": reps, "warmup": 2, "ms_per_step": ms_res,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["name"], "parcel_points": int(Ptot), "points_on_this_rank": int(Pm), "plots": int(C),
                   "plots_valid": int(nvalid), "plots_per_gpu": int(hi - lo), "points_per_plot": N, "batch": B, "batches_in_flight": depth,
                   "mosaic": [5, info["H"], info["W"]], "mosaic_covered_fraction": covered,
                   "hard_medium_vegetation_threshold": float(info["threshold"]),
                   "l2": "inputs larger than L2 (a 1.4 GB cloud; ~100 distinct 33 MB plot batches per pass)",
                   "parallelism": f"plot centres in contiguous blocks (x stripes) x{world}, one all-reduce of the [7,{info['H']},{info['W']}] f64 accumulators",
                   "allreduce_ms": ar_ms, "allreduce_bytes": int(7 * info["H"] * info["W"] * 8)},
        "points_per_s": Ptot / (ms_res / 1e3),
        "e2e": {"value": C / (ms_e2e / 1e3), "unit": "plots/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(mine_host.numel() * 4), "d2h_bytes_per_step": int(5 * info["H"] * info["W"] * 8),
                "api": "sn2.parcel.ParcelCloud(pinned host LAS array) + predict_parcel (mosaic read back to pinned host memory)"},
        "gpu_launches": launches, "clocks": clocks, "roofline": None,
    }
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="config 3: eager prefetch loop instead of the CUDA-graph step")
    ap.add_argument("--train-scaling", default="strong", choices=["strong", "weak"], help="--config 3: split 32 plots, or 32 per GPU")
    ap.add_argument("--extras", default="auto", choices=["auto", "all", "none"],
                    help="also measure configs 1, 3 (weak + strong), 4 in the same run and attach them to the JSON line "
                         "(auto: when --config is the default 2)")
    ap.add_argument("--pipeline", type=int, default=3, help="batches in flight for value / e2e (0 = serial steps only)")
    ap.add_argument("--prewarm-s", type=float, default=1.5, help="seconds of untimed steps before anything is timed (clock / cache ramp)")
    opts = ap.parse_args()
    cfg = CONFIGS[opts.config]
    # stdout carries exactly one JSON line: NCCL's version banner (NCCL_DEBUG=VERSION in some images) goes away,
    # an explicit NCCL_DEBUG=INFO / TRACE from the caller is respected
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    if opts.impl == "reference":
        return run_reference(opts, cfg)
    world, rank, local, dev = dist_ctx()
    if cfg["mode"] == "train":
        out = run_train(opts, cfg, opts.train_scaling)
    elif cfg["mode"] == "parcel":
        out = run_parcel(opts, cfg)
    else:
        out = run_infer(opts, cfg)
    extras = opts.extras == "all" or (opts.extras == "auto" and opts.config == 2)
    if extras:
        # The other BASELINE configs under the same (driver-run) command, each a bounded measurement of its own:
        # config 3 (the one path with collectives) weak AND strong, config 4 (plot-sharded parcel + one all-reduce),
        # config 1 (single plots).  Failures are reported in place, they never take the headline down.
        def guarded(fn):
            try:
                r = fn()
                torch.cuda.synchronize()
                return r
            except Exception as e:  # noqa: BLE001
                return {"error": f"{type(e).__name__}: {e}"[:300]}
        keep = ("metric", "value", "unit", "ms_per_step", "scaling", "config", "collectives", "e2e", "gpu_launches", "serial",
                "graph_eager_fallback_steps", "n_gpus", "steps")
        slim = lambda d: {k: d[k] for k in keep if k in d} if "error" not in d else d  # noqa: E731
        if opts.config != 3:
            out["train"] = {sc: slim(guarded(lambda sc=sc: run_train(opts, CONFIGS[3], sc, quick=True))) for sc in ("weak", "strong")}
        if opts.config != 4:
            out["parcel"] = slim(guarded(lambda: run_parcel(opts, CONFIGS[4])))
        if opts.config != 1 and world == 1:
            o1 = argparse.Namespace(**vars(opts))
            o1.config = 1
            out["single_plot"] = slim(guarded(lambda: run_infer(o1, CONFIGS[1], cpu_baseline=False)))
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        import torch.distributed as dist

        from sn2 import comm as sn2_comm

        sn2_comm.close_all()
        dist.destroy_process_group()


def run_infer(opts, cfg, cpu_baseline=True):
    """Configs 1 / 2: eval forward + both projections on B plots per step and rank (weak scaling: plots are independent)."""
    import torch.distributed as dist
    from model.project_to_2d import project_to_2d_rasters_batched, project_to_plotwise_coverages
    from sn2 import ops
    from sn2.pipeline import StageTimer, forward_eval
    from sn2.synth import synth_batch

    world, rank, local, dev = dist_ctx()
    B, N = cfg["B"], cfg["N"]
    W = max(opts.warmup, 3)
    args, net = make_model(N, local)
    data = synth_batch(opts.config, B, N, first_plot=rank * B)  # every rank its own plots
    host = {k: v.pin_memory() for k, v in data.items()}
    dev_in = {k: v.to(dev) for k, v in data.items()}
    D = args.diam_pix
    pw_host = torch.empty((B, 4), dtype=torch.float32).pin_memory()
    rs_host = torch.empty((B, 3, D, D), dtype=torch.float64).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step_resident(timer=None):
        cov, proba, g, cloud_d = forward_eval(net, dev_in["xyz"], dev_in["cloud"], dev, 2000, None, timer)
        T = timer.stage if timer is not None else None
        if T:
            with T("project_plotwise"):
                pw = ops.project_plotwise(cloud_d, cov, D)
            with T("project_rasters"):
                rs = ops.project_rasters(cloud_d, cov, "point_major", D, args.diam_meters)
        else:
            pw = ops.project_plotwise(cloud_d, cov, D)
            rs = ops.project_rasters(cloud_d, cov, "point_major", D, args.diam_meters)
        return pw, rs

    def step_e2e():
        with torch.no_grad():
            cov, proba = net(host)                                             # H2D of xyz + cloud inside
            pw = project_to_plotwise_coverages(cov, net.last_cloud_device, args)
            rs = project_to_2d_rasters_batched(net.last_cloud_device, cov, args)
        pw_host.copy_(pw, non_blocking=True)                                   # D2H of the step's results
        rs_host.copy_(rs, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, with_timer):
        timer = StageTimer() if with_timer else None
        evs = []
        barrier()
        with no_gc():
            for _ in range(steps):
                flush.fill_(1)  # L2 flush between timed iterations (outside the event pair)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                if with_timer:
                    fn(timer)
                else:
                    fn()
                b.record()
                evs.append((a, b))
            barrier()
        per = [a.elapsed_time(b) for a, b in evs]
        total_ms = sum(per)
        if os.environ.get("SN2_BENCH_TRACE") == "1":
            print(f"[trace] per-step ms: median {float(np.median(per)):.3f}, max {max(per):.3f}, "
                  f"steps > 1.5x median: {[round(v, 2) for v in per if v > 1.5 * float(np.median(per))]}", file=sys.stderr)
        return total_ms, timer

    with torch.no_grad():
        t_pre = time.perf_counter()
        while time.perf_counter() - t_pre < opts.prewarm_s:  # fresh boxes ramp clocks / page in lazily: not timed
            step_resident()
            torch.cuda.synchronize()
        for _ in range(W):
            step_resident()
            step_e2e()
        sampler = ClockSampler(local)
        sampler.start()
        sampler.wait_ready()
        t_pre = time.perf_counter()
        while time.perf_counter() - t_pre < float(os.environ.get("SN2_BENCH_SETTLE_S", "1.0")):  # untimed, sampler running
            step_resident()
            torch.cuda.synchronize()
        # untimed ASYNCHRONOUS passes: the loops above synchronise after every step, so the host never runs ahead and
        # the allocator pools stay shallow; the first free-running pass then grows them (cudaMalloc calls of
        # 50-200 ms in the middle of a step, seen as 1-3 outlier steps in the first timed pass only)
        timed(step_resident, max(25, opts.steps // 2), False)
        timed(step_e2e, 10, False)
        l0 = ops.LAUNCHES
        nalloc = lambda: torch.cuda.memory_stats(dev).get("num_device_alloc", 0)  # noqa: E731  cudaMalloc calls so far
        n0 = nalloc()
        ms_res, _ = timed(step_resident, opts.steps, False)
        launches = ops.LAUNCHES - l0
        if os.environ.get("SN2_BENCH_TRACE") == "1":
            ms_ = torch.cuda.memory_stats(dev)
            print(f"[trace] serial resident pass: {ms_res / opts.steps:.3f} ms/step, cudaMalloc calls during the pass: {nalloc() - n0}, "
                  f"reserved {torch.cuda.memory_reserved(dev) / 2**30:.1f} GiB, allocated {torch.cuda.memory_allocated(dev) / 2**30:.1f} GiB, "
                  f"retries {ms_.get('num_alloc_retries')}, cudaFree calls {ms_.get('num_device_free')}", file=sys.stderr)
        ms_e2e, _ = timed(step_e2e, opts.steps, False)
        # same steps once more with per-stage events (instrumented pass: feeds stage_ms_per_step / roofline only)
        _, timer = timed(step_resident, opts.steps, True)
        clocks = sampler.stop()

    # ---- throughput mode: two batches in flight (sn2.pipeline.InferencePipeline), K steps timed as one region.
    # Inputs rotate over 4 distinct batches (4 x 54.5 MB > the 126 MB L2), so no step finds its input in L2.
    from sn2.pipeline import InferencePipeline
    nrot = 4 if cfg["B"] > 1 else 64
    rot = [synth_batch(opts.config, B, N, first_plot=(rank * nrot + r) * B) for r in range(nrot)] if opts.pipeline else []
    rot_host = [{k: v.pin_memory() for k, v in d.items()} for d in rot]
    rot_dev = [{k: v.to(dev) for k, v in d.items()} for d in rot]

    # small batches (config 1: single plots) are launch bound on the host: more slots, one CUDA graph per slot
    small = B * N <= (1 << 18)
    depth = opts.pipeline if opts.pipeline else 0
    if small and opts.pipeline == 3 and not opts.no_graph:
        depth = 12
    use_graph = small and not opts.no_graph

    PIPE_REPS = 3

    def timed_pipe(batches, keep):
        """K submits timed as one region, repeated PIPE_REPS times on the same pipeline; returns every pass (ms).
        The reported value is the MEDIAN pass: a one-off allocator / driver stall (seen once in ~10 runs, a whole
        pass 2-3x slower with clocks unchanged) then shows in `passes_ms_per_step` instead of in the headline."""
        pipe = InferencePipeline(net, args, depth=depth, graph=use_graph)
        for i in range(max(W, 20, depth + 1)):
            pipe.submit(batches[i % nrot], keep_on_device=keep)
        pipe.drain()
        out = []
        for _ in range(PIPE_REPS):
            with no_gc():  # entered (and collected once) BEFORE the region starts
                barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for i in range(opts.steps):
                    pipe.submit(batches[i % nrot], keep_on_device=keep)
                for s_ in pipe.sets:
                    torch.cuda.current_stream().wait_stream(s_[0])
                b.record()
                barrier()
            out.append(a.elapsed_time(b))
        return out

    ms_pipe_res = ms_pipe_e2e = None
    pipe_passes = None
    if opts.pipeline:
        with torch.no_grad():
            res_passes = timed_pipe(rot_dev, True)
            e2e_passes = timed_pipe(rot_host, False)
        ms_pipe_res, ms_pipe_e2e = float(np.median(res_passes)), float(np.median(e2e_passes))
        pipe_passes = {"resident": [round(v / opts.steps, 4) for v in res_passes], "e2e": [round(v / opts.steps, 4) for v in e2e_passes],
                       "reported": "median pass"}

    t = torch.tensor([ms_res, ms_e2e, ms_pipe_res or 0.0, ms_pipe_e2e or 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_res, ms_e2e = float(t[0]), float(t[1])
    plots_total = B * world * opts.steps
    serial_value = plots_total / (ms_res / 1e3)
    serial_e2e = plots_total / (ms_e2e / 1e3)
    if opts.pipeline:
        ms_res_used, ms_e2e_used = float(t[2]), float(t[3])
    else:
        ms_res_used, ms_e2e_used = ms_res, ms_e2e
    value = plots_total / (ms_res_used / 1e3)
    e2e_value = plots_total / (ms_e2e_used / 1e3)

    # roofline of the dominant kernel (largest share of the resident step), live CUDA-event time
    stages = timer.totals_ms()
    per_step = {k: v / opts.steps for k, v in stages.items()}
    dom = max(per_step, key=per_step.get)
    M1, M2 = ops.m_of(N, args.ratio1), ops.m_of(ops.m_of(N, args.ratio1), args.ratio2)
    alg_bytes = {  # SURVEY.md §8d algorithmic bytes per plot, x B plots per launch
        "fps1": 12 * N + 4 * M1, "fps2": 12 * M1 + 4 * M2,
        "knn1": 12 * (N + M1) + 24 * N, "knn2": 12 * (M1 + M2) + 24 * M1,
        "fp1_head": 4 * 34 * M1 + 4 * 8 * N + 24 * N + 2 * 16 * N,
        "ingest": 4 * 13 * N + 4 * 12 * N,
    }
    hbm_peak, peak_src = peaks()
    traffic = None  # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
    tpath = os.path.join(ROOT, "profiles", "r2_ncu_traffic_cfg2.json")
    if opts.config == 2 and os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic = {"fps1": tj.get("fps_bucket_kernel<8, 32, 0, 1, 1, 2>"), "fps2": tj.get("fps_kernel<512, 8, 1>"),
                   "sa1_fused": tj.get("sa_fused_kernel<1, 0, 0>"), "fp1_head": tj.get("fp1_head_kernel")}.get(dom)
    roof = None
    if dom in alg_bytes:
        ach = alg_bytes[dom] * B / (per_step[dom] / 1e3) / 1e9
        roof = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                "traffic": traffic, "peak_source": peak_src, "ms_per_launch": per_step[dom],
                "algorithmic_bytes_per_launch": alg_bytes[dom] * B}
        if dom.startswith("fps"):
            n, m = (N, M1) if dom == "fps1" else (M1, M2)
            evals = float(n) * m * B
            roof["note"] = ("FPS is a serial chain of M dependent arg-max iterations per plot: latency/SIMT bound, not HBM bound; "
                            "on-chip model below")
            roof["fps_distance_evals_per_s"] = evals / (per_step[dom] / 1e3)
            roof["fps_ns_per_iteration"] = per_step[dom] * 1e6 / m

    # ---- every stage against both rooflines (SURVEY.md §8d formulas; E1 / E2 = edge counts of THIS batch) ----------
    from sn2.pipeline import ForwardTrace
    with torch.no_grad():
        tr = ForwardTrace()
        net(dev_in, trace=tr)
        E1, E2 = int(tr.tensors["rowptr1"][-1]) / B, int(tr.tensors["rowptr2"][-1]) / B
        del tr
    D2 = D * D
    per_plot = {  # stage: (algorithmic bytes, algorithmic flops [reference formulation], what bounds it)
        "ingest": (4 * 13 * N + 4 * 12 * N, 0, "hbm"),
        "fps1": (12 * N + 4 * M1, 8.0 * N * M1, "latency chain (M dependent arg-max iterations on one SM per plot)"),
        "fps2": (12 * M1 + 4 * M2, 8.0 * M1 * M2, "latency chain; runs on a side stream under sa1_fused"),
        "sa1_fused": (4 * 11 * N + 4 * M1 + 4 * 16 * M1, 864.0 * E1, "fp32 issue + gather latency"),
        "sa2_fused": (4 * 19 * M1 + 4 * M2 + 4 * 32 * M2, 1216.0 * E2, "fp32 issue + gather latency"),
        "global_sa": (4 * 35 * M2 + 256, 4480.0 * M2, "fp32 issue (tiny)"),
        "fp3": (4 * 32 * M2 + 256 + 4 * 64 * M2, 12288.0 * M2, "fp32 issue (tiny)"),
        "knn2": (12 * (M1 + M2) + 24 * M1, 0, "L1/L2 gather; side stream (time includes waiting for SMs)"),
        "fp2": (4 * 64 * M2 + 4 * 16 * M1 + 24 * M1 + 4 * 34 * M1, 5440.0 * M1, "fp32 issue"),
        "knn1": (12 * (N + M1) + 24 * N, 0, "L1/L2 gather; side stream (time includes waiting for SMs)"),
        "fp1_head": (4 * 34 * M1 + 32 * N + 24 * N + 32 * N, (2856.0 + 1088 + 160) * N, "fp32 issue"),
        "project_plotwise": (8 * N + 16 * N + 16, 0, "hbm (tiny)"),
        "project_rasters": (8 * N + 16 * N + 8 * 3 * D2, 0, "hbm (tiny)"),
    }
    fp32_peak = 148 * 128 * 2 * (clocks.get("sm_max_mhz") or 1965.0) * 1e6 / 1e12  # TFLOP/s: 128 FMA / clk / SM
    roofline_all = {}
    for k, (by, fl, bound) in per_plot.items():
        if k not in per_step or per_step[k] <= 0:
            continue
        sec = per_step[k] / 1e3
        e = {"ms": round(per_step[k], 4), "algorithmic_bytes": int(by * B), "hbm_gbs": by * B / sec / 1e9,
             "hbm_frac": by * B / sec / 1e9 / hbm_peak, "bound": bound}
        if fl:
            e["algorithmic_flops"] = fl * B
            e["fp32_tflops"] = fl * B / sec / 1e12
            e["fp32_issue_frac"] = fl * B / sec / 1e12 / fp32_peak
        roofline_all[k] = e

    out = {
        "metric": "plots/sec (PointNet2 eval forward + project_to_2d)", "value": value, "unit": "plots/s",
        "n_gpus": world, "steps": opts.steps, "warmup": W, "ms_per_step": ms_res_used / opts.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": static_config(cfg, world),
        "measurement": {"batches_in_flight": depth or 1, "cuda_graph_per_batch": bool(use_graph and opts.pipeline),
                        "passes_ms_per_step": pipe_passes,
                        "timed_region": ("K submits timed as one region (median of 3 passes)" if opts.pipeline else
                                         "per-step event pairs, L2 flushed between steps (256 MiB write outside the pairs)")},
        "points_per_s": value * N,
        "serial": {"value": serial_value, "ms_per_step": ms_res / opts.steps, "e2e_value": serial_e2e,
                   "e2e_ms_per_step": ms_e2e / opts.steps,
                   "note": "one batch at a time, L2 flushed between steps; stage_ms_per_step and roofline are measured in this mode"},
        "e2e": {"value": e2e_value, "unit": "plots/s", "ms_per_step": ms_e2e_used / opts.steps,
                "h2d_bytes_per_step": int(sum(v.numel() * 4 for v in host.values())),
                "d2h_bytes_per_step": int(pw_host.numel() * 4 + rs_host.numel() * 8),
                "api": (f"sn2.pipeline.InferencePipeline.submit / result ({depth} batches in flight): H2D from pinned host memory -> PointNet2 eval "
                        "forward -> project_to_plotwise_coverages + project_to_2d_rasters -> D2H into pinned host memory, per batch"
                        if opts.pipeline else
                        "model.point_net2.PointNet2.forward + model.project_to_2d (pinned host in, host out), one batch at a time")},
        "e2e_dropin": {"value": serial_e2e, "unit": "plots/s", "ms_per_step": ms_e2e / opts.steps,
                       "api": "model.point_net2.PointNet2.forward + model.project_to_2d.project_to_plotwise_coverages / "
                              "project_to_2d_rasters_batched, one call per batch, pinned host in, host out (the reference's own call sequence, "
                              "predict.py:103-126), L2 flushed between steps"},
        "gpu_launches": launches,
        "clocks": clocks,
        "stage_ms_per_step": {k: round(v, 4) for k, v in sorted(per_step.items(), key=lambda kv: -kv[1])},
        "stage_note": "stage times come from the serial pass; fps2 / knn2 / knn1 run on side streams concurrently with sa1_fused, so their "
                      "event pairs include waiting for SMs (knn2 alone: 0.07 ms under ncu)",
        "roofline": roof,
        "roofline_all": roofline_all,
    }
    if rank == 0 and world == 1 and not opts.no_cpu_baseline and cpu_baseline:
        sample_plots = 16 if cfg["B"] >= 16 else cfg["B"]
        kind, v, best = time_cpu(N, opts.config, sample_plots, repeats=2)
        out["cpu_baseline"] = {"value": v, "unit": "plots/s", "cores": os.cpu_count(), "kind": kind,
                               "sample": f"{sample_plots} plots x {N} pts, best of 2 after 1 warm-up ({best:.2f} s); "
                                         "reference model files verbatim on restated third-party ops" if kind == "reference"
                               else f"{sample_plots} plots x {N} pts, best of 2 after 1 warm-up ({best:.2f} s); oracle port"}
    return out


if __name__ == "__main__":
    main()
