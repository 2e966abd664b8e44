"""Host-side data-parallel logic on CPU: world_size-2 gloo runs of the plot sharding, the flat gradient
bucket (weighted all-reduce == full-batch gradient) and the result gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sn2 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
        B, N = 5, 11  # 5 plots of 11 points: uneven split 3 + 2
        g = torch.Generator().manual_seed(1)
        batch = {"cloud": torch.rand(B, N, 5, generator=g), "gt": torch.rand(B, 3, generator=g), "tag": "x"}

        def loss_of(d):  # mean over plots of a mean over points, like the reference losses
            return ((net(d["cloud"]).mean(dim=1) - d["gt"]) ** 2).mean()

        # single-process reference gradient on the full batch
        net.zero_grad()
        loss_of(batch).backward()
        want = torch.cat([p.grad.reshape(-1) for p in net.parameters()]).clone()

        bucket = parallel.GradBucket(net)
        bucket.zero()
        mine = parallel.shard_plots(batch, rank, world)
        assert mine["tag"] == "x"
        lo, hi = parallel.shard_bounds(B, rank, world)
        assert mine["cloud"].shape[0] == hi - lo
        loss_of(mine).backward()
        bucket.allreduce(hi - lo, B)
        got = bucket.flat.clone()
        ok_grad = torch.allclose(got, want, rtol=1e-5, atol=1e-7)
        ok_view = all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in net.parameters())

        res = parallel.gather_plot_results(mine["gt"] * 1.0)
        ok_gather = True if rank != 0 else bool(torch.equal(res, batch["gt"]))
        ret[rank] = (ok_grad, ok_view, ok_gather)
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_all_plots():
    for B in (1, 5, 32, 33):
        for world in (1, 2, 4, 8):
            spans = [parallel.shard_bounds(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_gloo_world2_gradient_bucket_and_gather():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    for rank in range(world):
        assert ret[rank] == (True, True, True), (rank, ret[rank])


def test_sync_batchnorm_conversion_keeps_state_dict_keys():
    from model.point_net2 import PointNet2
    from sn2.config import default_args

    net = PointNet2(default_args())
    keys = list(net.state_dict().keys())
    net = parallel.convert_sync_batchnorm(net)
    assert list(net.state_dict().keys()) == keys
    assert isinstance(net.sa1_module.conv.local_nn[0][2], torch.nn.SyncBatchNorm)


def _route_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from model.point_net2 import MLP
        from sn2.pipeline import sa_recompute_allowed

        plain = MLP([11, 16, 16]).train()
        sync = torch.nn.SyncBatchNorm.convert_sync_batchnorm(MLP([11, 16, 16])).train()
        os.environ.pop("SN2_SA_RECOMPUTE", None)
        os.environ.pop("SN2_SA_RECOMPUTE_DP", None)
        a, b = sa_recompute_allowed(plain), sa_recompute_allowed(sync)
        os.environ["SN2_SA_RECOMPUTE_DP"] = "1"
        c = sa_recompute_allowed(sync)
        os.environ.pop("SN2_SA_RECOMPUTE_DP")
        os.environ["SN2_SA_RECOMPUTE"] = "0"
        d = sa_recompute_allowed(plain)
        os.environ.pop("SN2_SA_RECOMPUTE")
        ret[rank] = (a, b, c, d)
    finally:
        dist.destroy_process_group()


def test_gloo_world2_sa_recompute_routing():
    """sn2.pipeline.sa_recompute_allowed: the recompute SA blocks (csrc/train_sa.cu) serve single-process training; a
    SyncBatchNorm with a live process group of two ranks keeps the materialising blocks unless SN2_SA_RECOMPUTE_DP=1
    (that combination was not verified on > 1 GPU in round 2); SN2_SA_RECOMPUTE=0 switches the blocks off."""
    from model.point_net2 import MLP
    from sn2.pipeline import sa_recompute_allowed

    # no process group: a SyncBatchNorm module behaves like BatchNorm1d -> allowed
    assert sa_recompute_allowed(torch.nn.SyncBatchNorm.convert_sync_batchnorm(MLP([19, 32])).train())
    world, port = 2, _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_route_worker, args=(world, port, ret), nprocs=world, join=True)
    for rank in range(world):
        assert ret[rank] == (True, False, True, False), (rank, ret[rank])
