"""CPU tests of the host side: drop-in module layout, checkpoint round trip, weight packing, C-ABI
exports.  No kernel is launched here (no GPU in this container)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from sn2 import _lib, ops, weights
from sn2.config import default_args

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_state_dict_is_reference_layout(tmp_path):
    from model.point_net2 import PointNet2
    from oracle.pointnet2_port import PointNet2Port

    args = default_args(stats_path=str(tmp_path))
    net = PointNet2(args)
    sd = net.state_dict()
    ref_keys = list(PointNet2Port(args).state_dict().keys())
    assert list(sd.keys()) == ref_keys and len(sd) == 53
    assert sum(p.numel() for p in net.parameters()) == 14997
    # checkpoint round trip through the reference's file naming / dict layout (a14)
    net.best_metric_epoch, net.best_metric_value = 12, 0.25
    net.save_state(args)
    path = os.path.join(str(tmp_path), "PCC_model_full.pt")
    ck = torch.load(path)
    assert set(ck) == {"best_metric_epoch", "state_dict", "best_metric_value"}
    other = PointNet2(args).load_state(path)
    assert other.best_metric_epoch == 12 and other.best_metric_value == 0.25
    for k in sd:
        assert torch.equal(sd[k], other.state_dict()[k])
    args.current_fold_id = 3
    net.save_state(args)
    assert os.path.exists(os.path.join(str(tmp_path), "PCC_model_fold_n=3.pt"))
    assert torch.equal(net.load_best_state(args).state_dict()["lin2.bias"], sd["lin2.bias"])


def test_reference_checkpoint_loads_when_reference_reachable(tmp_path):
    from oracle.ref_loader import load_reference, reference_root

    if reference_root() is None:
        pytest.skip("reference files not reachable")
    from model.point_net2 import PointNet2

    pn2, _ = load_reference()
    args = default_args(stats_path=str(tmp_path))
    ref = pn2.PointNet2(args)
    ref.save_state(args)
    ours = PointNet2(args).load_best_state(args)
    for (k1, v1), (k2, v2) in zip(ref.state_dict().items(), ours.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
    ours.save_state(args)
    ref.load_best_state(args)  # and back


def test_early_stopping_bookkeeping(tmp_path):
    from model.point_net2 import PointNet2

    args = default_args(stats_path=str(tmp_path), patience_in_epochs=2, epoch_to_start_early_stop=3)
    net = PointNet2(args)
    assert net.stop_early(1.0, 1, args) is False and net.best_metric_epoch == 1
    assert net.stop_early(2.0, 2, args) is False          # before epoch_to_start_early_stop
    assert net.stop_early(2.0, 3, args) is True and net.stopped_early
    assert net.stop_early(0.5, 4, args) is False and net.best_metric_value == 0.5


def test_layout_helpers_roundtrip():
    from model.point_net2 import PointNet2

    net = PointNet2(default_args(subsample_size=7))
    data = torch.arange(2 * 3 * 7, dtype=torch.float32).view(2, 3, 7)
    long = PointNet2.get_long_form(data)
    assert long.shape == (14, 3) and torch.equal(long[8], data[1, :, 1])
    assert torch.equal(net.get_batch_format(long), data)


def test_forward_fails_loudly_without_cuda():
    from model.point_net2 import PointNet2
    from model.project_to_2d import project_to_plotwise_coverages

    net = PointNet2(default_args(subsample_size=16)).eval()
    data = {"xyz": torch.zeros(1, 3, 16), "cloud": torch.zeros(1, 10, 16)}
    with pytest.raises(RuntimeError):
        net(data)
    with pytest.raises(RuntimeError):
        project_to_plotwise_coverages(torch.zeros(16, 4), data["cloud"], default_args())
    with pytest.raises(RuntimeError):
        ops.fps(torch.zeros(8, 3), None, 0.5)


def test_weight_packing_sizes_and_bn_fold():
    from model.point_net2 import PointNet2
    from sn2.synth import randomize_bn_

    net = PointNet2(default_args())
    randomize_bn_(net)
    packed = weights.pack_eval(net)
    assert {k: v.numel() for k, v in packed.items()} == weights.EXPECTED_SIZES
    blk = net.sa2_module.conv.local_nn[0]
    w = packed["sa2"]
    assert torch.equal(w[:19 * 32].view(19, 32), blk[0].weight.detach().t())
    s = w[19 * 32 + 32:19 * 32 + 64]
    t = w[19 * 32 + 64:]
    x = torch.rand(5, 32)
    want = torch.nn.functional.batch_norm(x, blk[2].running_mean, blk[2].running_var, blk[2].weight, blk[2].bias, False)
    torch.testing.assert_close(x * s + t, want, rtol=1e-5, atol=1e-6)
    v0 = weights.params_version(net)
    with torch.no_grad():
        net.lin1.bias.add_(1.0)
    assert weights.params_version(net) != v0


def test_dense_batch_validation():
    b = torch.tensor([0, 0, 1, 1])
    assert ops._dense_shape(b, 4) == (2, 2)
    with pytest.raises(RuntimeError):
        ops._dense_shape(torch.tensor([0, 0, 0, 1]), 4)
    with pytest.raises(RuntimeError):
        ops._dense_shape(torch.tensor([0, 1, 0, 1]), 4)
    assert ops.m_of(10000, 0.25) == 2500 and ops.r2_of(np.sqrt(8.0)) == 8.0


def test_library_exports_every_declared_symbol():
    """The .so loads and exports exactly what include/sn2.h declares (no compute call)."""
    hdr = open(os.path.join(ROOT, "include", "sn2.h")).read()
    declared = set(re.findall(r"\b(sn2_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    if not os.path.exists(_lib.LIB_PATH):
        subprocess.check_call(["python", os.path.join(ROOT, "stratanet2-vegetation-coverage-maps_b200", "build.py")])
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    loaded = _lib.load(require_cuda=False)
    assert loaded.sn2_abi_version() == 1
    assert b"invalid" in loaded.sn2_error_string(-1)
    assert loaded.sn2_fps_max_points() >= 16384


def test_run_mlp_dispatch_rules(monkeypatch):
    """Host logic of sn2.autograd_ops.run_mlp (no GPU: the library and the fused Function are stubbed): which blocks
    run fused, where the BatchNorm transform is deferred to the consumer, and the fixed-capacity guard."""
    import types

    from model.point_net2 import MLP
    from sn2 import autograd_ops as ao

    supported = {(11, 16), (16, 16), (19, 32), (80, 34), (42, 34)}
    fake = types.SimpleNamespace(sn2_lrb_supported=lambda co, ci: int((ci, co) in supported),
                                 sn2_linear_wgrad_supported=lambda co, ci: 0)
    monkeypatch.setattr(ao._lib, "load", lambda *a, **k: fake)
    calls = []

    def fake_apply(x, w, b, gamma, beta, bn, rows=None, in_ss=None, apply=True):
        calls.append(dict(ci=w.shape[1], co=w.shape[0], rows=rows, in_ss=in_ss, apply=apply))
        y = torch.zeros(x.shape[0], w.shape[0])
        return y if apply else (y, torch.full((4 * w.shape[0],), float(len(calls))))

    monkeypatch.setattr(ao.LinReluBN, "apply", staticmethod(fake_apply))
    mlp = MLP([11, 16, 16]).train()
    x = torch.zeros(65536, 11)

    y, ss = ao.run_mlp(mlp, x, defer_last=True)         # both fused; BN 1 applied by block 2, BN 2 by the caller
    assert [c["apply"] for c in calls] == [False, False] and calls[0]["in_ss"] is None
    assert calls[1]["in_ss"] is not None and float(calls[1]["in_ss"][0]) == 1.0 and float(ss[0]) == 2.0
    calls.clear()
    z = ao.run_mlp(mlp, x)                               # the last transform is materialised for a torch consumer
    assert [c["apply"] for c in calls] == [False, True] and torch.is_tensor(z)
    calls.clear()
    monkeypatch.setenv("SN2_DEFER_BN", "0")
    y, ss = ao.run_mlp(mlp, x, defer_last=True)
    assert [c["apply"] for c in calls] == [True, True] and ss is None
    monkeypatch.delenv("SN2_DEFER_BN")
    calls.clear()
    ao.run_mlp(mlp, torch.zeros(1000, 11))               # round 2: every block is fused whatever its row count ...
    assert len(calls) == 2
    calls.clear()
    monkeypatch.setenv("SN2_FUSED_MLP_MIN_ROWS", "65536")  # ... unless a floor is asked for (the model's own torch modules)
    ao.run_mlp(mlp, torch.zeros(1000, 11))
    assert calls == []
    monkeypatch.delenv("SN2_FUSED_MLP_MIN_ROWS")
    ao.run_mlp(mlp.eval(), x)                            # eval-mode BatchNorm is never fused
    assert calls == []
    mlp.train()
    odd = MLP([11, 24]).train()                          # a width the kernels do not cover
    assert ao.run_mlp(odd, x).shape == (65536, 24) and calls == []
    with pytest.raises(RuntimeError, match="fixed-capacity"):
        ao.run_mlp(odd, x, rows=torch.zeros(1, dtype=torch.int32))
    rows = torch.zeros(1, dtype=torch.int32)
    ao.run_mlp(mlp, x, rows=rows)
    assert all(c["rows"] is rows for c in calls) and len(calls) == 2


def test_graphed_step_rolls_back_its_warmup():
    """GraphedTrainStep._snapshot / _restore (the roll-back around the warm-up executions of a capture): parameters,
    buffers and pre-existing optimizer state come back, state created by the warm-up returns to zero."""
    from sn2.pipeline import GraphedTrainStep

    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.BatchNorm1d(3))
    opt = torch.optim.Adam(net.parameters(), lr=0.1)
    x = torch.randn(16, 4)

    def one_step():
        opt.zero_grad()
        net(x).pow(2).mean().backward()
        opt.step()

    gs = GraphedTrainStep(net, lambda b: None, optimizer=opt, device=torch.device("cpu"))
    before = {k: v.clone() for k, v in net.state_dict().items()}
    snap = gs._snapshot()                                  # optimizer state still empty
    one_step(); one_step()
    gs._restore(snap)
    for k, v in net.state_dict().items():
        assert torch.equal(v, before[k]), k
    for st in opt.state.values():
        assert all(float(v.abs().sum()) == 0.0 for v in st.values() if torch.is_tensor(v))
    one_step()                                             # now with real state: it must be restored, not zeroed
    mid_params = {k: v.clone() for k, v in net.state_dict().items()}
    mid_state = [{k: (v.clone() if torch.is_tensor(v) else v) for k, v in st.items()} for st in opt.state.values()]
    snap = gs._snapshot()
    one_step(); one_step(); one_step()
    gs._restore(snap)
    for k, v in net.state_dict().items():
        assert torch.equal(v, mid_params[k]), k
    for st, old in zip(opt.state.values(), mid_state):
        for k, v in st.items():
            if torch.is_tensor(v):
                assert torch.equal(v, old[k]), k


def test_sa_recompute_backward_identities():
    """The algebra behind the recompute SA1 block (csrc/train_sa.cu), checked in float64 against autograd on a small
    random problem: the dense gradient dz2 from the raw sums of BatchNorm 2; dW2 = G diag(s1) + db2 t1^T with
    G = dz2^T a1; BatchNorm 1's raw backward sums T1 = W2^T db2 and T2[k] = sum_o W2[o][k] G[o][k] (no reduction sweep
    of their own); da1 from (cA, cB, cC); and the per-point form of layer 1's weight gradient
    dW1 = sum_j du_j in_j^T - sum_i dc_i q_i^T."""
    torch.manual_seed(3)
    dt = torch.float64
    P, Q, C, eps = 40, 12, 16, 1e-5
    deg = torch.randint(1, 9, (Q,))
    rowptr = torch.zeros(Q + 1, dtype=torch.long)
    rowptr[1:] = torch.cumsum(deg, 0)
    E = int(rowptr[-1])
    row = torch.repeat_interleave(torch.arange(Q), deg)
    col = torch.randint(0, P, (E,))
    x = torch.randn(P, 8, dtype=dt)
    pos, qpos = torch.randn(P, 3, dtype=dt), torch.randn(Q, 3, dtype=dt)
    W1 = (torch.randn(16, 11, dtype=dt) * 0.4).requires_grad_(True)
    b1 = (torch.randn(16, dtype=dt) * 0.1).requires_grad_(True)
    W2 = (torch.randn(16, 16, dtype=dt) * 0.4).requires_grad_(True)
    b2 = (torch.randn(16, dtype=dt) * 0.1).requires_grad_(True)
    g1 = (torch.randn(16, dtype=dt) * 0.7 + 0.3).requires_grad_(True)
    g2 = (torch.randn(16, dtype=dt) * 0.7 + 0.3).requires_grad_(True)
    bt1 = torch.zeros(16, dtype=dt, requires_grad=True)
    bt2 = torch.zeros(16, dtype=dt, requires_grad=True)
    dout = torch.randn(Q, C, dtype=dt)

    def bn(a, g, bt):
        mean, var = a.mean(0), a.var(0, unbiased=False)
        inv = 1.0 / torch.sqrt(var + eps)
        return (a - mean) * inv * g + bt, mean, inv

    msg = torch.cat([x[col], pos[col] - qpos[row]], 1)
    a1 = torch.relu(msg @ W1.t() + b1)
    a1.retain_grad()
    h1, mean1, inv1 = bn(a1, g1, bt1)
    a2 = torch.relu(h1 @ W2.t() + b2)
    y2, mean2, inv2 = bn(a2, g2, bt2)
    out = torch.stack([y2[rowptr[i]:rowptr[i + 1]].max(0).values for i in range(Q)])
    arg = torch.stack([y2[rowptr[i]:rowptr[i + 1]].argmax(0) + rowptr[i] for i in range(Q)])
    (out * dout).sum().backward()

    with torch.no_grad():
        n = float(E)

        def coeff(g, mean, inv, S1, S2):  # dy = mask * (cA dz + cB y + cC), as bn_bwd_coeff
            sc = g * inv
            m1, m2 = S1 / n, inv * (S2 - mean * S1) / n
            kb = -inv * m2
            return sc, sc * kb, sc * (-m1 - mean * kb)

        dz = torch.zeros(E, C, dtype=dt)
        dz[arg, torch.arange(C).expand(Q, C)] = dout          # sparse upstream gradient
        amax = a2[arg, torch.arange(C).expand(Q, C)]
        S1, S2 = dout.sum(0), (dout * amax).sum(0)             # B0
        assert torch.allclose(g2.grad, inv2 * (S2 - mean2 * S1)) and torch.allclose(bt2.grad, S1)
        cA, cB, cC = coeff(g2, mean2, inv2, S1, S2)
        dz2 = (a2 > 0) * (cA * dz + cB * a2 + cC)              # B1, per edge
        G, db2 = dz2.t() @ a1, dz2.sum(0)
        s1, t1 = g1 * inv1, bt1 - mean1 * g1 * inv1
        assert torch.allclose(W2.grad, G * s1 + db2[:, None] * t1[None, :])
        assert torch.allclose(b2.grad, db2)
        T1, T2 = W2.t() @ db2, (W2 * G).sum(0)
        assert torch.allclose(bt1.grad, T1) and torch.allclose(g1.grad, inv1 * (T2 - mean1 * T1))
        cA, cB, cC = coeff(g1, mean1, inv1, T1, T2)
        da1 = cA * (dz2 @ W2) + cB * a1 + cC                   # B2, per edge (before relu')
        assert torch.allclose(a1.grad, da1)
        dz1 = (a1 > 0) * da1
        du = torch.zeros(P, C, dtype=dt).index_add_(0, col, dz1)
        dc = torch.zeros(Q, C, dtype=dt).index_add_(0, row, dz1)
        dW1 = du.t() @ torch.cat([x, pos], 1)
        dW1[:, 8:] -= dc.t() @ qpos
        assert torch.allclose(W1.grad, dW1) and torch.allclose(b1.grad, du.sum(0))


def test_ctypes_signatures_mirror_the_header():
    """Every prototype of include/sn2.h against sn2._lib.SIGNATURES, argument by argument: pointers -> c_void_p, int ->
    c_int, long long -> c_longlong, float -> c_float (a drifted binding would pass garbage through the C ABI)."""
    hdr = open(os.path.join(ROOT, "include", "sn2.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    hdr = re.sub(r"//[^\n]*", " ", hdr)
    protos = re.findall(r"\b(?:int|const char \*|size_t|unsigned(?: int)?)\s*\*?\s*(sn2_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", hdr)
    assert len(protos) == len(_lib.SIGNATURES), (len(protos), len(_lib.SIGNATURES))
    kind = {ctypes.c_void_p: "ptr", ctypes.c_int: "int", ctypes.c_longlong: "ll", ctypes.c_float: "float",
            ctypes.c_uint: "uint", ctypes.c_ulonglong: "ull", ctypes.c_size_t: "size", ctypes.c_double: "double"}
    for name, params in protos:
        params = " ".join(params.split())
        want = []
        if params not in ("", "void"):
            for p in params.split(","):
                p = p.strip()
                if "*" in p:
                    want.append("ptr")
                elif re.match(r"(const )?long long\b", p):
                    want.append("ll")
                elif re.match(r"(const )?unsigned long long\b", p):
                    want.append("ull")
                elif re.match(r"(const )?size_t\b", p):
                    want.append("size")
                elif re.match(r"(const )?unsigned\b", p):
                    want.append("uint")
                elif re.match(r"(const )?float\b", p):
                    want.append("float")
                elif re.match(r"(const )?double\b", p):
                    want.append("double")
                elif re.match(r"(const )?int\b", p):
                    want.append("int")
                else:
                    raise AssertionError(f"{name}: unrecognised parameter '{p}'")
        got = [kind[t] for t in _lib.SIGNATURES[name]]
        assert got == want, f"{name}: header {want} vs binding {got}"
