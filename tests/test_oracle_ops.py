"""CPU tests of the oracle: known-answer micro-cases derived from SURVEY.md Appendix A, and the C
searches against independent numpy restatements."""
import numpy as np
import torch

from oracle import thirdparty_ops as tp
from oracle.pointnet2_port import m_of


def test_fps_tie_lowest_index_and_count_rule():
    pos = torch.tensor([[0., 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]])
    assert tp.fps(pos, None, ratio=0.75).tolist() == [0, 3, 1]  # tie between 1 and 2 -> 1
    assert m_of(10000, 0.25) == 2500 and m_of(2500, 0.25) == 625 and m_of(16384, 0.25) == 4096
    assert m_of(10, 0.25) == 3  # ceil
    assert tp.fps(torch.zeros(8, 3), None, ratio=0.5).tolist() == [0, 0, 0, 0]


def test_fps_c_matches_numpy_and_batches():
    g = torch.Generator().manual_seed(0)
    pos = torch.rand(3 * 400, 3, generator=g)
    pos = torch.round(pos * 20) / 20  # quantised -> many exact ties
    batch = torch.arange(3).repeat_interleave(400)
    got = tp.fps(pos, batch, ratio=0.3)
    m = m_of(400, 0.3)
    assert m == 121  # float32(400) * float32(0.3) = 120.000005 -> ceil: the fp32 count rule of A1
    assert got.numel() == 3 * m
    for b in range(3):
        want = tp.fps_slow(pos[b * 400:(b + 1) * 400].numpy(), m) + b * 400
        assert np.array_equal(got[b * m:(b + 1) * m].numpy(), want)
    got = tp.fps(pos[:400], None, ratio=0.1, start=[7])
    assert np.array_equal(got.numpy(), tp.fps_slow(pos[:400].numpy(), 40, start=7))


def test_radius_strict_cap_order_and_threshold():
    x = torch.tensor([[0., 0, 0], [2, 0, 0], [1.9999999, 0, 0], [0, 0, 2], [5, 5, 5]])
    y = torch.tensor([[0., 0, 0]])
    assert tp.radius(x, y, 2.0, None, None, max_num_neighbors=10)[1].tolist() == [0, 2]
    assert tp.radius(x, y, 2.0, None, None, max_num_neighbors=1)[1].tolist() == [0]  # first K by index
    assert tp.radius_threshold(np.sqrt(2.0)) == 2.0 and tp.radius_threshold(np.sqrt(8.0)) == 8.0
    g = torch.Generator().manual_seed(1)
    pos = torch.rand(600, 3, generator=g) * 4
    q = pos[::5]
    e = tp.radius(pos, q, 0.9, None, None, max_num_neighbors=12)
    sl = tp.radius_slow(pos.numpy(), q.numpy(), 0.9, 12)
    assert np.array_equal(e[1].numpy(), np.concatenate(sl))
    assert np.array_equal(e[0].numpy(), np.repeat(np.arange(len(sl)), [len(s) for s in sl]))
    assert (e[0][1:] >= e[0][:-1]).all()


def test_radius_respects_batches():
    pos = torch.tensor([[0., 0, 0], [0.1, 0, 0], [0., 0, 0], [0.1, 0, 0]])
    batch = torch.tensor([0, 0, 1, 1])
    e = tp.radius(pos, pos, 1.0, batch, batch, max_num_neighbors=8)
    assert e.T.tolist() == [[0, 0], [0, 1], [1, 0], [1, 1], [2, 2], [2, 3], [3, 2], [3, 3]]


def test_knn_tie_rule_and_interpolate():
    x = torch.tensor([[1., 0, 0], [-1, 0, 0], [0, 1, 0], [0, 0, 3]])
    y = torch.tensor([[0., 0, 0]])
    idx, d2 = tp.knn_raw(x, y, 3)
    assert idx.tolist() == [[0, 1, 2]] and d2.tolist() == [[1., 1., 1.]]  # ties -> lower index first
    feats = torch.tensor([[1.], [2.], [6.], [100.]], requires_grad=True)
    out = tp.knn_interpolate(feats, x, y, k=3)
    assert torch.allclose(out, torch.tensor([[3.]]))
    out.sum().backward()
    assert torch.allclose(feats.grad.view(-1), torch.tensor([1 / 3, 1 / 3, 1 / 3, 0.]))
    # coincident source dominates through the 1e-16 clamp
    out = tp.knn_interpolate(torch.tensor([[5.], [1.], [1.]]), torch.tensor([[0., 0, 0], [1, 0, 0], [0, 1, 0]]), y, k=3)
    assert torch.allclose(out, torch.tensor([[5.]]))
    # k=1 from a single source per plot = broadcast (fp3)
    g = torch.tensor([[1.5, -2.0]])
    out = tp.knn_interpolate(g, torch.zeros(1, 3), torch.rand(5, 3) + 0.1, torch.zeros(1, dtype=torch.long), torch.zeros(5, dtype=torch.long), k=1)
    assert torch.allclose(out, g.expand(5, 2), rtol=1e-6)


def test_scatter_max_first_arg_and_backward():
    src = torch.tensor([[1., 5, 5, 2, 7], [0., -1, -1, 3, 3]], requires_grad=True)
    index = torch.tensor([0, 1, 1, 0, 2])
    out, arg = tp.scatter_max(src, index)
    assert out.tolist() == [[2., 5, 7], [3., -1, 3]]
    assert arg.tolist() == [[3, 1, 4], [3, 1, 4]]  # ties -> first index
    (out * torch.tensor([[1., 10, 100], [1000., 1e4, 1e5]])).sum().backward()
    assert src.grad.tolist() == [[0., 10, 0, 1, 100], [0., 1e4, 0, 1000, 1e5]]
    out, arg = tp.scatter_max(torch.tensor([1., 2.]), torch.tensor([0, 2]))
    assert out.tolist() == [1., 0., 2.] and arg.tolist() == [0, 2, 1]  # empty slot -> 0, arg = n
    m = tp.scatter_mean(torch.tensor([1., 3., 5.]), torch.tensor([0, 0, 2]))
    assert m.tolist() == [2., 0., 5.]
    p = tp.global_max_pool(torch.tensor([[1., 2], [3, 0], [-1, -2]]), torch.tensor([0, 0, 1]))
    assert p.tolist() == [[3., 2], [-1., -2]]


def test_pointconv_column_order_and_max():
    lin = torch.nn.Linear(4, 2, bias=False)
    with torch.no_grad():
        lin.weight.copy_(torch.tensor([[1., 0, 0, 0], [0., 1, 10, 100]]))  # col 0 = feature, cols 1..3 = rel pos
    conv = tp.PointConv(lin, add_self_loops=False)
    x = torch.tensor([[2.], [3.]])
    pos_src = torch.tensor([[1., 2, 3], [4., 5, 6]])
    pos_dst = torch.tensor([[1., 1, 1]])
    edge_index = torch.tensor([[0, 1], [0, 0]])
    out = conv(x, (pos_src, pos_dst), edge_index)
    assert out.tolist() == [[3., 3 + 40 + 500]]  # rel = pos_j - pos_i, features first
