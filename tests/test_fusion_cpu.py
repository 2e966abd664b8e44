"""CPU tests of the parcel tiling grid and of the oracle's restated fusion rule."""
import numpy as np

from oracle.fusion_port import _merge_pair, fuse_sequential, weight_image
from sn2.fusion import mosaic_frame, plot_centers


def test_plot_center_grid_matches_reference_formula():
    c = plot_centers(0.0, 1040.0, 0.0, 1040.0)
    step = 2 * np.cos(np.pi / 4) * 10 - 1.0
    assert abs(step - 13.142) < 1e-3
    n = int(np.ceil(1040 / step)) + 1
    assert n == 81 and c.shape == (n * n, 2)
    assert np.allclose(c[0], [step / 4, step / 4]) and np.allclose(c[1] - c[0], [0.0, step])


def test_weight_image_and_pairwise_rule():
    w = weight_image(20)
    assert w.shape == (20, 20) and np.isnan(w[0, 0]) and np.isclose(np.nanmax(w), 1.5 - np.hypot(0.025, 0.025))
    assert np.isnan(w).sum() == 400 - 316  # same 316 in-disk pixels as the fake ground points
    a = np.array([[[0.2]], [[1.0]]]); b = np.array([[[0.8]], [[3.0]]])  # [score, weight] single pixel
    out = _merge_pair(a, b)
    assert np.isclose(out[0, 0, 0], (0.2 * 1 + 0.8 * 3) / 4) and np.isclose(out[1, 0, 0], 4.0)
    n = np.array([[[np.nan]], [[np.nan]]])
    assert np.allclose(_merge_pair(n, b), b) and np.isnan(_merge_pair(n, n)).all()


def test_sequential_fusion_is_a_weighted_mean_when_disks_are_full():
    D = 20
    centers = plot_centers(0.0, 30.0, 0.0, 30.0)
    left, top, H, W, off = mosaic_frame(centers)
    rng = np.random.default_rng(1)
    disk = ~np.isnan(weight_image(D))
    r = rng.random((len(centers), 3, D, D)); r[:, :, ~disk] = np.nan
    out = fuse_sequential(r, off, H, W)
    out_rev = fuse_sequential(r[::-1], off[::-1], H, W)
    assert np.allclose(np.nan_to_num(out), np.nan_to_num(out_rev), rtol=1e-12)  # order independent here
    assert np.nanmin(out[:3]) >= 0 and np.nanmax(out[:3]) <= 1
