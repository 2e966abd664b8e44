"""CPU tests of the parcel host logic and of the parcel oracle (no GPU): centre grid, shape filter, sampling hash,
the fast hard-medium-vegetation threshold rule against the reference's 10 001-pass scan."""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stratanet2-vegetation-coverage-maps_b200")]

from oracle import parcel_port as pp  # noqa: E402
from sn2 import _lib  # noqa: E402
from sn2.config import default_args  # noqa: E402
from sn2.parcel import keep_points_in_shape, plot_centers_reference  # noqa: E402


def test_sampling_hash_matches_the_library():
    lib = _lib.load(require_cuda=False)
    for seed in (0, 1, 12345, 0xFFFFFFFF):
        j = np.arange(0, 50000, 7)
        want = np.array([lib.sn2_sample_hash(seed, int(x)) for x in j], dtype=np.uint32)
        assert np.array_equal(pp.sample_hash(seed, j), want)
    # well spread: the S smallest of ~10 316 values select about S / n of every tenth of the index range
    h = pp.sample_hash(7, np.arange(10316))
    pick = np.sort(np.argsort(h, kind="stable")[:10000])
    frac = [np.mean((pick >= lo) & (pick < lo + 1031)) for lo in range(0, 10310, 1031)]
    assert max(frac) - min(frac) < 0.01


def test_plot_centres_follow_the_reference_loop():
    """inference/prepare_utils.py:116-144: step 2*cos45*10 - 20/diam_pix, start = min + step/4, ceil(range/step)+1 per
    axis, and the first centre listed twice (the list is seeded with it before the loops)."""
    args = default_args()
    c = plot_centers_reference(0.0, 1040.0, 0.0, 1040.0, args)
    step = 2 * math.cos(math.pi / 4) * 10 - 20 / 20
    n = math.ceil(1040.0 / step) + 1
    assert n == 81 and c.shape == (1 + n * n, 2)
    assert np.array_equal(c[0], c[1]) and np.allclose(c[0], [step / 4, step / 4])
    assert np.allclose(c[2] - c[1], [0.0, step]) and np.allclose(c[1 + n] - c[1], [step, 0.0])  # y runs fastest


def test_keep_points_in_shape_is_a_buffered_polygon_test():
    square = np.array([[20.0, 20.0], [1020.0, 20.0], [1020.0, 1020.0], [20.0, 1020.0]])
    pts = np.array([[500.0, 500.0], [0.0, 500.0], [-11.0, 500.0], [1049.9, 1049.9], [-5.0, -5.0], [1045.0, 1045.0]])
    # 30 m buffer: inside; 20 m from the edge; 31 m from the edge; corner distance 42.3; corner 35.4; corner 35.4
    assert keep_points_in_shape(pts, square, 30.0).tolist() == [True, True, False, False, False, False]
    tri = np.array([[0.0, 0.0], [10.0, 0.0], [0.0, 10.0]])
    assert keep_points_in_shape(np.array([[2.0, 2.0], [6.0, 6.0], [5.5, 5.5]]), tri, 1.0).tolist() == [True, False, True]


def test_hard_medium_vegetation_rule_equals_the_threshold_scan():
    """The kernel's rule -- count(v > lin[i]) from a histogram of rank(v) -- against the reference's 10 001 passes."""
    rng = np.random.default_rng(3)
    img = rng.random((37, 41)) ** 2
    img[rng.random(img.shape) < 0.2] = np.nan
    img[0, :5] = [0.0, 1.0, 0.5, 0.1234, 0.9999]  # values sitting exactly on thresholds
    mosaic = np.stack([rng.random(img.shape), img, rng.random(img.shape), rng.random(img.shape)])
    mosaic[0][rng.random(img.shape) < 0.1] = np.nan
    out, thr, target = pp.insert_hard_med_veg_raster_band(mosaic.copy())
    lin = np.linspace(0, 1, 10001)
    v = img[~np.isnan(img)]
    rank = np.searchsorted(lin, v, side="left")              # number of thresholds strictly below v
    hist = np.bincount(rank, minlength=10002)
    suffix = np.cumsum(hist[::-1])[::-1]                     # suffix[k] = sum_{r >= k} hist[r]
    cnt_gt = suffix[1:10002]                                 # count(v > lin[i]) = sum_{r > i} hist[r]
    delta = np.abs(v.mean() - cnt_gt / v.size)
    assert lin[np.argmin(delta)] == thr
    fin, thr2, _ = pp.finalize_merged_raster(np.concatenate([mosaic, mosaic[3:4], mosaic[3:4]]))
    assert fin.shape[0] == 5 and thr2 == thr
    none = np.isnan(mosaic[:3]).all(axis=0)
    assert np.isnan(fin[:, none]).all() and not np.isnan(fin[:, ~none]).any()


def test_oracle_plot_preparation_shapes_and_rules():
    args = default_args(subsample_size=2048)
    rng = np.random.default_rng(0)
    P = 40000
    cloud = np.zeros((10, P), dtype=np.float32)
    cloud[0] = rng.random(P) * 40 + 700000.0   # Lambert-sized coordinates: float32 ulp = 0.0625 m
    cloud[1] = rng.random(P) * 40 + 6600000.0
    cloud[2] = rng.random(P) * 10 + 100.0
    cloud[3:8] = rng.integers(0, 65535, (5, P))
    cloud[8:10] = rng.integers(1, 7, (2, P))
    centers = np.array([[700020.0, 6600020.0], [700000.0, 6600000.0], [699900.0, 6600000.0]])
    plots, counts = pp.prepare_plots(cloud, centers, args)
    assert plots[2] is None and counts[2] == 0
    d = plots[0]
    assert d["xyz"].shape == (3, 2048) and d["cloud"].shape == (10, 2048) and d["xyz"].dtype == np.float32
    assert (np.diff(d["src"][d["src"] >= 0]) > 0).all()                      # ascending parcel index, order kept
    assert (d["xyz"][2] >= 0).all() and np.abs(d["xyz"][:2]).max() <= 10.0   # height above the local minimum; inside the disk
    fake = d["src"] < 0
    assert np.all(d["cloud"][8:, fake] == np.float32(-1) / np.float32(6))    # rescale_cloud also hits the fake points
    assert np.array_equal(d["cloud"][0], d["xyz"][0] / np.float32(10))
