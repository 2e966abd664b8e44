"""Phase times of the config-4 parcel pass (CUDA events): cloud upload + grid, plot extraction, inference + fusion, finalisation."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stratanet2-vegetation-coverage-maps_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from sn2.parcel import LAS_PARCEL_BUFFER, ParcelCloud, extract_plots, keep_points_in_shape, plot_centers_reference  # noqa: E402

dev = torch.device("cuda", 0)
args, net = bench.make_model(10000, 0)
full = bench.synthetic_parcel(dev)
centers = plot_centers_reference(float(full[0].min()), float(full[0].max()), float(full[1].min()), float(full[1].max()), args)
shape = np.array([[20.0, 20.0], [1020.0, 20.0], [1020.0, 1020.0], [20.0, 1020.0]])
centers = centers[keep_points_in_shape(centers, shape, LAS_PARCEL_BUFFER + args.diam_meters // 2)]


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


for it in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0 = ev()
    parcel = ParcelCloud(full, dev)
    e1 = ev()
    ex = extract_plots(parcel, centers, args)
    e2 = ev()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f"pass {it}: ParcelCloud (copy + extents + grid) {e0.elapsed_time(e1):.2f} ms, extract_plots ({centers.shape[0]} plots) {e1.elapsed_time(e2):.2f} ms, "
          f"host wall {1e3 * (t1 - t0):.1f} ms, valid {int(ex['valid'].sum())}, mean points {float(ex['n_points'].float().mean()):.0f}")
    del ex, parcel
